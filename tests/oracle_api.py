"""ctypes wrapper of the CPU oracle (oracle/libhs_oracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PARSE_OK, R_OK, EQ_OK, SMALL, STRICT = 1, 2, 4, 8, 16
L_ORDER = 2**252 + 27742317777372353535851937790883648493
P = 2**255 - 19


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    def __init__(self):
        from hotstuff_b200 import build
        path = build.build_oracle()
        self.lib = L = ctypes.CDLL(path)
        L.hso_verify_flags.restype = ctypes.c_uint
        L.hso_verify_flags_fast.restype = ctypes.c_uint
        for f in (L.hso_verify_rec128_batch, L.hso_verify_var_batch, L.hso_sign_batch, L.hso_keygen_batch, L.hso_digest32_batch, L.hso_digest32_batch_mt):
            f.restype = None

    def sha512(self, m):
        o = ctypes.create_string_buffer(64)
        self.lib.hso_sha512(bytes(m), ctypes.c_size_t(len(m)), o)
        return o.raw

    def digest32(self, m):
        return self.sha512(m)[:32]

    def keygen(self, seed):
        o = ctypes.create_string_buffer(32)
        self.lib.hso_keygen(bytes(seed), o)
        return o.raw

    def sign(self, seed, m):
        o = ctypes.create_string_buffer(64)
        self.lib.hso_sign(bytes(seed), bytes(m), ctypes.c_size_t(len(m)), o)
        return o.raw

    def flags(self, sig, pk, m, fast=False):
        f = self.lib.hso_verify_flags_fast if fast else self.lib.hso_verify_flags
        return int(f(bytes(sig), bytes(pk), bytes(m), ctypes.c_size_t(len(m))))

    def verify_strict(self, sig, pk, m):
        return bool(self.flags(sig, pk, m, fast=True) & STRICT)

    def keygen_batch(self, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros_like(seeds)
        self.lib.hso_keygen_batch(_vp(seeds), ctypes.c_size_t(seeds.shape[0]), _vp(out))
        return out

    def sign_batch(self, seeds, pks, key_idx, msgs, off, nthreads=None):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint8)
        pks = np.ascontiguousarray(pks, dtype=np.uint8)
        key_idx = np.ascontiguousarray(key_idx, dtype=np.uint32)
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = key_idx.shape[0]
        sigs = np.zeros((n, 64), dtype=np.uint8)
        self.lib.hso_sign_batch(_vp(seeds), _vp(pks), _vp(key_idx), _vp(msgs), _vp(off), ctypes.c_size_t(n), int(nthreads or os.cpu_count() or 1), _vp(sigs))
        return sigs

    def verify_rec128(self, recs, mode=0, nthreads=None):
        recs = np.ascontiguousarray(recs, dtype=np.uint8).reshape(-1, 128)
        n = recs.shape[0]
        bm = np.zeros(max(1, (n + 31) // 32), dtype=np.uint32)
        self.lib.hso_verify_rec128_batch(_vp(recs), ctypes.c_size_t(n), int(mode), int(nthreads or os.cpu_count() or 1), _vp(bm))
        return np.unpackbits(bm.view(np.uint8), bitorder="little")[:n].astype(bool)

    def verify_var(self, sig, pk, msgs, off, mode=0, nthreads=None):
        sig = np.ascontiguousarray(sig, dtype=np.uint8).reshape(-1, 64)
        pk = np.ascontiguousarray(pk, dtype=np.uint8).reshape(-1, 32)
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
        if msgs.size == 0:
            msgs = np.zeros(1, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = sig.shape[0]
        bm = np.zeros(max(1, (n + 31) // 32), dtype=np.uint32)
        self.lib.hso_verify_var_batch(_vp(sig), _vp(pk), _vp(msgs), _vp(off), ctypes.c_size_t(n), int(mode), int(nthreads or os.cpu_count() or 1), _vp(bm))
        return np.unpackbits(bm.view(np.uint8), bitorder="little")[:n].astype(bool)

    def verify_batch_shared_msg(self, digest, votes, nthreads=1):
        votes = np.ascontiguousarray(votes, dtype=np.uint8).reshape(-1, 96)
        n = votes.shape[0]
        bm = np.zeros(max(1, (n + 31) // 32), dtype=np.uint32)
        ok = self.lib.hso_verify_batch_shared_msg(bytes(digest), _vp(votes), ctypes.c_size_t(n), int(nthreads), _vp(bm))
        return bool(ok), np.unpackbits(bm.view(np.uint8), bitorder="little")[:n].astype(bool)

    def digest32_batch(self, data, off, nthreads=1):
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(bytes(data), dtype=np.uint8)
        data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        if data.size == 0:
            data = np.zeros(1, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = off.shape[0] - 1
        out = np.zeros((n, 32), dtype=np.uint8)
        self.lib.hso_digest32_batch_mt(_vp(data), _vp(off), ctypes.c_size_t(n), int(nthreads), _vp(out))
        return out

    # point helpers for adversarial fixtures
    def point_add(self, a, b):
        o = ctypes.create_string_buffer(32)
        return o.raw if self.lib.hso_point_add(bytes(a), bytes(b), o) else None

    def scalarmult(self, s, pt):
        o = ctypes.create_string_buffer(32)
        return o.raw if self.lib.hso_point_scalarmult(int(s).to_bytes(32, "little"), bytes(pt), o) else None

    def decompress_ok(self, enc):
        return bool(self.lib.hso_point_decompress_ok(bytes(enc)))

    def is_small_order(self, enc):
        return int(self.lib.hso_point_is_small_order(bytes(enc)))

    def sc_reduce64(self, h):
        o = ctypes.create_string_buffer(32)
        self.lib.hso_sc_reduce64(bytes(h), o)
        return int.from_bytes(o.raw, "little")


# ---------------------------------------------------------------------------------------------------- workloads
def make_workload(oracle, n, n_keys=64, msg_len=32, seed=1, corrupt_frac=0.0, nthreads=None):
    """Deterministic synthetic (sig, pk, msg) set: keys/messages from a seeded PRNG, RFC 8032 signatures from the
    oracle, then `corrupt_frac` of the records get one flipped bit at a seeded position (uniform over sig|pk|msg).
    Returns dict(sig (n,64), pk (n,32), msgs (n*msg_len,), off (n+1,), corrupted bool[n])."""
    rng = np.random.default_rng(seed)
    seeds = rng.integers(0, 256, size=(n_keys, 32), dtype=np.uint8)
    pks = oracle.keygen_batch(seeds)
    key_idx = (np.arange(n) % n_keys).astype(np.uint32)
    msgs = rng.integers(0, 256, size=(n, msg_len), dtype=np.uint8)
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(msg_len))
    sig = oracle.sign_batch(seeds, pks, key_idx, msgs.reshape(-1), off, nthreads)
    pk = pks[key_idx].copy()
    corrupted = np.zeros(n, dtype=bool)
    if corrupt_frac > 0:
        k = max(1, int(n * corrupt_frac))
        pos = rng.choice(n, size=k, replace=False)
        for i in pos:
            total_bits = (64 + 32 + msg_len) * 8
            b = int(rng.integers(0, total_bits))
            byte, bit = b >> 3, b & 7
            if byte < 64:
                sig[i, byte] ^= 1 << bit
            elif byte < 96:
                pk[i, byte - 64] ^= 1 << bit
            else:
                msgs[i, byte - 96] ^= 1 << bit
            corrupted[i] = True
    return dict(sig=sig, pk=pk, msgs=msgs.reshape(-1), off=off, corrupted=corrupted, key_idx=key_idx, pks=pks, seeds=seeds)


def to_rec128(w):
    n = w["sig"].shape[0]
    recs = np.zeros((n, 128), dtype=np.uint8)
    recs[:, :64] = w["sig"]
    recs[:, 64:96] = w["pk"]
    recs[:, 96:] = w["msgs"].reshape(n, 32)
    return recs


def make_adversarial(oracle, n, seed=1):
    """Randomised adversarial (sig, pk, 32-byte msg) records for differential testing: valid signatures mutated along the
    axes where Ed25519 implementations disagree — torsion components added to A and/or R (with the signature recomputed
    for the mixed key so the cofactorless equation may or may not hold), non-canonical S (S + l), small-order A / R,
    random byte strings as A / R (about half do not decompress), non-canonical field encodings, bit flips."""
    import hashlib
    rng = np.random.default_rng(seed)
    Y8 = 0x05fc536d880238b13933c6d305acdfd5f098eff289f4c345b027b2c28f95e826
    tors = []
    for ty in (0, 1, P - 1, P, P + 1, Y8, P - Y8):
        for sign in (0, 1):
            e = bytearray(int(ty).to_bytes(32, "little"))
            e[31] |= sign << 7
            tors.append(bytes(e))
    B_enc = int("6666666666666666666666666666666666666666666666666666666666666658", 16).to_bytes(32, "little")
    nk = 8
    seeds = [rng.bytes(32) for _ in range(nk)]
    pks = [oracle.keygen(s) for s in seeds]
    scal = []
    for s in seeds:
        h = hashlib.sha512(s).digest()
        a = int.from_bytes(bytes([h[0] & 248]) + h[1:31] + bytes([(h[31] & 127) | 64]), "little")
        scal.append((a, h[32:]))
    recs = np.zeros((n, 128), dtype=np.uint8)
    kinds = rng.integers(0, 10, n)
    for i in range(n):
        k = int(rng.integers(0, nk))
        m = rng.bytes(32)
        a, prefix = scal[k]
        A = pks[k]
        kind = int(kinds[i])
        if kind in (1, 2):                      # mixed-order key (and sometimes nonce): sign for A' = A + T
            A = oracle.point_add(pks[k], tors[int(rng.integers(0, len(tors)))])
        r = int.from_bytes(hashlib.sha512(prefix + m).digest(), "little") % L_ORDER
        R = oracle.scalarmult(r, B_enc)
        if kind == 2:
            R = oracle.point_add(R, tors[int(rng.integers(0, len(tors)))])
        kk = oracle.sc_reduce64(hashlib.sha512(R + A + m).digest())
        S = (r + kk * a) % L_ORDER
        sig = bytearray(R + int(S).to_bytes(32, "little"))
        pk = bytearray(A)
        if kind == 3 and S + L_ORDER < 2**256:  # non-canonical S
            sig[32:] = int(S + L_ORDER).to_bytes(32, "little")
        elif kind == 4:                         # small-order A or R
            if rng.integers(0, 2):
                pk = bytearray(tors[int(rng.integers(0, len(tors)))])
            else:
                sig[:32] = tors[int(rng.integers(0, len(tors)))]
        elif kind == 5:                         # random bytes as A or R
            if rng.integers(0, 2):
                pk = bytearray(rng.bytes(32))
            else:
                sig[:32] = rng.bytes(32)
        elif kind == 6:                         # single bit flip anywhere
            b = int(rng.integers(0, 128 * 8))
            buf = bytearray(bytes(sig) + bytes(pk) + m)
            buf[b >> 3] ^= 1 << (b & 7)
            sig, pk, m = buf[:64], buf[64:96], bytes(buf[96:])
        elif kind == 7:                         # high / non-canonical field bits
            which = int(rng.integers(0, 3))
            if which == 0:
                pk[31] ^= 0x80
            elif which == 1:
                sig[31] ^= 0x80
            else:
                sig[63] |= int(rng.integers(1, 8)) << 5
        elif kind == 8:                         # identity key: R = [S]B verifies for any message in batch-eq mode
            pk = bytearray((1).to_bytes(32, "little"))
            sig = bytearray(oracle.scalarmult(r, B_enc) + int(r).to_bytes(32, "little"))
        recs[i, :64] = np.frombuffer(bytes(sig), np.uint8)
        recs[i, 64:96] = np.frombuffer(bytes(pk), np.uint8)
        recs[i, 96:] = np.frombuffer(m, np.uint8)
    return recs
