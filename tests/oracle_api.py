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
