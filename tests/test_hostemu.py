"""Host emulation of the device headers (tests/hostemu): the exact field / scalar / SHA-512 / curve / window logic the
CUDA kernels run, compiled by g++ with the PTX primitives replaced by portable C, checked against Python big integers,
hashlib and the oracle.  Runs without a GPU; the GPU tests then only have to establish that the PTX primitives agree."""
import ctypes
import hashlib

import numpy as np

from oracle_api import L_ORDER, P, R_OK


def _fe(hostemu, op, a, b=0):
    o = ctypes.create_string_buffer(32)
    hostemu.emu_fe_op(op, int(a).to_bytes(32, "little"), int(b).to_bytes(32, "little"), o)
    return int.from_bytes(o.raw, "little")


def test_field_ops(hostemu):
    rng = np.random.default_rng(1)
    specials = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P, 2 * P + 1, 2**256 - 1, 2**256 - 38, 2**256 - 39, 2**255, 2**255 - 1]
    vals = specials + [int.from_bytes(rng.bytes(32), "little") for _ in range(200)]
    for a in vals:
        for b in vals[:18]:
            assert _fe(hostemu, 0, a, b) % P == a * b % P
            assert _fe(hostemu, 2, a, b) % P == (a + b) % P
            assert _fe(hostemu, 3, a, b) % P == (a - b) % P
        assert _fe(hostemu, 1, a) % P == a * a % P
        assert _fe(hostemu, 4, a) == a % P
        assert _fe(hostemu, 7, a) % P == (-a) % P
    for a in vals[:30]:
        assert _fe(hostemu, 5, a) % P == pow(a, P - 2, P)
        assert _fe(hostemu, 6, a) % P == pow(a, (P - 5) // 8, P)


def test_scalar_reduce_and_recode(hostemu):
    rng = np.random.default_rng(2)
    xs = [int.from_bytes(rng.bytes(64), "little") for _ in range(500)]
    xs += [0, L_ORDER - 1, L_ORDER, L_ORDER + 1, 2**512 - 1, (2**512 // L_ORDER) * L_ORDER, (2**512 // L_ORDER) * L_ORDER - 1, L_ORDER << 259]
    xs += [2**512 - 1 - int.from_bytes(rng.bytes(20), "little") for _ in range(50)]
    for x in xs:
        x %= 2**512
        o = ctypes.create_string_buffer(32)
        hostemu.emu_sc_reduce512(x.to_bytes(64, "little"), o)
        assert int.from_bytes(o.raw, "little") == x % L_ORDER
    for s in [0, 1, L_ORDER - 1, L_ORDER, L_ORDER + 1, 2**252, 2**253, 2**256 - 1]:
        assert hostemu.emu_sc_is_canonical(s.to_bytes(32, "little")) == (1 if s < L_ORDER else 0)
    edge = [0, L_ORDER - 1, 2**253 - 1, 2**252, int("7f" * 31, 16), int("80" * 31, 16), int("77" * 32, 16) % 2**253, int("88" * 32, 16) % 2**253]
    for W, msb in [(4, 1), (4, 0)] + [(w, 0) for w in range(8, 27)]:   # every width the engine can be configured with (11 and 23 divide 253)
        for s in edge + [int(rng.integers(0, 2**62)) * int(rng.integers(0, 2**62)) ** 3 % L_ORDER for _ in range(200)]:
            out = (ctypes.c_int * 80)()
            n = hostemu.emu_sc_digits(W, msb, s.to_bytes(32, "little"), out)
            d = list(out)[:n]
            assert all(-(1 << (W - 1)) <= x < (1 << (W - 1)) for x in d)
            assert sum(x << (W * i) for i, x in enumerate(d)) == s


def test_sha512_paths(hostemu):
    rng = np.random.default_rng(3)
    for ln in list(range(0, 260)) + [511, 512, 513, 1000]:
        m = rng.bytes(ln)
        o = ctypes.create_string_buffer(64)
        for pad in (0, 1, 4, 7):  # aligned and unaligned message pointers
            buf = ctypes.create_string_buffer(b"\0" * pad + m + b"x" * 16)
            hostemu.emu_sha512(ctypes.byref(buf, pad), ctypes.c_uint64(ln), o)
            assert o.raw == hashlib.sha512(m).digest(), (ln, pad)
        R, A = rng.bytes(32), rng.bytes(32)
        hostemu.emu_sha512_ram(R, A, m, ctypes.c_uint64(ln), o)
        assert o.raw == hashlib.sha512(R + A + m).digest(), ln


def test_sha512_schedule_table_paths(hostemu):
    """The two table-driven compressions of the Digest kernels: per-block K+W tables (k_digest32_long) and the
    host-expanded padding block of messages whose length is a multiple of 128 (k_digest32_fixed)."""
    rng = np.random.default_rng(33)
    o = ctypes.create_string_buffer(64)
    for ln in [0, 1, 111, 112, 127, 128, 129, 4095, 4096, 4097, 15300, 15301, 128 * 33, 128 * 64 + 5]:
        m = rng.bytes(ln)
        hostemu.emu_sha512_kw_path(m, ctypes.c_uint64(ln), o)
        assert o.raw == hashlib.sha512(m).digest(), ln
    for ln in [128, 256, 512, 1024, 128 * 33]:
        m = rng.bytes(ln)
        hostemu.emu_sha512_padkw_path(m, ctypes.c_uint64(ln), o)
        assert o.raw == hashlib.sha512(m).digest(), ln


def test_decompress_and_small_order(hostemu, oracle, golden):
    rng = np.random.default_rng(4)
    encs = [rng.bytes(32) for _ in range(300)] + [bytes.fromhex(t) for t in golden["torsion_encodings"]]
    encs += [int(y).to_bytes(32, "little") for y in list(range(0, 30)) + list(range(P - 30, P + 19))]
    d = (-121665 * pow(121666, P - 2, P)) % P
    for e in encs:
        x = ctypes.create_string_buffer(32)
        y = ctypes.create_string_buffer(32)
        ok = hostemu.emu_decompress(e, x, y)
        assert bool(ok) == oracle.decompress_ok(e), e.hex()
        if ok:
            xi, yi = int.from_bytes(x.raw, "little"), int.from_bytes(y.raw, "little")
            assert (-xi * xi + yi * yi - 1 - d * xi * xi * yi * yi) % P == 0
            assert yi == (int.from_bytes(e, "little") & (2**255 - 1)) % P
            assert xi == 0 or (xi & 1) == e[31] >> 7
        so = oracle.is_small_order(e)
        assert hostemu.emu_enc_is_small_order(e) == (1 if so == 1 else 0) or so == -1 and not ok, e.hex()


import pytest


@pytest.mark.parametrize("wa,wb", [(10, 12), (8, 8), (12, 16), (9, 14), (11, 12), (13, 10)])
def test_golden_vectors_generic_and_committee_paths(hostemu, golden, wa, wb):
    hostemu.emu_set_windows(wa, wb)
    for v in golden["vectors"]:
        sig, pk, msg = bytes.fromhex(v["sig"]), bytes.fromhex(v["pk"]), bytes.fromhex(v["msg"])
        want = v["flags"] & ~R_OK
        got = hostemu.emu_verify_generic(sig, pk, msg, ctypes.c_uint64(len(msg)))
        assert got == want, (v["name"], got, want)
    # committee path builds a 384 KB table per key on the CPU: sample the interesting ones
    names = ("rfc8032", "reference")
    sample = [v for v in golden["vectors"] if v["group"] in names + ("speccheck",) or v["name"].startswith(("adv_valid", "adv_flip_R", "adv_S_", "adv_torsion_pair_eq_1", "adv_identity_A", "adv_mixed_order_A", "adv_A_not_on_curve_0", "adv_zero"))]
    for v in sample:
        sig, pk, msg = bytes.fromhex(v["sig"]), bytes.fromhex(v["pk"]), bytes.fromhex(v["msg"])
        got = hostemu.emu_verify_committee(sig, pk, msg, ctypes.c_uint64(len(msg)))
        assert got == v["flags"] & ~R_OK, (v["name"], got)
        if len(msg) == 32 and wa <= 10:   # latency path (tree of lanes + projective compare with the decompressed R)
            assert hostemu.emu_verify_committee_tree(sig, pk, msg, ctypes.c_uint64(32)) == v["flags"] & ~R_OK, v["name"]


def test_random_parity_with_oracle(hostemu, oracle):
    rng = np.random.default_rng(6)
    for i in range(40):
        seed = rng.bytes(32)
        pk = oracle.keygen(seed)
        m = rng.bytes(32 if i % 2 else int(rng.integers(0, 300)))
        sig = oracle.sign(seed, m)
        for s, p, mm in ((sig, pk, m), (bytes([sig[0] ^ 1]) + sig[1:], pk, m), (sig, bytes([pk[0] ^ 2]) + pk[1:], m), (sig[:40] + bytes([sig[40] ^ 8]) + sig[41:], pk, m)):
            assert hostemu.emu_verify_generic(s, p, mm, ctypes.c_uint64(len(mm))) == oracle.flags(s, p, mm) & ~R_OK


def test_randomised_adversarial_differential_hostemu(hostemu, oracle):
    """Same randomised adversarial generator as the GPU test, at a size the host emulation finishes quickly."""
    from oracle_api import make_adversarial
    hostemu.emu_set_windows(10, 12)
    recs = make_adversarial(oracle, 600, seed=7)
    agree = 0
    for r in recs:
        sig, pk, m = r[:64].tobytes(), r[64:96].tobytes(), r[96:].tobytes()
        want = oracle.flags(sig, pk, m, fast=True) & ~R_OK
        assert hostemu.emu_verify_generic(sig, pk, m, ctypes.c_uint64(32)) == want
        agree += 1
    assert agree == 600


def test_load_generation_signer_matches_rfc8032(hostemu, oracle, golden):
    """sign_digest_core / keygen_core (the GPU load-generation signer) under host emulation: byte-identical to the oracle's RFC 8032
    signer (itself pinned on the RFC 8032 KATs and OpenSSL) for the reference's keys() seeds and random seeds, 32-byte digests."""
    hostemu.emu_set_windows(10, 12)
    rng = np.random.default_rng(8032)
    seeds = [bytes.fromhex(s) for s in golden["reference"]["seeds"]] + [rng.bytes(32) for _ in range(40)]
    for i, seed in enumerate(seeds):
        m = oracle.digest32(b"msg" + bytes([i]))
        pk, sig = ctypes.create_string_buffer(32), ctypes.create_string_buffer(64)
        hostemu.emu_sign_digest(seed, m, pk, sig)
        assert pk.raw == oracle.keygen(seed), i
        assert sig.raw == oracle.sign(seed, m), i
    assert golden["reference"]["hello_sig_key3"] == _emu_sign(hostemu, seeds[3], bytes.fromhex(golden["reference"]["hello_digest"]))


def _emu_sign(hostemu, seed, m):
    pk, sig = ctypes.create_string_buffer(32), ctypes.create_string_buffer(64)
    hostemu.emu_sign_digest(seed, m, pk, sig)
    return sig.raw.hex()
