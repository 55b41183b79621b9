#!/usr/bin/env python3
"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck): small tables (base window 12), a few hundred
records per entry point.  usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hotstuff_b200 import Engine
from oracle_api import Oracle, make_workload, to_rec128, make_adversarial

o = Oracle()
e = Engine(0, base_window=12, key_window=10)
w = make_workload(o, 700, n_keys=9, seed=1, corrupt_frac=0.1)
recs = np.concatenate([to_rec128(w), make_adversarial(o, 300, seed=2)])
want = o.verify_rec128(recs)
assert (e.verify_rec128(recs) == want).all()                       # generic kernel + key-cache collection
assert (e.verify_rec128(recs) == want).all()                       # learned tables: lookup + committee kernel + side pass
keys, inv = np.unique(recs[:, 64:96], axis=0, return_inverse=True)
e.committee_register(keys)
assert (e.verify_rec128(recs) == want).all()
assert (e.verify_rec128(recs[:37]) == want[:37]).all()             # latency path (k_verify_small)
assert (e.verify_committee(inv.astype(np.uint32), recs[:, :64].copy(), recs[:, 96:].copy(), msg_idx=np.arange(len(recs), dtype=np.uint32)) == want).all()
rng = np.random.default_rng(3)
lens = [0, 1, 111, 112, 513, 15300, 4096, 128 * 33]
off = np.zeros(len(lens) + 1, dtype=np.uint64); off[1:] = np.cumsum(lens)
data = rng.integers(0, 256, int(off[-1]), dtype=np.uint8)
import hashlib
d = e.digest32_batch(data, off)                                     # long-message kernel
assert all(d[i].tobytes() == hashlib.sha512(data[int(off[i]):int(off[i + 1])].tobytes()).digest()[:32] for i in range(len(lens)))
lens = list(range(0, 200, 7)); off = np.zeros(len(lens) + 1, dtype=np.uint64); off[1:] = np.cumsum(lens)
d = e.digest32_batch(data[:int(off[-1])], off)                      # generic digest kernel
for L in (512, 144):                                                # staged fixed-length digest + verify
    n = 333
    msgs = rng.integers(0, 256, (n, L), dtype=np.uint8)
    dg = o.digest32_batch(msgs.reshape(-1), np.arange(n + 1, dtype=np.uint64) * L)
    kidx = (np.arange(n) % 9).astype(np.uint32)
    sig = e.sign_digests(w["seeds"], e.keygen_batch(w["seeds"]), dg, key_idx=kidx)   # signer kernels
    assert (sig == o.sign_batch(w["seeds"], w["pks"], kidx, dg.reshape(-1), np.arange(n + 1, dtype=np.uint64) * 32)).all()
    assert e.verify_msgs(sig, msgs.reshape(-1), L, pk=w["pks"][kidx]).all()
# QC / TC / group front ends
pre = np.zeros((5, 40), np.uint8); pre[:, 0] = np.arange(5)
qd = o.digest32_batch(pre.reshape(-1), np.arange(6, dtype=np.uint64) * 40)
qi = np.repeat(np.arange(5, dtype=np.uint32), 9); vi = np.tile(np.arange(9, dtype=np.uint32), 5)
sg = o.sign_batch(w["seeds"], w["pks"], vi, qd[qi].reshape(-1), np.arange(46, dtype=np.uint64) * 32)
sg[7, 3] ^= 1
assert list(e.verify_qcs(pre, sg, qi, pk=w["pks"][vi])) == [False, True, True, True, True]
hq = np.arange(45, dtype=np.uint64); tr = np.arange(5, dtype=np.uint64) + 100
tp = b"".join(int(tr[t]).to_bytes(8, "little") + int(h).to_bytes(8, "little") for t, h in zip(qi, hq))
td = o.digest32_batch(tp, np.arange(46, dtype=np.uint64) * 16)
ts = o.sign_batch(w["seeds"], w["pks"], vi, td.reshape(-1), np.arange(46, dtype=np.uint64) * 32)
assert e.verify_tcs(tr, ts, hq, tc_idx=qi, pk=w["pks"][vi]).all()
poff = np.arange(6, dtype=np.uint64) * 40
assert list(e.verify_groups(pre.reshape(-1), poff, sg, qi, qi, 5, mode=np.ones(45, np.uint8), pk=w["pks"][vi])) == [False, True, True, True, True]
e.close()
print("sanitize_run ok")
