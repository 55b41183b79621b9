#!/usr/bin/env python3
"""Small driver for ncu: a few passes of the bench step (Digest + verify) on an n-record resident workload."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from hotstuff_b200 import Engine
from oracle_api import Oracle, make_workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
mode = sys.argv[2] if len(sys.argv) > 2 else "committee"
L = 512
o = Oracle(); e = Engine(0)
base = min(n, 1 << 16)            # a signed base set, tiled to n (input synthesis is not what is profiled)
w = make_workload(o, base, n_keys=4096, seed=3, msg_len=L)
d = o.digest32_batch(w["msgs"], w["off"], nthreads=16)
sig = o.sign_batch(w["seeds"], w["pks"], w["key_idx"], d.reshape(-1), np.arange(base + 1, dtype=np.uint64) * 32)
reps = (n + base - 1) // base
sig = np.tile(sig, (reps, 1))[:n]
w["key_idx"] = np.tile(w["key_idx"], reps)[:n]
w["msgs"] = np.tile(w["msgs"].reshape(base, L), (reps, 1))[:n].reshape(-1)
dev = torch.device("cuda", 0)
d_sig = torch.from_numpy(sig).to(dev); d_pk = torch.from_numpy(w["pks"][w["key_idx"]]).to(dev)
d_msgs = torch.from_numpy(w["msgs"]).to(dev); d_vidx = torch.from_numpy(w["key_idx"].astype(np.int32)).to(dev)
d_dig = torch.empty((n, 32), dtype=torch.uint8, device=dev); d_bm = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
if mode != "generic": e.committee_register(w["pks"])
for _ in range(3):
    e.verify_msgs_dev(d_sig, d_msgs, L, d_dig, d_bm, n, d_pk=None if mode == "indexed" else d_pk, d_vidx=d_vidx if mode == "indexed" else None)
torch.cuda.synchronize()
bits = np.unpackbits(d_bm.cpu().numpy().view(np.uint8), bitorder="little")[:n]
print("accepted", int(bits.sum()), "of", n)
