#!/usr/bin/env python3
"""Small driver for ncu: a few launches of the verify kernels on a 2^17-record resident workload."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from hotstuff_b200 import Engine
from oracle_api import Oracle, make_workload, to_rec128
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["committee", "generic"]
o = Oracle(); e = Engine(0)
w = make_workload(o, n, n_keys=4096, seed=3, corrupt_frac=0.01)
recs = to_rec128(w)
dev = torch.device("cuda", 0)
d_recs = torch.from_numpy(recs).to(dev); d_bm = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
d_sig = torch.from_numpy(w["sig"]).to(dev); d_vidx = torch.from_numpy(w["key_idx"].astype(np.int32)).to(dev)
d_dig = torch.from_numpy(recs[:, 96:].copy()).to(dev); d_midx = torch.arange(n, dtype=torch.int32, device=dev)
e.committee_register(w["pks"])
for _ in range(3):
    if "committee" in modes: e.verify_committee_dev(d_vidx, d_sig, d_dig, d_bm, n, d_midx=d_midx)
    if "generic" in modes: e.verify_rec128_dev(d_recs, d_bm, n)
torch.cuda.synchronize()
print("done")
