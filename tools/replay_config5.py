#!/usr/bin/env python3
"""BASELINE config[4] substitute (SURVEY §8d "Config 5"): the Rust node cannot be built here (no cargo), so this REPLAYS the
per-node crypto call stream of `fab local` (benchmark/fabfile.py:14-33 with rate 50,000 tx/s, 512 B tx, 4 nodes, batch 15,000 B)
against the C ABI and reports per-call latency.  Per node and per second: ~1,707 Digest calls over ~15.3 kB serialized batches
(mempool/src/processor.rs:30); per round: 1 strict verify (Block::verify, messages.rs:64), 1 verify_batch of 3 votes
(QC::verify, messages.rs:197) and, at the leader, 3 strict verifies (Vote::verify, messages.rs:144).
This is a replay of the call pattern, NOT a fab run."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hotstuff_b200 import Engine
from oracle_api import Oracle

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
o = Oracle(); e = Engine(0)
rng = np.random.default_rng(5)
seeds = rng.integers(0, 256, (4, 32), dtype=np.uint8); pks = o.keygen_batch(seeds)
e.committee_register(pks)
tx = 512; per_batch = 15000 // tx
batch = (0).to_bytes(4, "little") + per_batch.to_bytes(8, "little") + b"".join(tx.to_bytes(8, "little") + rng.bytes(tx) for _ in range(per_batch))
boff = np.array([0, len(batch)], dtype=np.uint64)
lat = {"digest_batch_15kB": [], "verify_strict_1": [], "verify_batch_3": [], "verify_strict_3": []}
def timed(key, fn):
    t0 = time.perf_counter(); r = fn(); lat[key].append((time.perf_counter() - t0) * 1e6); return r
for r in range(rounds):
    bd = o.digest32(b"block" + r.to_bytes(8, "little"))
    recs = np.zeros((4, 128), dtype=np.uint8)
    for i in range(4):
        recs[i, :64] = np.frombuffer(o.sign(seeds[i].tobytes(), bd), np.uint8); recs[i, 64:96] = pks[i]; recs[i, 96:] = np.frombuffer(bd, np.uint8)
    votes = np.concatenate([recs[1:4, 64:96], recs[1:4, :64]], axis=1)
    for _ in range(2):   # ~1.7 batches per round at 1,000 rounds/s
        d = timed("digest_batch_15kB", lambda: e.digest32_batch(batch, boff))
    assert timed("verify_strict_1", lambda: e.verify_strict_batch(recs[:1]))[0]
    assert timed("verify_batch_3", lambda: e.verify_batch_shared_msg(bd, votes))
    assert timed("verify_strict_3", lambda: e.verify_strict_batch(recs[1:4])).all()
import hashlib
assert d[0].tobytes() == hashlib.sha512(batch).digest()[:32]
out = {"what": "replay of the per-node crypto call stream of fab local @ 50k tx/s, 512 B tx, 4 nodes (not a fab run)", "rounds": rounds,
       "latency_us": {k: {"p50": float(np.percentile(v, 50)), "p99": float(np.percentile(v, 99)), "mean": float(np.mean(v)), "calls": len(v)} for k, v in lat.items()}}
per_round = sum(np.mean(v) * (2 if k.startswith("digest") else 1) for k, v in lat.items())
out["crypto_us_per_round"] = float(per_round); out["rounds_per_s_sustainable_single_caller"] = 1e6 / per_round
print(json.dumps(out))
