"""N > 1 path on CPU: world_size-2 `gloo` run of the sharding + bitmap all-gather logic (hotstuff_b200/sharding.py), with
the oracle standing in for the per-rank verifier (the GPU engine plugs into the same verify_fn slot in bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    from oracle_api import Oracle, make_workload, to_rec128
    from hotstuff_b200.sharding import shard_range, verify_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle()
    w = make_workload(o, n, n_keys=9, seed=77, corrupt_frac=0.1, nthreads=2)   # identical on every rank (seeded)
    recs = to_rec128(w)

    def verify_fn(lo, hi):
        bits = o.verify_rec128(recs[lo:hi], nthreads=2)
        padded = np.concatenate([bits, np.zeros((-len(bits)) % 32, dtype=bool)])
        return np.frombuffer(np.packbits(padded, bitorder="little").tobytes(), dtype=np.uint32).copy()

    got = verify_sharded(verify_fn, n, rank, world)
    want = o.verify_rec128(recs, nthreads=2)
    lo, hi, per = shard_range(n, rank, world)
    assert per % 32 == 0 and (hi - lo) <= per
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.array([(got == want).all(), (~want).sum() > 0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 64, 33])
def test_gloo_world2_sharded_bitmap_allgather(tmp_path, n):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        res = np.load(os.path.join(str(tmp_path), "ok_%d.npy" % r))
        assert res[0], "rank %d assembled a wrong bitmap" % r


def test_shard_ranges_cover_and_align():
    from hotstuff_b200.sharding import shard_range
    for n in (0, 1, 31, 32, 33, 1000, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 4, 8):
            covered = 0
            for r in range(world):
                lo, hi, per = shard_range(n, r, world)
                assert lo % 32 == 0 or lo == n
                assert lo == min(n, r * per)
                covered += hi - lo
            assert covered == n
