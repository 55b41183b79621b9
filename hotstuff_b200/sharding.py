"""Multi-GPU sharding of independent signatures: one process per GPU, contiguous index ranges whose
length is a multiple of 32 so bitmap words never straddle ranks, and ONE collective — an all-gather of the per-rank
accept bitmaps — so that every rank (every validator process) ends up with the full bitmap.  Backend-agnostic
(`nccl` on the GPUs, `gloo` in the CPU tests)."""
import numpy as np


def shard_range(n, rank, world):
    """Records [lo, hi) owned by `rank`; every shard but the last has ceil(n / world) rounded up to 32 records."""
    per = -(-n // world)
    per = (per + 31) // 32 * 32
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def all_gather_bitmap(local_words, n, world, group=None):
    """local_words: int32/uint32 tensor with the shard's bitmap (ceil(shard/32) words).  Returns the full bitmap
    (ceil(n/32) words) on every rank.  Shards are padded to the common per-rank word count for the collective."""
    import torch
    import torch.distributed as dist
    per_words = shard_range(n, 0, world)[2] // 32
    buf = torch.zeros(per_words, dtype=local_words.dtype, device=local_words.device)
    buf[: local_words.numel()] = local_words
    if world == 1:
        return buf[: (n + 31) // 32]
    out = torch.empty(per_words * world, dtype=local_words.dtype, device=local_words.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out[: (n + 31) // 32]


def verify_sharded(verify_fn, n, rank, world, device=None, group=None):
    """verify_fn(lo, hi) -> uint32 numpy array / tensor of ceil((hi-lo)/32) words for records [lo, hi).
    Returns bool[n] (numpy) assembled from every rank's shard."""
    import torch
    lo, hi, _ = shard_range(n, rank, world)
    words = verify_fn(lo, hi) if hi > lo else np.zeros(0, dtype=np.uint32)
    if isinstance(words, np.ndarray):
        words = torch.from_numpy(words.view(np.int32).copy())
    if device is not None:
        words = words.to(device)
    full = all_gather_bitmap(words, n, world, group)
    return np.unpackbits(full.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
