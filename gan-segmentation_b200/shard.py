"""Index sharding of a generation job over ranks (one process per GPU, no collective on the data path).

The unit of work is a latent with a *global* index g: its z and its noise planes come from Philox streams
keyed by (seed, g), so the images do not depend on how the index range is split -- the property the
reference gets implicitly from one in-process RNG (image_generator.py:92-95) and loses with >1 process.
"""
from __future__ import annotations


def shard_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of range(n_total) for ``rank``; sizes differ by at most one
    (split_and_load(even_split=False) semantics, image_generator.py:95)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def batches(lo, hi, batch):
    """[(first_sample, n)] covering [lo, hi) in steps of ``batch`` (last one may be short, :93)."""
    return [(s, min(batch, hi - s)) for s in range(lo, hi, batch)]


def step_first_sample(step, rank, world, batch):
    """bench.py's weak-scaling schedule: in step s rank r generates samples [(s*world + r)*batch, +batch)."""
    return (step * world + rank) * batch


def reduce_max_time(ms, group=None):
    """Max over ranks of a per-rank elapsed time (what bench.py reports for N > 1)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(ms)], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend(group) == 'nccl':
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
