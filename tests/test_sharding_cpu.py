"""N>1 host logic on CPU: world_size-2 gloo processes shard a generation job with no data-path collective;
the union of the shards is the whole index range, and the timing reduction is the max over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gan_segmentation_b200.shard import shard_range, batches, step_first_sample, reduce_max_time


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, batch, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    mine = [g for first, n in batches(lo, hi, batch) for g in range(first, first + n)]
    # gather only the bookkeeping (index lists), never sample data
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = reduce_max_time(10.0 + rank)
    steps = [step_first_sample(s, rank, world, batch) for s in range(3)]
    all_steps = [None] * world
    dist.all_gather_object(all_steps, steps)
    if rank == 0:
        out.put((gathered, t, all_steps))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    world, n_total, batch = 2, 37, 8
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, t, all_steps = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = [g for part in gathered for g in part]
    assert sorted(flat) == list(range(n_total)) and len(set(flat)) == n_total
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1
    assert t == 11.0                                   # max over ranks
    starts = sorted(s for part in all_steps for s in part)
    assert starts == [i * batch for i in range(3 * world)]   # steps tile the index space without overlap


def test_shard_range_properties():
    for n in (0, 1, 7, 10000):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert batches(3, 20, 8) == [(3, 8), (11, 8), (19, 1)]
