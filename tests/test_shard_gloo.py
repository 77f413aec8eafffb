"""world_size-2 gloo test (CPU) of the multi-process path bench.py uses: contiguous event shards per
rank, no data-path collective, counters summed and time max-reduced over ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_events, q):
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module("nps-waveform-analysis_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = pkg.shard_range(n_events, rank, world)
    # stand-in for the per-rank GPU work: a deterministic per-event quantity
    local = np.arange(lo, hi, dtype=np.int64)
    counters = torch.tensor([hi - lo, int((local % 7 == 0).sum())], dtype=torch.int64)
    t_local = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((counters.tolist(), float(t_local[0]), (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    n_events, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_events, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    counters, tmax, r0 = q.get()
    assert counters[0] == n_events                       # shards cover every event exactly once
    assert counters[1] == len([e for e in range(n_events) if e % 7 == 0])
    assert abs(tmax - 0.020) < 1e-12                     # max over ranks
    assert r0 == (0, 500)
