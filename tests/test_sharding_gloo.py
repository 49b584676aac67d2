"""world_size-2 CPU test (gloo) of the N>1 path of bench.py: batch shards are independent, the only
cross-rank operation is the MAX of the timings, and the aggregate is world x per-rank bytes / max time."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    calls, shapes, starts, F = bench.make_calls(1, seed=bench.shard_seed(rank), layers=1,
                                               modalities=(("det", 32, 13),))
    ms = bench.max_over_ranks(10.0 * (rank + 1), torch.device("cpu"))          # rank 1 is the slow one
    gbs = bench.whole_job_gbs(world, 1_000_000, ms)
    digest = float(np.abs(calls[0]["loc"]).sum())
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, ms, gbs, digest, F))
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


def test_two_rank_sharding_and_aggregation():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ms0, g0, d0, F0), (r1, ms1, g1, d1, F1) = sorted(res)
    assert ms0 == ms1 == 20.0                                   # MAX over ranks
    assert g0 == g1 == 2 * 1_000_000 / 20e-3 / 1e9               # whole-job aggregate, weak scaling
    assert d0 != d1                                              # each rank owns a different batch shard
    assert F0 == F1 == 112200
