"""CPU, world_size 2, gloo: env sharding and the metrics all-reduce (the only collective)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, per_env, out):
    sys.path.insert(0, str(ROOT))
    from mapf_gpt_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, cnt = parallel.shard_range(per_env.shape[0], rank, world)
    red = parallel.reduce_metrics(parallel.local_metric_sums(per_env[first:first + cnt]))
    table = parallel.gather_rows(per_env[first:first + cnt], per_env.shape[0], first)
    if rank == 0:
        red["table_equal"] = bool(np.array_equal(table, per_env))
        out.put(red)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly():
    from mapf_gpt_b200 import parallel
    for total in (1, 7, 256, 1024, 1000):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_metrics_allreduce_equals_single_process_sum():
    from mapf_gpt_b200 import parallel
    rng = np.random.default_rng(0)
    per_env = np.zeros((37, 10))
    per_env[:, 0] = rng.integers(10, 129, 37)
    per_env[:, 1] = rng.integers(0, 2, 37)
    per_env[:, 2] = rng.random(37)
    per_env[:, 3] = rng.integers(100, 9000, 37)
    per_env[:, 4] = rng.integers(10, 129, 37)
    per_env[:, 6] = per_env[:, 0] * 64
    per_env[:, 7] = 64
    per_env[:, 8] = rng.random(37)
    per_env[5, 7] = 0                                    # an unused slot is ignored
    single = parallel.reduce_metrics(parallel.local_metric_sums(per_env))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, per_env, q)) for r in range(2)]
    [p.start() for p in procs]
    red = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert red["episodes"] == single["episodes"] == 36
    assert red.pop("table_equal")                          # gather_rows: every rank sees every episode row
    for k in single:
        assert abs(red[k] - single[k]) < 1e-9, k
