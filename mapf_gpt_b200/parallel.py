"""Env sharding and episode-metric reduction for one-process-per-GPU runs.

The rollout path has no exchange step: envs are independent, each rank owns a contiguous
range of global env ids (its Philox streams are keyed by the GLOBAL id, so env e behaves the
same on any number of GPUs).  The only collective is one all-reduce(sum) of a small metrics
vector at episode end; with backend "nccl" it runs over NVLink/NVSwitch, tests use
"gloo".  (The reference fans episodes out over Dask worker processes and gathers JSON,
eval_configs/01-random/01-random.yaml:145-149.)
"""
from __future__ import annotations

import numpy as np

METRIC_NAMES = ["episodes", "CSR", "ISR", "SoC", "makespan", "ep_length", "agent_steps", "agents", "avg_agents_density"]


def shard_range(n_envs_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [first, first+count) of global env ids for `rank` (strong scaling split)."""
    base, rem = divmod(n_envs_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def local_metric_sums(per_env: np.ndarray) -> np.ndarray:
    """engine.metrics() rows [ep_length, CSR, ISR, SoC, makespan, on_goal, agent_steps, n_agents, avg_agents_density, ...]
    -> the additive vector that is all-reduced (METRIC_NAMES order)."""
    m = np.asarray(per_env, dtype=np.float64)
    live = m[:, 7] > 0
    m = m[live]
    dens = m[:, 8].sum() if m.shape[1] > 8 else 0.0
    return np.array([m.shape[0], m[:, 1].sum(), m[:, 2].sum(), m[:, 3].sum(), m[:, 4].sum(), m[:, 0].sum(),
                     m[:, 6].sum(), m[:, 7].sum(), dens], dtype=np.float64)


def reduce_metrics(local_sums: np.ndarray, device=None) -> dict:
    """all-reduce(sum) across the default process group (no-op when not initialised)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.asarray(local_sums, dtype=np.float64).copy())
    if dist.is_available() and dist.is_initialized():
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t = t.cpu()
    s = t.numpy()
    n = max(s[0], 1.0)
    return {"episodes": int(s[0]), "CSR": s[1] / n, "ISR": s[2] / n, "SoC": s[3] / n, "makespan": s[4] / n,
            "ep_length": s[5] / n, "agent_steps": s[6], "agents": s[7], "avg_agents_density": s[8] / n}


def gather_rows(local_rows: np.ndarray, total: int, first: int, device=None) -> np.ndarray:
    """Per-episode result rows of every rank on every rank: rank r holds rows [first, first + len(local_rows)) of a
    [total, cols] table; one all-reduce(sum) of the zero-padded table (benchmark.py gathers its episode records with it,
    where the reference gathers JSON from Dask workers)."""
    import torch
    import torch.distributed as dist
    local_rows = np.asarray(local_rows, dtype=np.float64)
    full = np.zeros((total, local_rows.shape[1]), dtype=np.float64)
    full[first:first + local_rows.shape[0]] = local_rows
    if dist.is_available() and dist.is_initialized():
        t = torch.from_numpy(full)
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        full = t.cpu().numpy()
    return full
