"""cpu_rollout.py -- TEST/BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's CPU implementation of the hot path, timed by bench.py (`cpu_baseline` and
`--impl reference`) and used by tests as the end-to-end checker:

  tokens   : the reference's own ObservationGenerator compiled into oracle/_ref when it was
             built here (kind "reference"), else oracle/obs_oracle.c (kind "port")
  forward  : oracle/gpt_oracle.py -- the same ATen fp32 calls as mapf_gpt/model.py (the
             Python reference cannot travel to the GPU box; pinned by tests/golden)
  sampling : torch.multinomial with a CPU generator seeded 0 (inference.py:69-70, model.py:257)
  env step : oracle/pogema_oracle.c (pogema itself is not installable; PARITY UNPINNED)

Per step and per env this mirrors MAPFGPTInference.act (inference.py:148-172):
update_agents(last actions) -> generate_observations -> torch.tensor(long) -> act -> step.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

import oracle
from oracle import gpt_oracle


class CpuRollout:
    def __init__(self, grid, starts, goals, sd, n_layer, n_head, threads=None, use_ref=True, batch_size=2048):
        """starts/goals: [E, n, 2] int32; grid: [H, W] padded obstacles."""
        self.grid = np.ascontiguousarray(grid, dtype=np.int32)
        self.pos = np.array(starts, dtype=np.int32).copy()
        self.goals = np.array(goals, dtype=np.int32)
        self.E, self.n = self.pos.shape[:2]
        self.sd, self.n_layer, self.n_head = sd, n_layer, n_head
        self.batch_size = batch_size
        ref = oracle.load_ref_module() if use_ref else None
        self.kind = "reference" if ref is not None else "port"
        self.gens = []
        for e in range(self.E):
            if ref is not None:
                g = ref.ObservationGenerator(self.grid.tolist(), ref.InputParameters(20, 13, 5, 256, 5, 5, 64, False))
                g.create_agents([tuple(p) for p in self.pos[e].tolist()], [tuple(p) for p in self.goals[e].tolist()])
            else:
                g = oracle.ObsOracle(self.grid)
                g.create_agents(self.pos[e], self.goals[e])
            self.gens.append(g)
        # the reference generator pins OpenMP to one thread process-wide (observation_generator.h:115),
        # which also throttles torch; give the forward all host cores back (the "fair" baseline).
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.last = np.full((self.E, self.n), -1, dtype=np.int32)
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(0)
        self.is_ref = ref is not None

    def tokens(self) -> np.ndarray:
        out = np.empty((self.E, self.n, 256), dtype=np.int64)
        for e, g in enumerate(self.gens):
            if self.is_ref:
                g.update_agents([tuple(p) for p in self.pos[e].tolist()], [tuple(p) for p in self.goals[e].tolist()],
                                self.last[e].tolist())
                out[e] = np.asarray(g.generate_observations(), dtype=np.int64)
            else:
                g.update_agents(self.pos[e], self.goals[e], self.last[e])
                out[e] = g.generate_observations()
        return out

    def step(self, do_sample=True, q=None):
        """One timestep for all envs.  Returns (tokens, actions)."""
        toks = self.tokens()
        rows = torch.from_numpy(toks.reshape(-1, 256))
        acts = []
        for i in range(0, rows.shape[0], self.batch_size):
            qq = None if q is None else q[i:i + self.batch_size]
            acts.append(gpt_oracle.act(self.sd, self.n_layer, self.n_head, rows[i:i + self.batch_size],
                                       do_sample=do_sample, generator=self.gen, q=qq))
        acts = torch.cat(acts).numpy().astype(np.int32).reshape(self.E, self.n)
        self.last = acts
        for e in range(self.E):
            self.pos[e], _ = oracle.pogema_step_soft(self.grid, self.pos[e], acts[e])
        return toks, acts


def time_cpu_rollout(grid, starts, goals, sd, n_layer, n_head, steps, warmup=1, threads=None):
    """agent-steps/s of the CPU path on a bounded sample; returns (value, info)."""
    r = CpuRollout(grid, starts, goals, sd, n_layer, n_head, threads=threads)
    for _ in range(warmup):
        r.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        r.step()
    dt = time.perf_counter() - t0
    n = r.E * r.n * steps
    return n / dt, {"kind": r.kind, "cores": r.threads, "seconds": dt, "agent_steps": n,
                    "ms_per_step": 1e3 * dt / steps}
