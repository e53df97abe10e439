#!/usr/bin/env python
"""bench.py -- agent-steps/s of the MAPF-GPT rollout path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)

Workload (config.workload): BASELINE.json configs[1] = validation-mazes-seed-000, 64 agents x
1024 envs PER GPU (weak scaling; envs are independent, no data-path collective), MAPF-GPT-2M,
seeded random-init weights, synthetic starts/goals.  A "step" = one timestep of
all envs: update_agents -> tokenizer -> GPT forward -> sample -> POGEMA soft step.

  value  whole-job agent-steps/s, state resident in HBM, timed with CUDA events on the
         engine's stream, max over ranks
  e2e    same metric through the C ABI with HOST buffers (mg_engine_act_host + mg_engine_env_step):
         positions/goals H2D and actions/positions D2H inside the timed region
  roofline / cpu_baseline: see DESIGN.md section "Measurement"
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "agent-steps/sec on POGEMA mazes (64 agents x 1024 envs per GPU, MAPF-GPT-2M)"
UNIT = "agent-steps/s"


def flops_per_agent_step(L, C, T=256, V=67):
    """F_ref of SURVEY 8d: L*(24*T*C^2 + 4*T^2*C) + 2*C*V (multiply-add = 2 FLOP)."""
    return L * (24 * T * C * C + 4 * T * T * C) + 2 * C * V


def flops_executed(L, C, T=256, V=67, pruned=True, block0_table=False):
    """What the kernels execute: with last-block pruning the last block runs attention for one
    query and c_proj + MLP for one token per sequence (its QKV GEMM still covers all tokens); with block 0 tabulated
    per (token, position) at model load its c_attn GEMM (6 T C^2) is a lookup, not arithmetic."""
    f = flops_per_agent_step(L, C, T, V)
    if pruned:
        f -= (4 * T * T * C - 4 * T * C) + 18 * C * C * (T - 1)
    if block0_table:
        f -= 6 * T * C * C
    return f


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.load(open(p))
        return {"tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"], "hbm_gbs": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_ev = gpu_index, [], threading.Event()

    def run(self):
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_instances(map_name, n_agents, n_envs, first_env, seed=0):
    from mapf_gpt_b200 import maps
    m = maps.load_map(map_name)
    st = np.empty((n_envs, n_agents, 2), np.int32)
    gl = np.empty((n_envs, n_agents, 2), np.int32)
    for e in range(n_envs):
        st[e], gl[e] = maps.sample_instance(m, n_agents, seed, first_env + e)
    return m["grid"], st, gl


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    import torch
    from mapf_gpt_b200 import weights as W
    from oracle import cpu_rollout
    cfg = W.model_config(args.model)
    sd = W.random_init(cfg, 1234)
    envs = args.ref_envs
    grid, st, gl = build_instances(args.map, args.agents, envs, 0)
    cores = os.cpu_count() or 1
    r = cpu_rollout.CpuRollout(grid, st, gl, sd, cfg.n_layer, cfg.n_head, threads=cores)
    for _ in range(args.warmup):
        r.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.step()
    dt = time.perf_counter() - t0
    val = envs * args.agents * args.steps / dt
    sample = (f"{envs} env x {args.agents} agents x {args.steps} steps of the same workload "
              f"({r.kind} tokenizer, torch-fp32 forward on {cores} threads, C soft-step)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.envs),     # the workload; each step of this arm is a bounded sample of it
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": r.kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "torch_threads": torch.get_num_threads()}
    print(json.dumps(line), flush=True)


def workload_config(args, envs_per_gpu):
    return {"workload": f"{args.map}, {args.agents} agents x {envs_per_gpu} envs per GPU, MAPF-GPT-{args.model}, "
                        f"obs radius 5, soft collisions, sampling",
            "map": args.map, "agents": args.agents, "envs_per_gpu": envs_per_gpu, "policy": f"MAPF-GPT-{args.model}",
            "weights": "seeded random init (pretrained weights need network)",
            "cache": "inputs larger than L2 (activation working set of one step >> 126 MB)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="2M", choices=["2M", "6M", "85M"])
    ap.add_argument("--map", default="validation-mazes-seed-000")
    ap.add_argument("--agents", type=int, default=64)
    ap.add_argument("--envs", type=int, default=1024, help="envs per GPU")
    ap.add_argument("--ref-envs", type=int, default=8, help="envs in the CPU sample (--impl reference / cpu_baseline)")
    ap.add_argument("--cpu-steps", type=int, default=12, help="timesteps of the cpu_baseline sample (~10-20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--quick", action="store_true", help="profiling runs: no e2e / cpu baseline, warm-up as given")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.quick:
        args.no_e2e = args.no_cpu_baseline = True
    elif args.impl == "ours":
        args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mapf_gpt_b200 import engine as E, weights as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    cfg = W.model_config(args.model)
    sd = W.random_init(cfg, 1234)
    E_gpu, n = args.envs, args.agents
    first_env = rank * E_gpu
    grid, st, gl = build_instances(args.map, n, E_gpu, first_env)
    H, Wd = grid.shape
    eng = E.RolloutEngine(E_gpu, n, H, Wd, device=local_rank)
    eng.load_model(sd, cfg)
    eng.set_seed(0)
    eng.set_env_offset(first_env)
    eng.reset(0, grid, st, gl)
    eng.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.synchronize()

    # ---- device-resident rollout: warm-up, then exactly K timed steps
    eng.rollout(args.warmup, E.MODE_PHILOX)
    eng.synchronize()
    eng.set_profiling(True)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    eng.rollout(args.steps, E.MODE_PHILOX)
    eng.synchronize()
    clocks = sampler.stop()
    total_ms, phases = eng.last_timing()
    barrier()
    launches = eng.launch_count() - launches0
    ktimes = eng.kernel_times()
    eng.set_profiling(False)
    t = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    agent_steps = world * E_gpu * n * args.steps
    value = agent_steps / (total_ms_max * 1e-3)

    # ---- e2e through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        pos = eng.positions()
        acts = eng.act_host(pos, gl, E.MODE_PHILOX)          # warm
        pos = eng.env_step(None)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            acts = eng.act_host(pos, gl, E.MODE_PHILOX)      # H2D pos+goal, D2H actions
            pos = eng.env_step(None)                         # device step, D2H positions
        eng.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": agent_steps / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(pos.nbytes + gl.nbytes), "d2h_bytes_per_step": int(acts.nbytes + pos.nbytes),
               "api": "mg_engine_act_host + mg_engine_env_step (C ABI, host numpy buffers)"}

    # ---- episode metrics: the only cross-GPU exchange (one all-reduce of 8 doubles)
    met = eng.metrics()
    msum = torch.tensor(np.concatenate([[met.shape[0]], met[:, 1:7].sum(0), [met[:, 7].sum()]]), dtype=torch.float64,
                        device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(msum, op=dist.ReduceOp.SUM)

    if rank == 0:
        pk = peaks()
        F = flops_per_agent_step(cfg.n_layer, cfg.n_embd)
        pruned = (cfg.n_embd in (160, 256) and os.environ.get("MAPF_GPT_B200_GENERIC", "0")[:1] != "1"
                  and os.environ.get("MAPF_GPT_B200_NO_PRUNE") != "1")
        fused_path = cfg.n_embd in (160, 256) and os.environ.get("MAPF_GPT_B200_GENERIC", "0")[:1] != "1"
        fuse_qkv = fused_path and os.environ.get("MAPF_GPT_B200_NO_QKV_FUSION") is None
        table0 = fuse_qkv and os.environ.get("MAPF_GPT_B200_NO_BLOCK0_TABLE") is None
        Fx = flops_executed(cfg.n_layer, cfg.n_embd, pruned=pruned, block0_table=table0)
        C, L, T = cfg.n_embd, cfg.n_layer, 256
        rows_per_launch = min(E_gpu * n, 8192) * T           # the engine forwards in chunks of 8192 sequences
        kflops = {"gemm_qkv": 2 * 3 * C * C, "gemm_attn_proj": 2 * C * C, "gemm_fc_gelu": 2 * 4 * C * C,
                  "gemm_mlp_proj": 2 * 4 * C * C, "attention": 4 * T * C,
                  "post_attn_fused": 2 * 9 * C * C}   # per token
        if fuse_qkv:
            kflops["post_attn_fused"] += 2 * 3 * C * C       # + the next block's c_attn
        x_from_tab = table0 and cfg.n_layer >= 2 and os.environ.get("MAPF_GPT_B200_BLOCK0_X_VIA_HBM") is None
        kbytes = {"block0_lookup": (6 if x_from_tab else 10) * C, "embed": 6 * C, "attention_last_token": 4 * C}
        kern = {}
        if table0 and "embed" in ktimes:       # block 0 runs as the (token, position) lookup, not embedding + LN + GEMM
            ktimes = {("block0_lookup" if k == "embed" else k): v for k, v in ktimes.items()}
        for k, v in ktimes.items():
            if v["launches"]:
                avg = v["ms"] / v["launches"]
                ent = {"ms_total": round(v["ms"], 3), "launches": v["launches"], "avg_ms": round(avg, 4),
                       "share": round(v["ms"] / total_ms, 4)}
                if k in kflops:
                    ent["tflops"] = round(kflops[k] * rows_per_launch / (avg * 1e-3) / 1e12, 1)
                if k in kbytes:     # HBM-bound kernels: algorithmic bytes per token against the measured copy peak
                    ent["hbm_gbs"] = round(kbytes[k] * rows_per_launch / (avg * 1e-3) / 1e9, 1)
                    ent["hbm_frac"] = round(ent["hbm_gbs"] / pk["hbm_gbs"], 3)
                kern[k] = ent
        dom = max((k for k in kern if k in kflops), key=lambda k: kern[k]["ms_total"])
        achieved = kern[dom]["tflops"]
        traffic = None
        tfile = ROOT / "profiles" / "dram_traffic.json"       # dram__bytes_read+write per launch from the committed ncu capture
        if tfile.exists():
            traffic = json.load(open(tfile)).get(f"{args.model}:{dom}")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(args, E_gpu),
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": round(achieved / pk["tflops"], 4), "traffic": traffic, "peak_source": pk["source"],
                         "whole_step_tflops": round(value / world * Fx / 1e12, 1),
                         "whole_step_frac": round(value / world * Fx / 1e12 / pk["tflops"], 4),
                         "flops_per_agent_step_executed": Fx, "flops_per_agent_step_reference": F,
                         "last_block_pruned": pruned, "block0_lookup_table": table0},
            "kernels": kern, "phases_ms_last_step": {"observe": phases[0], "forward": phases[1], "sample_step": phases[2]},
            "clocks": clocks,
            "episode_metrics_sum": {"envs": msum[0].item(), "CSR": msum[1].item(), "ISR": msum[2].item(),
                                    "SoC": msum[3].item(), "makespan": msum[4].item(), "agent_steps": msum[6].item()},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_rollout
            gridc, stc, glc = build_instances(args.map, n, args.ref_envs, 0)
            v, info = cpu_rollout.time_cpu_rollout(gridc, stc, glc, sd, cfg.n_layer, cfg.n_head, steps=args.cpu_steps)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                                    "sample": f"{args.ref_envs} env x {n} agents x {args.cpu_steps} steps of the same "
                                              f"workload, {info['seconds']:.1f} s ({info['kind']} tokenizer + torch-fp32 "
                                              f"forward + C soft-step)"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
