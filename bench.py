#!/usr/bin/env python
"""bench.py -- agent-steps/s of the MAPF-GPT rollout path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)

Headline workload (config.workload): BASELINE.json configs[1] = validation-mazes-seed-000, 64 agents x 1024 envs PER GPU
(weak scaling; envs are independent, no data-path collective), MAPF-GPT-2M, seeded random-init weights, synthetic
starts/goals.  A "step" = one timestep of all envs: update_agents -> tokenizer -> GPT forward -> sample ->
POGEMA soft step.

  value          whole-job agent-steps/s (steps EXECUTED by non-finished envs, from the engine's own counters), state
                 resident in HBM, timed with CUDA events on the engine's stream, max over ranks
  e2e            same metric through the C ABI with HOST buffers (mg_engine_act_host + mg_engine_env_step):
                 positions/goals H2D and actions/positions D2H inside the timed region
  roofline       dominant tensor-core kernel and the whole step against MEASURED_PEAKS.json (DESIGN.md "Measurement")
  other_configs  the other BASELINE.json configs, a few steps each, same measurements: C3 (wfi_warehouse 192 x 512, 6M),
                 C4 per-GPU shard (Berlin tile 256 x 32, 85M; at --gpus 8 this is C4), mazes at the literal metric shape
                 (256 agents x 256 envs), two C5 points (8 and 512 agents on a Berlin tile, 85M) and C1 as the latency of
                 MAPFGPTInference.act() (32 agents x 1 env, obs dicts in, list out, median of 60 calls)
  cpu_baseline   the UNMODIFIED reference (mapf_gpt/inference.py + model.py + compiled generator from baseline/_ref,
                 oracle/ref_runtime.py) on the host cores in the form that uses them best: its own scaling model, one single-thread
                 worker process per core, one env each, one act() per step (= what --impl reference times);
  cpu_act_batch_threads: one process, act_batch over 8 envs, torch on all threads;
  cpu_as_shipped one process, one act() per env, OpenMP pinned to one thread by the generator (as shipped)
  stock_gpu      the same reference object with device='cuda' (inference.py:58-60): stock PyTorch fp32 kernels and a
                 bf16-autocast variant, tokenizer on the host as shipped -- the "stock PyTorch on the same B200" bar
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
GENERIC = os.environ.get("MAPF_GPT_B200_GENERIC", "0")[:1] == "1"


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
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md): NVML when nvidia_ml_py is importable (a sample
    costs microseconds, so short timed regions still get many), else the nvidia-smi query of the recipe."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_ev = gpu_index, [], threading.Event()
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:                                    # NVML counts physical devices
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if gpu_index < len(ids) and ids[gpu_index].strip().isdigit():
                    idx = int(ids[gpu_index])
            self._nv, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
        except Exception:
            self._h = None

    def run(self):
        while not self._stop_ev.is_set():
            if self._h is not None:
                try:
                    nv = self._nv
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, name in self.NVML_REASONS.items():
                        if bits & bit:
                            self.reasons.add(name)
                except Exception:
                    self._h = None
                    continue
                self._stop_ev.wait(0.02)
                continue
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    r = [c.strip() for c in line.split(",")]
                    if len(r) >= 8 and r[1].replace(".", "").isdigit():
                        self.sm.append(float(r[1]))
                        self.mx.append(float(r[2]))
                        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                        self.reasons.update(names[i] for i in range(4) if r[4 + i].lower().startswith("active"))
            except Exception:
                pass
            self._stop_ev.wait(0.15)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": round(max(self.power), 1) if self.power else None,
                "source": "nvml" if self.power else "nvidia-smi"}


def build_instances(map_name, n_agents, n_envs, first_env, seed=0):
    from mapf_gpt_b200 import maps
    m = maps.load_map(map_name)
    st = np.empty((n_envs, n_agents, 2), np.int32)
    gl = np.empty((n_envs, n_agents, 2), np.int32)
    for e in range(n_envs):
        st[e], gl[e] = maps.sample_instance(m, n_agents, seed, first_env + e)
    return m["grid"], st, gl


def workload_config(model, map_name, agents, envs_per_gpu):
    return {"workload": f"{map_name}, {agents} agents x {envs_per_gpu} envs per GPU, MAPF-GPT-{model}, "
                        f"obs radius 5, soft collisions, sampling",
            "map": map_name, "agents": agents, "envs_per_gpu": envs_per_gpu, "policy": f"MAPF-GPT-{model}",
            "weights": "seeded random init (pretrained weights need network)",
            "cache": "inputs larger than L2 (activation working set of one step >> 126 MB)"}


# --------------------------------------------------------------------------------------------- baselines (CPU / stock GPU)
def make_reference_rollout(model, map_name, agents, envs, device="cpu", mode="act_batch", threads=None, autocast=False):
    """The unmodified reference when baseline/_ref exists (kind "reference"), else the oracle port (kind "port")."""
    from mapf_gpt_b200 import weights as W
    from oracle import ref_runtime
    cfg = W.model_config(model)
    sd = W.random_init(cfg, 1234)
    grid, st, gl = build_instances(map_name, agents, envs, 0)
    if ref_runtime.available():
        return ref_runtime.ReferenceRollout(grid, st, gl, sd, cfg, device=device, mode=mode, torch_threads=threads,
                                            autocast_bf16=autocast), "reference"
    if device != "cpu":
        return None, "unavailable"
    from oracle import cpu_rollout
    return cpu_rollout.CpuRollout(grid, st, gl, sd, cfg.n_layer, cfg.n_head, threads=threads), "port"


def time_steps(r, steps, warmup, cuda=False):
    import torch
    for _ in range(warmup):
        r.step()
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        r.step()
    if cuda:
        torch.cuda.synchronize()
    return time.perf_counter() - t0


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores (rank 0 only), in the form
    that uses the cores best: its own scaling model, one single-thread worker process per core, each stepping one env through
    act() (eval_configs: num_process workers); the oracle port with torch threads when baseline/_ref is missing."""
    if rank != 0:
        return
    import torch
    from oracle import ref_runtime
    cores = os.cpu_count() or 1
    if ref_runtime.available():
        val, secs = ref_runtime.time_fair_processes(args.model, args.map, args.agents, cores, args.steps, timeout_s=600, warmup=max(args.warmup, 1))
        kind, envs, dt = "reference", cores, secs
        what = (f"unmodified mapf_gpt/inference.py + model.py + compiled observation generator (baseline/_ref): {cores} worker processes "
                f"(the reference's num_process model), one env and one torch thread each, one act() per step")
    else:
        envs = args.ref_envs
        r, kind = make_reference_rollout(args.model, args.map, args.agents, envs, "cpu", "act_batch", cores)
        dt = time_steps(r, args.steps, args.warmup)
        val = envs * args.agents * args.steps / dt
        what = f"oracle port: compiled/C tokenizer + torch-fp32 forward on {cores} threads"
    sample = (f"{envs} envs x {args.agents} agents x {args.steps} steps of the same workload; {what}; env = C soft-step (pogema not installable)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.model, args.map, args.agents, args.envs),   # each step of this arm is a bounded sample of it
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "device": "cpu (the GPU engine is compared with the reference's CPU path; see stock_gpu in the ours arm for cuda)",
            "torch_threads": torch.get_num_threads()}
    print(json.dumps(line), flush=True)


def comparators(args):
    """cpu_baseline (fair), cpu_as_shipped and stock_gpu on bounded samples of the headline workload (rank 0, N = 1)."""
    import torch
    out = {}
    cores = os.cpu_count() or 1
    envs = args.ref_envs
    r, kind = make_reference_rollout(args.model, args.map, args.agents, envs, "cpu", "act_batch", cores)
    dt = time_steps(r, args.cpu_steps, 1)
    threads = {"value": envs * args.agents * args.cpu_steps / dt, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{envs} envs x {args.agents} agents x {args.cpu_steps} steps of the same workload, {dt:.1f} s; "
                         f"{'unmodified reference MAPFGPTInference.act_batch (baseline/_ref)' if kind == 'reference' else 'oracle port'}"
                         f", ONE process, torch fp32 on {cores} threads (re-enabled after the generator pins OpenMP to 1), C soft-step env"}
    out["cpu_baseline"] = threads
    if kind == "reference":
        out["cpu_act_batch_threads"] = threads
        try:    # the reference's own scaling model: one single-thread worker process per core, one env each (BASELINE.md 4.5 "ref-fair")
            from oracle import ref_runtime
            v, secs = ref_runtime.time_fair_processes(args.model, args.map, args.agents, cores, 5)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                                   "sample": f"{cores} worker processes x 1 env x {args.agents} agents x 5 steps of the same workload, {secs:.1f} s; "
                                             f"unmodified reference (baseline/_ref), one act() per step and one torch thread per process (its "
                                             f"own num_process scaling model: the fastest way it uses the host cores), C soft-step env"}
        except Exception as ex:
            out["cpu_fair_processes_error"] = f"{type(ex).__name__}: {ex}"
        r, _ = make_reference_rollout(args.model, args.map, args.agents, 1, "cpu", "act", None)
        torch.set_num_threads(1)
        dt = time_steps(r, 3, 1)
        out["cpu_as_shipped"] = {"value": args.agents * 3 / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                                 "sample": f"1 env x {args.agents} agents x 3 steps, {dt:.1f} s; one act() per env per step, OpenMP "
                                           f"pinned to 1 thread by ObservationGenerator (observation_generator.h:115), as shipped"}
        torch.set_num_threads(cores)
        sg = {"unit": UNIT, "kind": "reference", "device": "cuda:0",
              "what": "unmodified reference MAPFGPTInference(device='cuda').act_batch: host tokenizer (1 thread) + stock PyTorch "
                      "forward (SDPA) + torch.multinomial, 2048-row chunks (inference.py:87-101)"}
        genvs = max(1, 2048 // args.agents)
        for name, ac in (("fp32", False), ("bf16_autocast", True)):
            r, _ = make_reference_rollout(args.model, args.map, args.agents, genvs, "cuda", "act_batch", cores, autocast=ac)
            dt = time_steps(r, 6, 2, cuda=True)
            sg[name] = genvs * args.agents * 6 / dt
            # forward only: the reference network on resident int64 tokens (no tokenizer, no list conversion)
            net = r.algo.net
            idx = torch.randint(0, 67, (2048, 256), device="cuda")
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if ac else torch.autocast("cuda", enabled=False)
            with ctx:
                for _ in range(2):
                    net.act(idx, generator=r.algo.torch_generator)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(5):
                    net.act(idx, generator=r.algo.torch_generator)
                torch.cuda.synchronize()
            sg[name + "_forward_only"] = 2048 * 5 / (time.perf_counter() - t0)
            del r
        sg["sample"] = f"{genvs} envs x {args.agents} agents x 6 steps (one 2048-row chunk per step); forward_only = GPT.act on 2048 resident rows"
        sg["tf32"] = bool(torch.backends.cuda.matmul.allow_tf32)
        out["stock_gpu"] = sg
    return out


# --------------------------------------------------------------------------------------------- one measured configuration
def kernel_table(model_cfg, ktimes, total_ms, seqs_per_gpu, pk):
    C, T = model_cfg.n_embd, 256
    fused_path = C in (160, 256) and not GENERIC
    pruned = os.environ.get("MAPF_GPT_B200_NO_PRUNE") != "1"     # last-block pruning: fused and (since round 2) generic path
    fuse_qkv = fused_path and os.environ.get("MAPF_GPT_B200_NO_QKV_FUSION") is None
    table0 = fuse_qkv and os.environ.get("MAPF_GPT_B200_NO_BLOCK0_TABLE") is None
    chunk = int(os.environ.get("MAPF_GPT_B200_CHUNK_SEQS", "8192")) // 128 * 128
    rows_per_launch = min(seqs_per_gpu, max(chunk, 128)) * T  # the engine forwards in chunks of 8192 sequences (env override)
    kflops = {"gemm_qkv": 2 * 3 * C * C, "gemm_attn_proj": 2 * C * C, "gemm_fc_gelu": 2 * 4 * C * C,
              "gemm_mlp_proj": 2 * 4 * C * C, "attention": 4 * T * C, "post_attn_fused": 2 * 9 * C * C}   # per token
    if fuse_qkv:
        kflops["post_attn_fused"] += 2 * 3 * C * C           # + the next block's c_attn
    x_from_tab = table0 and model_cfg.n_layer >= 2 and os.environ.get("MAPF_GPT_B200_BLOCK0_X_VIA_HBM") is None
    kbytes = {"block0_lookup": (6 if x_from_tab else 10) * C, "embed": 6 * C, "attention_last_token": 4 * C, "layernorm": 6 * C}
    if table0 and "embed" in ktimes:       # block 0 runs as the (token, position) lookup, not embedding + LN + GEMM
        ktimes = {("block0_lookup" if k == "embed" else k): v for k, v in ktimes.items()}
    kern = {}
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
    return kern, kflops, pruned, table0


def run_config(model, map_name, agents, envs, steps, warmup, rank, world, local_rank, do_e2e=True):
    """One BASELINE config on this rank's GPU: device-resident rollout (value), C-ABI host-buffer loop (e2e), kernel table."""
    import torch
    import torch.distributed as dist
    from mapf_gpt_b200 import engine as E, parallel, weights as W
    dev = f"cuda:{local_rank}"
    cfg = W.model_config(model)
    first_env = rank * envs
    grid, st, gl = build_instances(map_name, agents, envs, first_env)
    H, Wd = grid.shape
    eng = E.RolloutEngine(envs, agents, H, Wd, device=local_rank)
    eng.load_model(W.random_init(cfg, 1234), cfg)
    eng.set_seed(0)
    eng.set_env_offset(first_env)
    eng.reset(0, grid, st, gl)
    eng.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.synchronize()

    def executed():
        return float(eng.metrics()[:, 6].sum())

    eng.rollout(warmup, E.MODE_PHILOX)
    eng.synchronize()
    a0 = executed()
    # Stream lanes (2M fused path): attention of one chunk and the post-attention kernel of the other share the SMs, so a
    # per-kernel CUDA-event duration inside the timed region would measure the overlap, not the kernel.  The timed region then
    # runs without per-kernel events, and the kernel table comes from a second, single-lane pass of the same K steps.
    overlapped = eng.num_lanes() > 1
    eng.set_profiling(not overlapped)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    eng.rollout(steps, E.MODE_PHILOX)
    eng.synchronize()
    clocks = sampler.stop()
    total_ms, phases = eng.last_timing()
    barrier()
    launches = eng.launch_count() - launches0 - 1            # minus the metrics kernel of executed()
    a1 = executed()
    ktotal_ms = total_ms
    if overlapped:
        eng.set_profiling(True)
        eng.rollout(steps, E.MODE_PHILOX)
        eng.synchronize()
        ktotal_ms, phases = eng.last_timing()
    ktimes = eng.kernel_times()
    eng.set_profiling(False)
    t = torch.tensor([total_ms, -total_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([a1 - a0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total_ms_max = float(t[0].item())
    agent_steps = float(cnt.item())
    value = agent_steps / (total_ms_max * 1e-3)

    e2e = None
    if do_e2e:
        pos = eng.positions()
        acts = eng.act_host(pos, gl, E.MODE_PHILOX)          # warm
        pos = eng.env_step(None)
        b0 = executed()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            acts = eng.act_host(pos, gl, E.MODE_PHILOX)      # H2D pos+goal, D2H actions
            pos = eng.env_step(None)                         # device step, D2H positions
        eng.synchronize()
        dt = time.perf_counter() - t0
        b1 = executed()
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        c2 = torch.tensor([b1 - b0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(c2, op=dist.ReduceOp.SUM)
        e2e = {"value": float(c2.item()) / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(pos.nbytes + gl.nbytes), "d2h_bytes_per_step": int(acts.nbytes + pos.nbytes),
               "api": "mg_engine_act_host + mg_engine_env_step (C ABI, host numpy buffers)"}

    # episode metrics: the only cross-GPU exchange (one all-reduce of a 9-double vector, mapf_gpt_b200/parallel.py)
    red = parallel.reduce_metrics(parallel.local_metric_sums(eng.metrics()), device=dev if world > 1 else None)
    pk = peaks()
    kern, kflops, pruned, table0 = kernel_table(cfg, ktimes, ktotal_ms, envs * agents, pk)
    F = flops_per_agent_step(cfg.n_layer, cfg.n_embd)
    Fx = flops_executed(cfg.n_layer, cfg.n_embd, pruned=pruned, block0_table=table0)
    dom = max((k for k in kern if k in kflops), key=lambda k: kern[k]["ms_total"])
    achieved = kern[dom]["tflops"]
    traffic = None
    tfile = ROOT / "profiles" / "dram_traffic.json"           # dram__bytes_read+write per launch from the committed ncu capture
    if tfile.exists() and envs * agents >= 8192:
        traffic = json.load(open(tfile)).get(f"{model}:{dom}")
    res = {
        "value": value, "unit": UNIT, "ms_per_step": total_ms_max / steps, "steps": steps, "warmup": warmup,
        "agent_steps_executed": agent_steps, "agent_steps_nominal": world * envs * agents * steps,
        "config": workload_config(model, map_name, agents, envs), "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": round(achieved / pk["tflops"], 4), "traffic": traffic, "peak_source": pk["source"],
                     "whole_step_tflops": round(value / world * Fx / 1e12, 1),
                     "whole_step_frac": round(value / world * Fx / 1e12 / pk["tflops"], 4),
                     "flops_per_agent_step_executed": Fx, "flops_per_agent_step_reference": F,
                     "last_block_pruned": pruned, "block0_lookup_table": table0},
        "kernels": kern, "phases_ms_last_step": {"observe": phases[0], "forward": phases[1], "sample_step": phases[2]},
        "clocks": clocks, "episode_metrics_mean": red,
    }
    if overlapped:
        res["stream_lanes"] = {"lanes": eng.num_lanes(), "timed_region_ms_per_step": total_ms_max / steps,
                               "single_lane_ms_per_step_with_kernel_events": ktotal_ms / steps,
                               "note": "value / ms_per_step: two stream lanes, attention of one chunk overlapping the post-attention "
                                       "kernel of the other on the same SMs; `kernels` and roofline.achieved: CUDA-event durations from a "
                                       "second single-lane pass of the same steps (kernels serialized), taken right after the timed region"}
    eng.close()
    return res


def c1_act_latency(local_rank, calls=60):
    """Config C1 through the drop-in: MAPFGPTInference.act(obs dicts) for random-000, 32 agents, 1 env, 2M (what
    example.py:65 and every Dask worker of the reference's benchmark do); env = the engine's own soft step on a second engine."""
    from mapf_gpt_b200 import engine as E, weights as W
    from mapf_gpt_b200.inference import MAPFGPTInference, MAPFGPTInferenceConfig
    cfg = W.model_config("2M")
    sd = W.random_init(cfg, 1234)
    grid, st, gl = build_instances("validation-random-seed-000", 32, 1, 0)
    out = {"workload": "validation-random-seed-000, 32 agents x 1 env, MAPF-GPT-2M, MAPFGPTInference.act(obs dicts) -> list",
           "calls": calls}
    for sampling in ("torch", "philox"):
        algo = MAPFGPTInference(MAPFGPTInferenceConfig(device=f"cuda:{local_rank}"), net=(sd, cfg), sampling=sampling)
        algo.reset_states()
        env = E.RolloutEngine(1, 32, *grid.shape, device=local_rank)
        env.reset(0, grid, st, gl)
        pos = st[0].copy()
        goals_t = [tuple(int(v) for v in g) for g in gl[0]]
        lat = []
        for i in range(calls + 5):
            obs = [{"global_obstacles": grid, "global_xy": (int(pos[k, 0]), int(pos[k, 1])), "global_target_xy": goals_t[k]}
                   for k in range(32)]
            t0 = time.perf_counter()
            acts = algo.act(obs)
            lat.append(time.perf_counter() - t0)
            pos = env.env_step(np.asarray(acts, np.int32)[None])[0]
        lat = np.array(lat[5:]) * 1e3
        tot, _ = algo._engine.last_timing()
        out[f"ms_per_act_{sampling}_sampling"] = {"median": float(np.median(lat)), "p10": float(np.percentile(lat, 10)),
                                                 "p90": float(np.percentile(lat, 90)), "device_ms_last_call": float(tot)}
        out[f"agent_steps_per_s_{sampling}_sampling"] = 32e3 / float(np.median(lat))
        env.close()
        algo.reset_states()
        algo._engine.close()
    return out


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
    ap.add_argument("--cpu-steps", type=int, default=10, help="timesteps of the cpu_baseline sample (~10-20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--quick", action="store_true", help="profiling runs: headline config only, no e2e / baselines, warm-up as given")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.quick:
        args.no_e2e = args.no_cpu_baseline = args.no_other_configs = True
    elif args.impl == "ours":
        args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"), timeout=datetime.timedelta(seconds=240))

    head = run_config(args.model, args.map, args.agents, args.envs, args.steps, args.warmup, rank, world, local_rank,
                      do_e2e=not args.no_e2e)
    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic"}
    line.update({k: head[k] for k in ("config", "e2e", "gpu_launches", "roofline", "kernels", "phases_ms_last_step", "clocks",
                                      "episode_metrics_mean", "agent_steps_executed", "agent_steps_nominal")})

    if not args.no_other_configs:
        others = {}
        plan = [("C3_warehouse_192x512_6M", "6M", "wfi_warehouse", 192, 512, 4),
                ("C4_shard_berlin_256x32_85M", "85M", "Berlin_1_256_05", 256, 32, 3),
                ("C5_berlin_8x32_85M", "85M", "Berlin_1_256_05", 8, 32, 3),
                ("C5_berlin_512x16_85M", "85M", "Berlin_1_256_05", 512, 16, 3),
                ("mazes_256x256_2M", "2M", "validation-mazes-seed-000", 256, 256, 6)]
        for name, model, mp, n, envs, steps in plan:
            try:
                r = run_config(model, mp, n, envs, steps, 3, rank, world, local_rank, do_e2e=True)
                keep = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "e2e", "gpu_launches", "clocks", "config")}
                keep["roofline"] = {k: r["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "whole_step_tflops", "whole_step_frac")}
                keep["kernel_shares"] = {k: v["share"] for k, v in r["kernels"].items() if v["share"] >= 0.02}
                others[name] = keep
            except Exception as ex:                            # a config that does not fit is reported, not hidden
                others[name] = {"error": f"{type(ex).__name__}: {ex}"}
        if rank == 0:
            try:
                others["C1_act_latency_random_32x1_2M"] = c1_act_latency(local_rank)
            except Exception as ex:
                others["C1_act_latency_random_32x1_2M"] = {"error": f"{type(ex).__name__}: {ex}"}
        if world > 1:
            dist.barrier()
        line["other_configs"] = others

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line.update(comparators(args))
            line["comparison_note"] = ("e2e / value are the GPU engine; cpu_baseline, cpu_as_shipped (reference on host cores) and "
                                       "stock_gpu (reference on this B200 with stock PyTorch kernels) are baselines, each on a "
                                       "bounded sample of the same workload")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
