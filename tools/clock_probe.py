"""SM clock each of the two big kernels really runs at inside a long, power-capped rollout (2M, mazes 64 x 1024).

A mid-grid CTA of `post_attn_kernel` and CTA 100 of `attn_persistent_kernel` record clock64() (SM cycles) and %globaltimer (ns) at
their start and end (timeline slots 2112.. / 2120..; only with the test hook enabled).  cycles / ns = the SM clock during that CTA.
Run once cold (single forward after idle) and once at the end of a sustained rollout; NVML's sampled clock is printed beside it.
"""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import bench
from mapf_gpt_b200 import engine as E, weights as W, _lib

L = _lib.lib()
cfg = W.model_config("2M")
grid, st, gl = bench.build_instances("validation-mazes-seed-000", 64, 1024, 0)
eng = E.RolloutEngine(1024, 64, *grid.shape)
eng.load_model(W.random_init(cfg, 1234), cfg)
eng.set_seed(0)
eng.reset(0, grid, st, gl)
eng.synchronize()


def probe(tag, steps):
    L.mg_test_timeline(eng._h, 1, None)
    s = bench.ClockSampler(0)
    s.start()
    eng.rollout(steps, E.MODE_PHILOX)
    eng.synchronize()
    clocks = s.stop()
    tot, _ = eng.last_timing()
    out = np.zeros((17, 128), np.int64)
    L.mg_test_timeline(eng._h, 0, out.ctypes.data_as(C.c_void_p))
    f = out.reshape(-1)
    res = {"tag": tag, "steps": steps, "ms_per_step": round(tot / steps, 2), "nvml_sm_mhz_median": clocks.get("sm_mhz"),
           "nvml_power_w_max": clocks.get("power_w_max")}
    for name, o in (("post_attn", 2112), ("attention", 2120)):
        dc, dt = int(f[o + 2] - f[o]), int(f[o + 3] - f[o + 1])
        res[name] = {"cycles": dc, "ns": dt, "sm_mhz": round(dc / dt * 1e3, 1) if dt > 0 else None}
    print(res, flush=True)
    return res


eng.rollout(3, E.MODE_PHILOX)       # warm-up (allocations)
eng.synchronize()
probe("sustained: last launches of a 12-step rollout", 12)
r = probe("sustained again", 12)
# CUDA-event kernel averages of the same sustained state -> what a CTA slot spends per tile vs the CTA's own lifetime
eng.set_profiling(True)
eng.rollout(8, E.MODE_PHILOX)
eng.synchronize()
kt = eng.kernel_times()
eng.set_profiling(False)
post = kt["post_attn_fused"]["ms"] / kt["post_attn_fused"]["launches"]
att = kt["attention"]["ms"] / kt["attention"]["launches"]
mhz = r["post_attn"]["sm_mhz"]
slots = 2 * 148
tiles_per_slot = 2 * 8192 / slots
per_tile_slot = post * 1e-3 * mhz * 1e6 / tiles_per_slot
print({"post_attn_ms": round(post, 4), "attention_ms": round(att, 4), "cycles_per_tile_of_a_CTA_slot": round(per_tile_slot),
       "cycles_of_one_CTA (start stamp .. end stamp)": r["post_attn"]["cycles"],
       "gap_cycles_per_tile (CTA teardown + launch of the next cluster, plus the launch's ramp and tail)": round(per_tile_slot - r["post_attn"]["cycles"]),
       "attention_kernel_cycles_from_events": round(att * 1e-3 * r["attention"]["sm_mhz"] * 1e6), "attention_CTA_cycles": r["attention"]["cycles"]})
eng.close()
