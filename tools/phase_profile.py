"""Average cycles per phase of post_attn_kernel over ALL CTAs of one forward (8192 sequences, MAPF-GPT-2M).

Needs a library built with the phase accumulators compiled in:
    make -C mapf_gpt_b200/csrc OUT=../libprof.so EXTRA=-DMG_PHASE_PROF
    MAPF_GPT_B200_LIB_PATH=$PWD/mapf_gpt_b200/libprof.so python tools/phase_profile.py
(the single-CTA view is tools/timeline.py).  Worker numbers are thread 0's laps, issuer numbers lane 0's."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, ".")
from mapf_gpt_b200 import engine as E, weights as W, _lib
cfg = W.model_config("2M")
eng = E.RolloutEngine(1, 1, 11, 11)
eng.load_model(W.random_init(cfg), cfg)
n_seq = 8192
toks = np.random.default_rng(0).integers(0, 67, (n_seq, 256)).astype(np.int8)
L = _lib.lib()
for _ in range(3):
    eng.forward_tokens(toks)                  # warm
L.mg_test_timeline(eng._h, 1, None)
eng.set_profiling(True)
reps = 4
for _ in range(reps):
    eng.forward_tokens(toks)
eng.synchronize()
kt = eng.kernel_times()
print("avg kernel ms:", {k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in kt.items() if v["launches"]})
eng.set_profiling(False)
out = np.zeros((17, 128), np.int64)
L.mg_test_timeline(eng._h, 0, out.ctypes.data_as(C.c_void_p))
acc = out.reshape(-1)[2048:2048 + 64]
tiles = reps * ((cfg.n_layer - 1) * n_seq * 2 + n_seq // 128)
names = {0: "wrk residual tile -> TMEM", 1: "wrk wait att tile + c_proj", 2: "wrk LN2", 3: "wrk wait FC chunk (x8)",
         4: "wrk FC chunk -> registers (x8)", 5: "wrk GELU (x8)", 6: "wrk wait hidden buffer (x8)", 7: "wrk hidden -> smem (x8)",
         8: "wrk wait last proj2", 9: "wrk x' -> HBM + stats", 10: "wrk LN1_next -> smem", 11: "wrk wait qkv half-tile (x6)",
         12: "wrk qkv half-tile -> HBM (x6)", 13: "wrk tail", 20: "mma wait weight ring (all stages)",
         21: "mma issue + wait workers", 22: "mma wait att tile + residual"}
tw = sum(int(acc[k]) for k in range(0, 16)) / tiles
tm = sum(int(acc[k]) for k in range(20, 26)) / tiles
for k, n in names.items():
    v = int(acc[k]) / tiles
    print(f"{v:9.0f} cycles/tile  {100 * v / (tw if k < 16 else tm):5.1f} %  {n}")
print(f"{tw:9.0f} cycles/tile  worker total (after the TMEM rendezvous)")
print(f"{tm:9.0f} cycles/tile  issuer total")
