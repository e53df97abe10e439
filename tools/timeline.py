"""Print the clock64() timeline of the fused post-attention kernel (first CTAs of one launch)."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, ".")
from mapf_gpt_b200 import engine as E, weights as W, _lib
cfg = W.model_config("2M")
eng = E.RolloutEngine(1, 1, 11, 11)
eng.load_model(W.random_init(cfg), cfg)
toks = np.random.default_rng(0).integers(0, 67, (8192, 256)).astype(np.int8)
L = _lib.lib()
eng.forward_tokens(toks)                      # warm
L.mg_test_timeline(eng._h, 1, None)
eng.set_profiling(True)
eng.forward_tokens(toks)
eng.synchronize()
print("kernel times of this single forward (cold clocks, no sustained power cap):",
      {k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in eng.kernel_times().items() if v["launches"]})
eng.set_profiling(False)
out = np.zeros((17, 128), np.int64)
L.mg_test_timeline(eng._h, 0, out.ctypes.data_as(C.c_void_p))
names = {0: "mma:start", 1: "mma:att ready", 2: "mma:proj issued", 40: "mma:all issued",
         50: "wrk:proj done", 51: "wrk:epi1 pass1", 52: "wrk:ln2 arrive", 90: "wrk:done seen", 91: "wrk:end"}
names.update({100: "att:start", 101: "att:QK landed", 102: "att:P+V ready", 103: "att:PV issued", 110: "att:S seen",
              111: "att:max done", 112: "att:P arrive", 113: "att:O seen", 114: "att:end"})
if "--classic-attn" not in sys.argv:   # persistent attention kernel: item 40 of the CTA, both query tiles
    for k in range(100, 128):
        names.pop(k, None)
    for qt in range(2):
        names.update({100 + 3 * qt: f"att{qt}:mma S issue", 101 + 3 * qt: f"att{qt}:mma P ready", 102 + 3 * qt: f"att{qt}:mma PV issued",
                      110 + 8 * qt: f"att{qt}:wrk S seen", 111 + 8 * qt: f"att{qt}:wrk max done", 112 + 8 * qt: f"att{qt}:wrk max exchanged",
                      113 + 8 * qt: f"att{qt}:wrk P arrive", 114 + 8 * qt: f"att{qt}:wrk O seen", 115 + 8 * qt: f"att{qt}:wrk end"})
names.update({92: "wrk:x' stored", 93: "wrk:qa arrive (LN1_next in smem)"})
names.update({94 + hh: f"wrk:qkv half-tile {hh} seen" for hh in range(6)})
names.update({126: "att:CTA start", 127: "att:CTA end"})
for j in range(8):
    names[10 + 2 * j] = f"mma:FC({j}) issued"; names[11 + 2 * j] = f"mma:P2({j}) issued"
    names[60 + 3 * j] = f"wrk:a1f({j}) seen"; names[61 + 3 * j] = f"wrk:a1e({j}) arrive"; names[62 + 3 * j] = f"wrk:hf({j}) arrive"
for cta in range(2):
    for grp, lo, hi in (("post_attn", 0, 100), ("attention", 100, 128)):
        ev = [(int(out[cta, k]), names[k]) for k in names if out[cta, k] and lo <= k < hi]
        ev.sort()
        if not ev:
            continue
        t0 = ev[0][0]
        print(f"--- CTA {cta} {grp}")
        prev = t0
        for t, n in ev:
            print(f"{t - t0:8d}  (+{t - prev:6d})  {n}")
            prev = t
    continue
    ev = []
    t0 = ev[0][0]
    print(f"--- CTA {cta}")
    prev = t0
    for t, n in ev:
        print(f"{t - t0:8d}  (+{t - prev:6d})  {n}")
        prev = t

flat = out.reshape(-1)
n_cta = int((flat[1536:2048] != 0).sum())
gt = flat[1536:1536 + n_cta]
if n_cta:
    aux = flat[1024:1024 + n_cta]     # narrow kernel: SM id; wide kernel: items processed by the CTA
    g = gt - gt.min()
    st = flat[512:512 + n_cta] - gt.min()
    print(f"persistent attention: {n_cta} CTAs; end times (ns after the first to finish): median {int(np.median(g))} max {int(g.max())}; "
          f"durations ns: min {int((g - st).min())} median {int(np.median(g - st))} max {int((g - st).max())}")
    print("per-CTA aux (items processed / SM id): min %d median %d max %d; CTA0 %d" % (aux.min(), np.median(aux), aux.max(), aux[0]))
    print("histogram of aux:", np.histogram(aux, 8))
