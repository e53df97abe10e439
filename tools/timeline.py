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
eng.forward_tokens(toks)
out = np.zeros((4, 128), np.int64)
L.mg_test_timeline(eng._h, 0, out.ctypes.data_as(C.c_void_p))
names = {0: "mma:start", 1: "mma:att ready", 2: "mma:proj issued", 3: "mma:ln2 ready", 40: "mma:all issued",
         50: "wrk:proj done", 51: "wrk:epi1 pass1", 52: "wrk:ln2 arrive", 90: "wrk:done seen", 91: "wrk:end"}
names.update({100: "att:start", 101: "att:QK landed", 102: "att:P+V ready", 103: "att:PV issued", 110: "att:S seen",
              111: "att:max done", 112: "att:P arrive", 113: "att:O seen", 114: "att:end"})
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
