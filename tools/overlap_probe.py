"""Would running two forward chunks concurrently (attention of one next to post-attention of the other on the same SMs) beat
running them back to back?  Two engines, two host threads, two streams: aggregate sequences/s vs one engine alone."""
import sys, time, threading
import numpy as np
sys.path.insert(0, ".")
from mapf_gpt_b200 import engine as E, weights as W
cfg = W.model_config("2M")
sd = W.random_init(cfg)
n_seq, iters = 8192, 12
toks = np.random.default_rng(0).integers(0, 67, (n_seq, 256)).astype(np.int8)
engs = []
for _ in range(2):
    e = E.RolloutEngine(1, 1, 11, 11)
    e.load_model(sd, cfg)
    e.forward_tokens(toks)
    engs.append(e)

def run(e, n):
    for _ in range(n):
        e.forward_tokens(toks)

t0 = time.perf_counter(); run(engs[0], iters); t1 = time.perf_counter()
solo = iters * n_seq / (t1 - t0)
th = [threading.Thread(target=run, args=(e, iters)) for e in engs]
t0 = time.perf_counter()
for t in th: t.start()
for t in th: t.join()
t1 = time.perf_counter()
duo = 2 * iters * n_seq / (t1 - t0)
print(f"one engine: {solo:,.0f} seq/s   two concurrent engines: {duo:,.0f} seq/s   ratio {duo / solo:.3f}")
