"""Logit error and sampled-action flip rate of the bf16 tensor-core forward against the fp32 oracle at three logit scales
(weights x1, x3, x10 of the seeded random init with perturbed LayerNorm gains), per model size.  GPU only.

For every scale: tokens of a real rollout state (mazes, 64 agents x 32 envs = 2048 rows after 3 steps), engine logits vs
oracle/gpt_oracle.py in fp32 on cuda (TF32 off), actions = argmax(softmax(l)/q) with the same Exp(1) draws q for both.
Prints one JSON line per (model, scale): max |dlogit|, logit std (the scale the error has to be read against), flip rate."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mapf_gpt_b200 import engine as E, maps, weights as W   # noqa: E402
from oracle import gpt_oracle as G                          # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
m = maps.load_map("validation-mazes-seed-000")
n, envs = 64, 32
st = np.stack([maps.sample_instance(m, n, 0, e)[0] for e in range(envs)])
gl = np.stack([maps.sample_instance(m, n, 0, e)[1] for e in range(envs)])
for name in (sys.argv[1:] or ["2M", "6M", "85M"]):
    cfg = W.model_config(name)
    for scale in (1.0, 3.0, 10.0):
        sd = W.scale_weights(W.perturb_layernorm(W.random_init(cfg)), scale)
        eng = E.RolloutEngine(envs, n, *m["grid"].shape)
        eng.load_model(sd, cfg)
        eng.reset(0, m["grid"], st, gl)
        eng.rollout(3, E.MODE_PHILOX)
        eng.update_agents()
        toks = eng.generate_observations().reshape(-1, 256)
        q = np.random.default_rng(0).exponential(size=(envs, n, 5)).astype(np.float32)
        acts, lg = eng.act(E.MODE_SUPPLIED_Q, q, want_logits=True)
        eng.close()
        sdd = {k: v.cuda() for k, v in sd.items()}
        ref = torch.cat([G.forward_logits(sdd, cfg.n_layer, cfg.n_head, torch.from_numpy(toks[i:i + 256].astype(np.int64)).cuda())[:, :5]
                         for i in range(0, len(toks), 256)]).cpu().numpy()
        lg = lg.reshape(-1, 5)
        p = torch.softmax(torch.from_numpy(ref), -1).numpy() / q.reshape(-1, 5)
        ref_act = p.argmax(-1)
        greedy_flip = float((lg.argmax(-1) != ref.argmax(-1)).mean())
        print(json.dumps({"model": name, "weight_scale": scale, "rows": len(toks), "max_abs_logit_err": float(np.abs(lg - ref).max()),
                          "mean_abs_logit_err": float(np.abs(lg - ref).mean()), "logit_std": float(ref.std()),
                          "max_abs_logit": float(np.abs(ref).max()),
                          "sampled_action_flip_rate": float((acts.reshape(-1) != ref_act).mean()), "greedy_action_flip_rate": greedy_flip}),
              flush=True)
