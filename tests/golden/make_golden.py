"""Generate the committed golden vectors by running the REFERENCE itself in the build container.

  * tokens : the reference ObservationGenerator compiled from /root/reference into oracle/_ref
             (oracle/Makefile), driven through update_agents/generate_observations
  * logits / sampled actions : /root/reference/mapf_gpt/model.py (GPT.forward / GPT.act, CPU fp32)
             loaded with the seeded weights of mapf_gpt_b200.weights.random_init

Run:  python tests/golden/make_golden.py      (needs /root/reference; writes tests/golden/*.npz)
The GPU box has no /root/reference: tests only read the .npz files.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

import oracle  # noqa: E402
from mapf_gpt_b200 import maps, weights as W  # noqa: E402

OUT = Path(__file__).resolve().parent
SCENARIOS = [  # (map, agents, steps, seed, solid padding)
    ("validation-random-seed-000", 32, 6, 1, True),
    ("validation-mazes-seed-000", 64, 5, 2, False),
    ("wfi_warehouse", 192, 2, 3, True),
    ("Berlin_1_256_03", 256, 2, 4, False),
    ("puzzle-00", 4, 8, 5, True),
]


def obs_golden():
    oracle.build()
    ref = oracle.load_ref_module()
    assert ref is not None, "oracle/_ref was not built (needs /root/reference)"
    P = ref.InputParameters(20, 13, 5, 256, 5, 5, 64, False)
    out = {}
    for k, (name, n, steps, seed, solid) in enumerate(SCENARIOS):
        m = maps.load_map(name, solid_padding=solid)
        grid = m["grid"]
        st, gl = maps.sample_instance(m, n, seed)
        g = ref.ObservationGenerator(grid.astype(int).tolist(), P)
        g.create_agents([tuple(x) for x in st.tolist()], [tuple(x) for x in gl.tolist()])
        rng = np.random.default_rng(100 + seed)
        pos = st.copy()
        act = np.full(n, -1, np.int32)
        P_, A_, T_ = [], [], []
        for t in range(steps):
            g.update_agents([tuple(x) for x in pos.tolist()], [tuple(x) for x in gl.tolist()], act.tolist())
            T_.append(np.asarray(g.generate_observations(), dtype=np.int8))
            P_.append(pos.copy())
            A_.append(act.copy())
            act = rng.integers(0, 5, n).astype(np.int32)            # "chosen" actions (history)
            mv = np.where(rng.random(n) < 0.8, act, rng.integers(0, 5, n)).astype(np.int32)  # executed moves
            pos, _ = oracle.pogema_step_soft(grid, pos, mv)
        out[f"s{k}_grid"] = grid
        out[f"s{k}_goals"] = gl
        out[f"s{k}_pos"] = np.stack(P_)
        out[f"s{k}_act"] = np.stack(A_)
        out[f"s{k}_tokens"] = np.stack(T_)
    # the reference's only known-answer scenario: int main(), observation_generator.cpp:530-544
    g = ref.ObservationGenerator(np.zeros((256, 256), int).tolist(), P)
    g.create_agents([(120, 120)], [(20, 200)])
    g.update_agents([(120, 120)], [(20, 200)], [0])
    row = np.asarray(g.generate_observations(), dtype=np.int8)[0]
    out["main_row"] = row
    assert hashlib.sha256(row.tobytes()).hexdigest() == \
        "896eb85aa89a369759917e5903f237dc28387e6b7d431fbd6703f302a97585e1"
    np.savez_compressed(OUT / "obs_golden.npz", **out)
    print("obs_golden.npz", (OUT / "obs_golden.npz").stat().st_size, "bytes")


def gpt_golden():
    from mapf_gpt.model import GPT, GPTConfig as RefCfg   # the reference, imported read-only
    obs = np.load(OUT / "obs_golden.npz")
    rows = np.concatenate([obs["s0_tokens"][3, :4], obs["s1_tokens"][2, :4], obs["s3_tokens"][1, :4]]).astype(np.int64)
    out = {"tokens": rows.astype(np.int8)}
    for name in ("2M", "6M", "85M"):
        cfg = W.model_config(name)
        for tag, scale in (("init", 1.0), ("sharp", 3.0)):
            if name == "85M" and tag == "init":
                continue
            sd = W.random_init(cfg, 1234)
            if tag == "sharp":
                sd = W.scale_weights(W.perturb_layernorm(sd), scale)
            net = GPT(RefCfg(**cfg.__dict__))
            missing = net.load_state_dict(sd, strict=False)
            assert not missing.missing_keys and not missing.unexpected_keys
            net.eval()
            idx = torch.from_numpy(rows)
            with torch.no_grad():
                logits, _ = net(idx)
            gen = torch.Generator(device="cpu")
            gen.manual_seed(0)
            acts = net.act(idx, generator=gen)
            greedy = net.act(idx, do_sample=False)
            out[f"{name}_{tag}_logits"] = logits[:, 0, :].numpy()
            out[f"{name}_{tag}_actions"] = acts.numpy().astype(np.int32)
            out[f"{name}_{tag}_greedy"] = greedy.numpy().astype(np.int32)
            out[f"{name}_{tag}_digest"] = np.frombuffer(bytes.fromhex(W.state_dict_digest(sd)), dtype=np.uint8)
            print(name, tag, "logits[0,:5]", logits[0, 0, :5].tolist())
    np.savez_compressed(OUT / "gpt_golden.npz", **out)
    print("gpt_golden.npz", (OUT / "gpt_golden.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    obs_golden()
    gpt_golden()
