"""CPU: oracle/gpt_oracle.py (torch fp32 restatement of mapf_gpt/model.py) against golden logits
and sampled actions produced by the reference model itself (tests/golden/make_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from mapf_gpt_b200 import weights as W
from oracle import gpt_oracle as G

GOLD = np.load(Path(__file__).parent / "golden" / "gpt_golden.npz")
CASES = [("2M", "init"), ("2M", "sharp"), ("6M", "init"), ("6M", "sharp"), ("85M", "sharp")]


def make_sd(name, tag):
    cfg = W.model_config(name)
    sd = W.random_init(cfg, 1234)
    if tag == "sharp":
        sd = W.scale_weights(W.perturb_layernorm(sd), 3.0)
    return cfg, sd


@pytest.mark.parametrize("name,tag", CASES)
def test_logits_and_actions_match_reference(name, tag):
    cfg, sd = make_sd(name, tag)
    digest = bytes(GOLD[f"{name}_{tag}_digest"]).hex()
    assert W.state_dict_digest(sd) == digest, "seeded weights differ from the ones the golden run used"
    idx = torch.from_numpy(GOLD["tokens"].astype(np.int64))
    logits = G.forward_logits(sd, cfg.n_layer, cfg.n_head, idx)
    ref = torch.from_numpy(GOLD[f"{name}_{tag}_logits"])
    assert logits.shape == ref.shape == (idx.shape[0], 67)
    assert float((logits - ref).abs().max()) < 2e-5          # fp32 vs fp32, SDPA vs explicit softmax
    gen = torch.Generator(device="cpu")
    gen.manual_seed(0)
    acts = G.act(sd, cfg.n_layer, cfg.n_head, idx, generator=gen)
    assert acts.tolist() == GOLD[f"{name}_{tag}_actions"].tolist()
    assert G.act(sd, cfg.n_layer, cfg.n_head, idx, do_sample=False).tolist() == GOLD[f"{name}_{tag}_greedy"].tolist()


def test_multinomial_is_argmax_p_over_q():
    """torch.multinomial(p,1,g) == argmax(p / q), q = exponential_(1, g) of p's shape."""
    cfg, sd = make_sd("2M", "sharp")
    idx = torch.from_numpy(GOLD["tokens"].astype(np.int64))
    g1, g2 = torch.Generator().manual_seed(0), torch.Generator().manual_seed(0)
    a = G.act(sd, cfg.n_layer, cfg.n_head, idx, generator=g1)
    q = torch.empty((idx.shape[0], 67)).exponential_(1, generator=g2)
    b = G.act(sd, cfg.n_layer, cfg.n_head, idx, q=q)
    assert a.tolist() == b.tolist()


def test_param_counts():
    # 1 589 440 / 6 378 496 / 85 201 920 parameters (wte tied, wpe included)
    for name, n in (("2M", 1_589_440), ("6M", 6_378_496), ("85M", 85_201_920)):
        cfg = W.model_config(name)
        sd = W.random_init(cfg)
        assert sum(v.numel() for k, v in sd.items() if k != "lm_head.weight") == n


def test_oracle_rollout_equals_the_unmodified_reference_inference():
    """oracle/cpu_rollout.py (ported forward + sampler) against the UNMODIFIED mapf_gpt/inference.py + model.py + compiled
    generator loaded from baseline/_ref (oracle/ref_runtime.py): same obs dicts in, identical sampled actions out, step after
    step -- pins the port of GPT.act / _forward_batch / _prepare_inputs, not only its logits."""
    import numpy as np
    import pytest
    from oracle import cpu_rollout, ref_runtime
    from mapf_gpt_b200 import maps, weights as W
    if not ref_runtime.available():
        pytest.skip("baseline/_ref not built (needs /root/reference once: make -C oracle baseline_ref)")
    cfg = W.model_config("2M")
    sd = W.scale_weights(W.perturb_layernorm(W.random_init(cfg)), 3.0)
    m = maps.load_map("validation-random-seed-000")
    st, gl = maps.sample_instance(m, 12, 5)
    port = cpu_rollout.CpuRollout(m["grid"], st[None], gl[None], sd, cfg.n_layer, cfg.n_head)
    ref = ref_runtime.ReferenceRollout(m["grid"], st[None], gl[None], sd, cfg, device="cpu", mode="act")
    for t in range(6):
        _, acts = port.step()
        racts = ref.step()
        assert (np.asarray(racts[0]) == acts[0]).all() and (ref.pos == port.pos).all(), t
