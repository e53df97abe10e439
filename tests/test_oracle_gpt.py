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
