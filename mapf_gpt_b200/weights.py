"""Checkpoint handling for the policy network (host side, plumbing only).

Mirrors what the reference does at construction (mapf_gpt/inference.py:72-85):
torch.load -> checkpoint["model"] with the "_orig_mod." prefix stripped
(inference.py:33-44) + checkpoint["model_args"] -> GPTConfig (model.py:107-115).
Pretrained weights need the network (inference.py:53-56), so benchmarks and tests use
`random_init`, a seeded numpy re-statement of the reference init distribution
(model.py:141-145,159-165): N(0, 0.02), c_proj N(0, 0.02/sqrt(2L)), LayerNorm gain 1.
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, asdict

import numpy as np
import torch


@dataclass
class GPTConfig:  # same fields and defaults as model.py:107-115
    block_size: int = 161
    vocab_size: int = 67
    n_layer: int = 8
    n_head: int = 8
    n_embd: int = 256
    dropout: float = 0.0
    bias: bool = False


# experiment_setup/config-{2M,6M,85M}.py:7-13
MODEL_SHAPES = {
    "2M": dict(n_layer=5, n_head=5, n_embd=160),
    "6M": dict(n_layer=8, n_head=8, n_embd=256),
    "85M": dict(n_layer=12, n_head=12, n_embd=768),
}


def model_config(name: str) -> GPTConfig:
    return GPTConfig(block_size=256, vocab_size=67, dropout=0.0, bias=False, **MODEL_SHAPES[name])


def strip_prefix(state_dict: dict, prefix: str = "_orig_mod.") -> dict:
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in state_dict.items()}


def random_init(cfg: GPTConfig, seed: int = 1234) -> dict:
    """Seeded fp32 state_dict with the reference's key names."""
    rng = np.random.default_rng(seed)
    C, L = cfg.n_embd, cfg.n_layer

    def normal(shape, std):
        return torch.from_numpy((rng.standard_normal(shape, dtype=np.float32) * np.float32(std)))

    sd = {}
    sd["transformer.wte.weight"] = normal((cfg.vocab_size, C), 0.02)
    sd["transformer.wpe.weight"] = normal((cfg.block_size, C), 0.02)
    pstd = 0.02 / math.sqrt(2 * L)
    for l in range(L):
        p = f"transformer.h.{l}."
        sd[p + "ln_1.weight"] = torch.ones(C)
        sd[p + "attn.c_attn.weight"] = normal((3 * C, C), 0.02)
        sd[p + "attn.c_proj.weight"] = normal((C, C), pstd)
        sd[p + "ln_2.weight"] = torch.ones(C)
        sd[p + "mlp.c_fc.weight"] = normal((4 * C, C), 0.02)
        sd[p + "mlp.c_proj.weight"] = normal((C, 4 * C), pstd)
    sd["transformer.ln_f.weight"] = torch.ones(C)
    sd["lm_head.weight"] = sd["transformer.wte.weight"]  # tied, model.py:138
    return sd


def perturb_layernorm(sd: dict, seed: int = 7, scale: float = 0.1) -> dict:
    """Make LayerNorm gains non-trivial so tests exercise them."""
    rng = np.random.default_rng(seed)
    out = dict(sd)
    for k in sd:
        if k.endswith("ln_1.weight") or k.endswith("ln_2.weight") or k.endswith("ln_f.weight"):
            out[k] = sd[k] + torch.from_numpy(rng.standard_normal(sd[k].shape, dtype=np.float32) * np.float32(scale))
    return out


def scale_weights(sd: dict, factor: float) -> dict:
    """Scale the matmul weights (not embeddings / LN) - random 0.02-std weights give
    near-uniform attention and tiny logits; tests use a larger scale to get sharp
    softmaxes and decisive logits."""
    out = dict(sd)
    for k in sd:
        if k.endswith("c_attn.weight") or k.endswith("c_fc.weight") or k.endswith("c_proj.weight"):
            out[k] = sd[k] * factor
    return out


def state_dict_digest(sd: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def save_checkpoint(path, sd: dict, cfg: GPTConfig) -> None:
    """Reference .pt layout (train.py:301-308): {'model': state_dict, 'model_args': {...}}."""
    torch.save({"model": {k: v.clone() for k, v in sd.items()}, "model_args": asdict(cfg)}, path)


def load_checkpoint(path, map_location="cpu"):
    ck = torch.load(path, map_location=map_location)
    sd = strip_prefix(ck["model"])
    cfg = GPTConfig(**ck["model_args"])
    if "lm_head.weight" not in sd:
        sd["lm_head.weight"] = sd["transformer.wte.weight"]
    return sd, cfg
