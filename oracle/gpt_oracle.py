"""gpt_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain PyTorch fp32 restatement of the reference policy network's inference path
(mapf_gpt/model.py): GPT.forward (model.py:167-189) with LayerNorm (model.py:11-20),
NonCausalSelfAttention (model.py:23-72), MLP (model.py:75-89), Block (model.py:92-104)
and GPT.act (model.py:244-260).  Functional: it takes the checkpoint's state_dict
(keys of SURVEY App. D.3) instead of building nn.Modules.

Parity status: PINNED against the reference itself.  tests/golden/make_golden.py
imports /root/reference/mapf_gpt/model.py in the build container, loads the same
seeded weights into the reference GPT and stores its logits / sampled actions in
tests/golden/gpt_*.npz; tests/test_oracle_gpt.py compares this file with them
(fp32 on CPU, tolerance 2e-5 abs on logits; actions exact).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _ln(x, w):
    # model.py:19-20: F.layer_norm(input, weight.shape, weight, bias=None, 1e-5)
    return F.layer_norm(x, w.shape, w, None, 1e-5)


def block_forward(sd, l: int, x, n_head: int, return_parts: bool = False):
    """One Block (model.py:101-104)."""
    p = f"transformer.h.{l}."
    B, T, C = x.shape
    h = _ln(x, sd[p + "ln_1.weight"])
    qkv = F.linear(h, sd[p + "attn.c_attn.weight"])                  # model.py:50
    q, k, v = qkv.split(C, dim=2)
    hs = C // n_head
    q = q.view(B, T, n_head, hs).transpose(1, 2)
    k = k.view(B, T, n_head, hs).transpose(1, 2)
    v = v.view(B, T, n_head, hs).transpose(1, 2)
    att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hs))          # model.py:58-66, is_causal=False
    att = F.softmax(att, dim=-1)
    y = (att @ v).transpose(1, 2).contiguous().view(B, T, C)
    x = x + F.linear(y, sd[p + "attn.c_proj.weight"])                # model.py:71,102
    h2 = _ln(x, sd[p + "ln_2.weight"])
    m = F.gelu(F.linear(h2, sd[p + "mlp.c_fc.weight"]))              # exact erf GELU, model.py:80,85-86
    x = x + F.linear(m, sd[p + "mlp.c_proj.weight"])                 # model.py:87,103
    return x


@torch.no_grad()
def forward_logits(sd, n_layer: int, n_head: int, idx: torch.Tensor, return_hidden: bool = False):
    """idx (B,T) long -> logits (B, vocab) of the LAST position (model.py:186)."""
    wte = sd["transformer.wte.weight"]
    wpe = sd["transformer.wpe.weight"]
    B, T = idx.shape
    x = wte[idx] + wpe[torch.arange(T, device=idx.device)]          # model.py:171-175
    hidden = [x]
    for l in range(n_layer):
        x = block_forward(sd, l, x, n_head)
        hidden.append(x)
    x = _ln(x, sd["transformer.ln_f.weight"])                        # model.py:178
    logits = F.linear(x[:, -1, :], wte)                              # tied lm_head, model.py:138,186
    return (logits, hidden) if return_hidden else logits


@torch.no_grad()
def act(sd, n_layer, n_head, idx, do_sample=True, generator=None, q=None):
    """GPT.act (model.py:244-260).  `q`: optional pre-drawn Exp(1) tensor (B,vocab) to
    make the multinomial draw explicit: multinomial(p,1) == argmax(p / q)."""
    logits = forward_logits(sd, n_layer, n_head, idx)
    masked = torch.full_like(logits, float("-inf"))
    masked[:, :5] = logits[:, :5]
    probs = F.softmax(masked, dim=-1)
    if not do_sample:
        return probs.argmax(dim=-1)
    if q is not None:
        return (probs / q).argmax(dim=-1)
    return torch.multinomial(probs, num_samples=1, generator=generator).squeeze(-1)
