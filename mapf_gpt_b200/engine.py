"""RolloutEngine: numpy-facing wrapper of the C ABI (include/mapf_gpt_b200.h).

Holds E environment slots x N agents in HBM and runs the whole MAPF-GPT step on the
device: update_agents -> tokenizer -> GPT forward -> sample -> POGEMA soft step.
Host arrays are slot-major and padded to the engine's agent capacity.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .weights import GPTConfig

KERNEL_CLASSES = ["bfs", "observe", "embed", "layernorm", "gemm_qkv", "attention", "gemm_attn_proj",
                  "gemm_fc_gelu", "gemm_mlp_proj", "head", "sample_step", "post_attn_fused", "attention_last_token",
                  "post_attn_last_token"]

MODE_GREEDY, MODE_PHILOX, MODE_SUPPLIED_Q = 0, 1, 2


def flatten_weights(sd: dict, cfg: GPTConfig) -> np.ndarray:
    """state_dict (SURVEY App. D.3 keys) -> the flat fp32 buffer mg_engine_load_model expects."""
    def f(k):
        t = sd[k]
        t = t.detach().cpu().float().numpy() if hasattr(t, "detach") else np.asarray(t, dtype=np.float32)
        return np.ascontiguousarray(t, dtype=np.float32).ravel()

    parts = [f("transformer.wte.weight"), f("transformer.wpe.weight")]
    for l in range(cfg.n_layer):
        p = f"transformer.h.{l}."
        parts += [f(p + "ln_1.weight"), f(p + "attn.c_attn.weight"), f(p + "attn.c_proj.weight"),
                  f(p + "ln_2.weight"), f(p + "mlp.c_fc.weight"), f(p + "mlp.c_proj.weight")]
    parts.append(f("transformer.ln_f.weight"))
    return np.concatenate(parts)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RolloutEngine:
    def __init__(self, max_envs: int, max_agents: int, H: int, W: int, device: int = 0, params=None):
        self._L = _lib.lib()
        p = _lib.MgParams()
        self._L.mg_default_params(C.byref(p))
        if params:
            for k, v in params.items():
                setattr(p, k, int(v))
        self.E, self.N, self.H, self.W = max_envs, max_agents, H, W
        self._h = self._L.mg_engine_create(device, max_envs, max_agents, H, W, C.byref(p))
        if not self._h:
            raise _lib.MgError(-1, self._L.mg_last_error().decode())
        self.cfg = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.mg_engine_destroy(self._h)
            self._h = None

    __del__ = close

    # ---- policy network
    def load_model(self, sd: dict, cfg: GPTConfig):
        if cfg.dropout != 0.0 or cfg.bias:
            raise ValueError("only dropout=0, bias=False checkpoints are supported (all shipped models)")
        flat = flatten_weights(sd, cfg)
        mc = _lib.MgModelConfig(cfg.block_size, cfg.vocab_size, cfg.n_layer, cfg.n_head, cfg.n_embd)
        _lib.check(self._L.mg_engine_load_model(self._h, C.byref(mc), _ptr(flat), flat.size))
        self.cfg = cfg

    # ---- environments
    @property
    def num_envs(self) -> int:
        return self._L.mg_engine_num_envs(self._h)

    def clear(self):
        """Forget every env slot; the loaded model and the per-map tables stay on the device."""
        _lib.check(self._L.mg_engine_clear(self._h))

    def reset(self, first_env: int, obstacles, pos, goal):
        """obstacles [n_envs,H,W] (or [H,W] broadcast), pos/goal [n_envs,n_agents,2]."""
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        goal = np.ascontiguousarray(goal, dtype=np.int32)
        if pos.ndim == 2:
            pos, goal = pos[None], goal[None]
        n_envs, n_agents = pos.shape[:2]
        ob = np.asarray(obstacles)
        if ob.ndim == 2:
            ob = np.broadcast_to(ob, (n_envs,) + ob.shape)
        ob = np.ascontiguousarray(ob != 0, dtype=np.uint8)
        if ob.shape != (n_envs, self.H, self.W):
            raise ValueError(f"obstacles shape {ob.shape} != {(n_envs, self.H, self.W)}")
        _lib.check(self._L.mg_engine_reset(self._h, first_env, n_envs, n_agents, _ptr(ob), _ptr(pos), _ptr(goal)))

    def _pad(self, a, last):
        """[num_envs, n, ...] -> contiguous int32 [num_envs, N, ...]"""
        if a is None:
            return None
        a = np.asarray(a, dtype=np.int32)
        shape = (self.num_envs, self.N) + ((last,) if last else ())
        if a.shape == shape:
            return np.ascontiguousarray(a)
        out = np.zeros(shape, dtype=np.int32)
        out[:, :a.shape[1]] = a
        return out

    def update_agents(self, pos=None, goal=None, actions=None):
        p, g, a = self._pad(pos, 2), self._pad(goal, 2), self._pad(actions, 0)
        _lib.check(self._L.mg_engine_update_agents(self._h, _ptr(p), _ptr(g), _ptr(a)))

    def generate_observations(self, fetch: bool = True):
        out = np.empty((self.num_envs, self.N, 256), dtype=np.int8) if fetch else None
        _lib.check(self._L.mg_engine_generate_observations(self._h, _ptr(out)))
        return out

    def act(self, mode: int = MODE_PHILOX, q=None, want_logits: bool = False):
        ne = self.num_envs
        acts = np.empty((ne, self.N), dtype=np.int32)
        logits = np.empty((ne, self.N, 5), dtype=np.float32) if want_logits else None
        qq = None if q is None else np.ascontiguousarray(q, dtype=np.float32)
        _lib.check(self._L.mg_engine_act(self._h, mode, _ptr(qq), _ptr(acts), _ptr(logits)))
        return (acts, logits) if want_logits else acts

    def act_host(self, pos, goal, mode: int = MODE_PHILOX, q=None):
        p, g = self._pad(pos, 2), self._pad(goal, 2)
        acts = np.empty((self.num_envs, self.N), dtype=np.int32)
        qq = None if q is None else np.ascontiguousarray(q, dtype=np.float32)
        _lib.check(self._L.mg_engine_act_host(self._h, _ptr(p), _ptr(g), mode, _ptr(qq), _ptr(acts)))
        return acts

    def forward_tokens(self, tokens) -> np.ndarray:
        t = np.ascontiguousarray(tokens, dtype=np.int8)
        assert t.ndim == 2 and t.shape[1] == 256
        out = np.empty((t.shape[0], 5), dtype=np.float32)
        _lib.check(self._L.mg_engine_forward_tokens(self._h, _ptr(t), t.shape[0], _ptr(out)))
        return out

    def eval_tokens(self, tokens, targets):
        """Dataset rows -> (cross-entropy of the training objective per row, arg-max action per row): model.py:180-183 with
        targets only at position 255 (dataset/fast_data_loader.py:57).  tokens int8 [n, 256], targets int8 [n] (-1 = ignore)."""
        t = np.ascontiguousarray(tokens, dtype=np.int8)
        y = np.ascontiguousarray(targets, dtype=np.int8)
        assert t.ndim == 2 and t.shape[1] == 256 and y.shape == (t.shape[0],)
        loss = np.empty(t.shape[0], dtype=np.float32)
        pred = np.empty(t.shape[0], dtype=np.int32)
        _lib.check(self._L.mg_engine_eval_tokens(self._h, _ptr(t), _ptr(y), t.shape[0], _ptr(loss), _ptr(pred)))
        return loss, pred

    def env_step(self, actions=None, fetch: bool = True):
        a = self._pad(actions, 0)
        out = np.empty((self.num_envs, self.N, 2), dtype=np.int32) if fetch else None
        _lib.check(self._L.mg_engine_env_step(self._h, _ptr(a), _ptr(out)))
        return out

    def rollout(self, n_steps: int, mode: int = MODE_PHILOX):
        _lib.check(self._L.mg_engine_rollout(self._h, n_steps, mode))

    def synchronize(self):
        _lib.check(self._L.mg_engine_synchronize(self._h))

    def set_active(self, mask=None):
        """Restrict the next calls to the slots whose mask entry is non-zero (None = all slots)."""
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        _lib.check(self._L.mg_engine_set_active(self._h, _ptr(m)))

    def set_seed(self, seed: int):
        _lib.check(self._L.mg_engine_set_seed(self._h, seed))

    def set_env_offset(self, off: int):
        _lib.check(self._L.mg_engine_set_env_offset(self._h, off))

    def set_max_episode_steps(self, n: int):
        _lib.check(self._L.mg_engine_set_max_episode_steps(self._h, n))

    # ---- read-back
    def positions(self) -> np.ndarray:
        out = np.empty((self.num_envs, self.N, 2), dtype=np.int32)
        _lib.check(self._L.mg_engine_get_positions(self._h, _ptr(out)))
        return out

    def tokens(self) -> np.ndarray:
        out = np.empty((self.num_envs, self.N, 256), dtype=np.int8)
        _lib.check(self._L.mg_engine_get_tokens(self._h, _ptr(out)))
        return out

    def cost2go(self, env: int, agent: int) -> np.ndarray:
        out = np.empty((self.H, self.W), dtype=np.uint16)
        _lib.check(self._L.mg_engine_get_cost2go(self._h, env, agent, _ptr(out)))
        return out

    def partial(self, env: int, agent: int):
        """((left, right, top, bottom), field[rows, cols]) of one agent's cost-to-go window."""
        b = np.zeros(4, dtype=np.int32)
        n = self._L.mg_engine_get_partial(self._h, env, agent, _ptr(b), None, 0)
        if n < 0:
            _lib.check(n)
        buf = np.empty(n, dtype=np.uint16)
        n2 = self._L.mg_engine_get_partial(self._h, env, agent, _ptr(b), _ptr(buf), n)
        if n2 < 0:
            _lib.check(n2)
        left, right, top, bottom = (int(v) for v in b)
        return (left, right, top, bottom), buf.reshape(right - left + 1, bottom - top + 1)

    def metrics(self) -> np.ndarray:
        """[num_envs, 10]: ep_length, CSR, ISR, SoC, makespan, on_goal_now, agent_steps, n_agents,
        avg_agents_density, density samples"""
        out = np.empty((self.num_envs, _lib.MG_METRIC_COLS), dtype=np.float64)
        _lib.check(self._L.mg_engine_get_metrics(self._h, _ptr(out)))
        return out

    # ---- timing
    def set_profiling(self, on: bool):
        _lib.check(self._L.mg_engine_set_profiling(self._h, int(on)))

    def last_timing(self):
        tot = C.c_float(0)
        ph = (C.c_float * 3)()
        _lib.check(self._L.mg_engine_last_timing(self._h, C.byref(tot), ph))
        return tot.value, list(ph)

    def launch_count(self) -> int:
        return int(self._L.mg_engine_launch_count(self._h))

    def num_lanes(self) -> int:
        """Stream lanes of the forward (2 = attention and post-attention kernels of alternate chunks overlap)."""
        return int(self._L.mg_engine_num_lanes(self._h))

    def kernel_times(self) -> dict:
        buf = (C.c_float * (2 * len(KERNEL_CLASSES)))()
        self._L.mg_engine_kernel_times(self._h, buf, len(buf))
        return {k: {"ms": buf[2 * i], "launches": int(buf[2 * i + 1])} for i, k in enumerate(KERNEL_CLASSES)}


def test_gemm(A, B, variant: int = 0):
    """torch bf16 CUDA tensors A[M,K], B[N,K] -> C[M,N] fp32 via the production tcgen05 GEMM."""
    import torch
    M, K = A.shape
    N = B.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    _lib.check(_lib.lib().mg_test_gemm(A.device.index or 0, A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, variant))
    return out


def test_attention(q, k, v, variant: int = 0):
    """torch bf16 CUDA tensors [n_seq, n_head, 256, hs] -> same shape, via the production kernels.
    variant 0: max-subtracting softmax, 1: max-free softmax on pre-scaled q (engine default), 2: classic kernel."""
    import torch
    n_seq, n_head, T, hs = q.shape
    assert T == 256
    out = torch.empty_like(q)
    _lib.check(_lib.lib().mg_test_attention_ex(q.device.index or 0, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                               out.data_ptr(), n_seq, n_head, hs, variant))
    return out


def set_precision(mode: str) -> str:
    """'bf16' (tensor-core path) or 'fp32' (CUDA-core verification path) for engines created afterwards."""
    prev = _lib.lib().mg_set_precision({"bf16": 0, "fp32": 1}[mode])
    return "fp32" if prev == 1 else "bf16"
