"""Drop-in for mapf_gpt/inference.py: same config fields, same class surface
(`MAPFGPTInference(cfg, net=None)`, `.act`, `.act_batch`, `.reset_states`, `build`),
backed by the B200 engine instead of the C++ generator + PyTorch model.

Reference behaviour mirrored (mapf_gpt/inference.py):
  * :13-31   config fields and defaults (pydantic, extra=forbid)
  * :48-85   weight loading from the reference .pt layout, RNG seeded 0
  * :127-146 one generator state per env slot key `pos`, created on first sight
             (history "n"x5, last actions -1), then update_agents + generate_observations
  * :87-101  forward in chunks of cfg.batch_size rows; one multinomial draw per chunk
  * :151-172 act_batch bookkeeping (_last_actions feed the action history)
  * :174-177 reset_states drops all slots and reseeds the sampling generator to 0
Differences, on purpose: there is no CPU/MPS fallback (a missing GPU or library is an
error), and `sampling="philox"` offers an in-kernel RNG for throughput runs.
"""
from __future__ import annotations

from pathlib import Path
from typing import Literal, Optional

import numpy as np
from pydantic import BaseModel

try:  # the harness's own base class when it is installed
    from pogema_toolbox.algorithm_config import AlgoBase  # type: ignore
except Exception:  # pogema_toolbox absent: field-compatible stand-in
    class AlgoBase(BaseModel):
        name: Optional[str] = None
        num_process: int = 3
        device: Optional[str] = "cuda"
        parallel_backend: Optional[str] = "multiprocessing"
        seed: Optional[int] = 0
        preprocessing: Optional[str] = None

from . import _lib
from .engine import MODE_GREEDY, MODE_PHILOX, MODE_SUPPLIED_Q, RolloutEngine
from .weights import GPTConfig, load_checkpoint

HF_WEIGHT_NAMES = ["MAPF-GPT-2M.pt", "MAPF-GPT-6M.pt", "MAPF-GPT-85M.pt", "MAPF-GPT-DDG-2M.pt"]  # inference.py:54


class MAPFGPTInferenceConfig(AlgoBase, extra="forbid"):
    name: Literal["MAPF-GPT"] = "MAPF-GPT"
    num_agents: int = 13
    num_previous_actions: int = 5
    cost2go_value_limit: int = 20
    agents_radius: int = 5
    cost2go_radius: int = 5
    path_to_weights: Optional[str] = "weights/MAPF-GPT-2M.pt"
    device: Optional[str] = None
    context_size: int = 256
    mask_actions_history: bool = False
    mask_goal: bool = False
    mask_cost2go: bool = False
    mask_greed_action: bool = False
    repo_id: str = "aandreychuk/MAPF-GPT"
    grid_step: int = 64
    save_cost2go: bool = False
    batch_size: int = 2048
    num_process: int = 8


def _device_index(device: Optional[str]) -> int:
    if device is None or device == "cuda":
        return 0
    if device.startswith("cuda:"):
        return int(device.split(":", 1)[1])
    raise RuntimeError(f"device {device!r} requested, but mapf_gpt_b200 runs on CUDA (B200) only: "
                       f"there is no CPU/MPS fallback")


class MAPFGPTInference:
    def __init__(self, cfg: MAPFGPTInferenceConfig, net=None, *, sampling: str = "torch",
                 do_sample: bool = True, max_envs: Optional[int] = None):
        """net: optional (state_dict, GPTConfig) pair used instead of loading cfg.path_to_weights
        (the reference accepts a pre-built module here, inference.py:79-80).
        sampling: "torch" reproduces torch.multinomial's draw with a torch.Generator seeded 0;
        "philox" uses the engine's counter-based stream (not the reference's numbers)."""
        self.cfg = cfg
        self._L = _lib.lib()  # ImportError when the CUDA library is missing
        if self._L.mg_device_count() < 1:
            raise RuntimeError("no CUDA device visible: mapf_gpt_b200 has no CPU fallback")
        self._dev = _device_index(cfg.device)
        self.cfg.device = f"cuda:{self._dev}" if cfg.device not in (None, "cuda") else "cuda"
        if sampling not in ("torch", "philox"):
            raise ValueError("sampling must be 'torch' or 'philox'")
        self._sampling, self._do_sample = sampling, do_sample
        self._max_envs = max_envs

        if net is not None:
            self._sd, self._gpt_cfg = net
        else:
            path = Path(cfg.path_to_weights)
            if not path.exists() and path.name in HF_WEIGHT_NAMES:
                try:
                    from huggingface_hub import hf_hub_download
                    hf_hub_download(repo_id=cfg.repo_id, filename=path.name, local_dir=path.parent)
                except Exception as ex:
                    raise FileNotFoundError(f"{path} not found and could not be downloaded from "
                                            f"{cfg.repo_id}: {ex}") from ex
            self._sd, self._gpt_cfg = load_checkpoint(path)
        import torch
        self._torch = torch
        self.torch_generator = torch.Generator(device=f"cuda:{self._dev}")
        self.torch_generator.manual_seed(0)
        self._engine: Optional[RolloutEngine] = None
        self._tok_engine: Optional[RolloutEngine] = None
        self._slots: dict = {}          # slot key -> env index
        self._slot_n: dict = {}         # slot key -> agent count
        self._last_actions: dict = {}
        self._step = 0
        self._all_active = True

    # ------------------------------------------------------------------ reference surface
    @staticmethod
    def build():
        """Pre-build hook (benchmark.py:26): compile the CUDA library once, before fan-out."""
        _lib.build()
        _lib.lib()

    def reset_states(self):
        # O(1) like the reference (inference.py:174-177): the engine, its loaded model and the per-map tables stay on the
        # device; only the env slots are forgotten.  A later episode that does not fit re-sizes the engine (_ensure_slots).
        if self._engine is not None:
            self._engine.clear()
        self._slots, self._slot_n, self._last_actions = {}, {}, {}
        self.torch_generator.manual_seed(0)
        self._step = 0
        self._all_active = True

    def act(self, observations):
        return self.act_batch([observations])[0]

    def act_batch(self, observations_list, positions=None):
        if positions is None:
            positions = list(range(len(observations_list)))
        if len(observations_list) == 0:
            return []
        if not isinstance(observations_list[0][0], dict):
            return self._act_tokens(observations_list, positions)
        self._ensure_slots(observations_list, positions)
        eng = self._engine
        E, N = eng.num_envs, eng.N
        # slots not present in this call keep their state: resend what was sent last
        pos = self._pos[:E].copy()
        goal = self._goals[:E].copy()
        counts = []
        for key, obs in zip(positions, observations_list):
            e = self._slots[key]
            n = len(obs)
            if n != self._slot_n[key]:
                raise ValueError(f"slot {key!r}: agent count changed from {self._slot_n[key]} to {n}")
            pos[e, :n] = np.asarray([o["global_xy"] for o in obs], dtype=np.int32)
            goal[e, :n] = np.asarray([o["global_target_xy"] for o in obs], dtype=np.int32)
            counts.append(n)
        self._pos[:E], self._goals[:E] = pos, goal
        mode, q = self._draw(positions, counts, E, N)
        present = np.zeros(E, dtype=np.uint8)
        present[[self._slots[k] for k in positions]] = 1
        if not present.all() or not self._all_active:          # only the addressed slots are updated (inference.py:158-160)
            eng.set_active(None if present.all() else present)
            self._all_active = bool(present.all())
        acts = eng.act_host(pos, goal, mode, q)
        results = []
        for key, n in zip(positions, counts):
            a = acts[self._slots[key], :n].tolist()
            self._last_actions[key] = list(a)
            results.append(a)
        self._step += 1
        return results

    # ------------------------------------------------------------------ internals
    def _draw(self, positions, counts, E, N):
        """Sampling noise laid out for the engine's padded [E, N] rows."""
        if not self._do_sample:
            return MODE_GREEDY, None
        if self._sampling == "philox":
            return MODE_PHILOX, None
        torch = self._torch
        total = int(sum(counts))
        qs = []
        for i in range(0, total, self.cfg.batch_size):       # one draw per chunk (inference.py:88-95)
            b = min(self.cfg.batch_size, total - i)
            qs.append(torch.empty((b, 67), dtype=torch.float32, device=self.torch_generator.device)
                      .exponential_(1, generator=self.torch_generator)[:, :5])
        qcat = torch.cat(qs).cpu().numpy()
        q = np.ones((E, N, 5), dtype=np.float32)
        off = 0
        for key, n in zip(positions, counts):
            q[self._slots[key], :n] = qcat[off:off + n]
            off += n
        return MODE_SUPPLIED_Q, q

    def _ensure_slots(self, observations_list, positions):
        new = [(k, o) for k, o in zip(positions, observations_list) if k not in self._slots]
        if not new:
            return
        grids = [np.asarray(o[0]["global_obstacles"]) for o in observations_list]
        H = max(g.shape[0] for g in grids)
        Wd = max(g.shape[1] for g in grids)
        N = max(len(o) for o in observations_list)
        E = self._max_envs or len(observations_list)
        eng = self._engine
        if eng is not None and not self._slots and (E > eng.E or N > eng.N or H > eng.H or Wd > eng.W):
            eng.close()                  # a fresh episode (reset_states) that outgrew the engine: size a new one
            self._engine = None
        if self._engine is None:
            params = dict(cost2go_value_limit=self.cfg.cost2go_value_limit, num_agents=self.cfg.num_agents,
                          num_previous_actions=self.cfg.num_previous_actions, context_size=self.cfg.context_size,
                          obs_radius=self.cfg.cost2go_radius, agents_radius=self.cfg.agents_radius,
                          grid_step=self.cfg.grid_step, save_cost2go=int(self.cfg.save_cost2go))
            self._engine = RolloutEngine(E, N, H, Wd, device=self._dev, params=params)
            self._engine.load_model(self._sd, self._gpt_cfg)
            self._pos = np.zeros((E, N, 2), dtype=np.int32)
            self._goals = np.zeros((E, N, 2), dtype=np.int32)
        eng = self._engine
        for key, obs in new:
            e = len(self._slots)
            n = len(obs)
            g = np.asarray(obs[0]["global_obstacles"])
            if e >= eng.E or n > eng.N or g.shape[0] > eng.H or g.shape[1] > eng.W:
                raise RuntimeError(
                    f"slot {key!r} does not fit the engine sized at the first act_batch call "
                    f"(capacity {eng.E} envs x {eng.N} agents, grid {eng.H}x{eng.W}); construct "
                    f"MAPFGPTInference(..., max_envs=...) or call reset_states() first")
            grid = np.ones((eng.H, eng.W), dtype=np.uint8)   # pad with obstacles, origin unchanged
            grid[:g.shape[0], :g.shape[1]] = g != 0
            p = np.asarray([o["global_xy"] for o in obs], dtype=np.int32)
            t = np.asarray([o["global_target_xy"] for o in obs], dtype=np.int32)
            eng.reset(e, grid[None], p[None], t[None])
            self._slots[key], self._slot_n[key] = e, n
            self._last_actions[key] = [-1] * n
            self._pos[e, :n], self._goals[e, :n] = p, t

    def _act_tokens(self, observations_list, positions):
        """Pre-tokenized rows (inference.py:146): forward + sample only."""
        rows = np.concatenate([np.asarray(o, dtype=np.int64) for o in observations_list]).astype(np.int8)
        if self._tok_engine is None:
            self._tok_engine = RolloutEngine(1, 1, 11, 11, device=self._dev)
            self._tok_engine.load_model(self._sd, self._gpt_cfg)
        logits = self._tok_engine.forward_tokens(rows)          # the engine's kernels; ids outside [0, 67) raise (MG_ERR_VOCAB)
        lg = logits.astype(np.float64)
        p = np.exp(lg - lg.max(-1, keepdims=True))
        p /= p.sum(-1, keepdims=True)
        if self._do_sample:
            # torch.multinomial(probs, 1, generator) == argmax(probs / q), q ~ Exp(1) of shape (rows, 67) drawn per chunk of
            # batch_size rows (inference.py:88-95); the masked entries have probability 0 and never win
            torch = self._torch
            qs = [torch.empty((min(self.cfg.batch_size, len(rows) - i), 67), dtype=torch.float32, device=self.torch_generator.device)
                  .exponential_(1, generator=self.torch_generator)[:, :5] for i in range(0, len(rows), self.cfg.batch_size)]
            q = torch.cat(qs).cpu().numpy().astype(np.float64)
            acts = (p / q).argmax(-1).tolist()
        else:
            acts = p.argmax(-1).tolist()
        out, off = [], 0
        for key, o in zip(positions, observations_list):
            out.append(acts[off:off + len(o)])
            self._last_actions[key] = list(out[-1])
            off += len(o)
        return out
