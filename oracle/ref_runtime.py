"""ref_runtime.py -- TEST/BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Loads the UNMODIFIED reference package (mapf_gpt/inference.py + model.py + its compiled observation generator) from the
git-ignored baseline/_ref (built by oracle/Makefile `baseline_ref`) so that bench.py's baseline arms run the reference's
own public API -- MAPFGPTInference(cfg).act_batch(obs dicts) -- on the host cores (`cpu_baseline`, `--impl reference`) and
on the B200 with stock PyTorch kernels (`stock_gpu`, reference device='cuda', inference.py:58-60).

The reference imports three packages that are not installable offline; they are replaced by the minimal stand-ins of
SURVEY App. E, none of which touches the timed path:
  pogema_toolbox.algorithm_config.AlgoBase   pydantic base with the toolbox's fields
  pogema_toolbox.registry.ToolboxRegistry    info()/warning() loggers
  cppimport.import_hook                      empty (the generator is already compiled next to the package)
  loguru.logger                              std logging (only when loguru is missing)
POGEMA itself is replaced by oracle/pogema_oracle.c (PARITY UNPINNED) as the environment that feeds act().
"""
from __future__ import annotations

import logging
import sys
import tempfile
import time
import types
from pathlib import Path
from typing import Optional

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF_DIR = ROOT / "baseline" / "_ref"
_MOD = None


def available() -> bool:
    return (REF_DIR / "mapf_gpt" / "inference.py").exists() and any((REF_DIR / "mapf_gpt").glob("observation_generator*.so"))


def _install_stubs():
    from pydantic import BaseModel
    if "pogema_toolbox" not in sys.modules:
        class AlgoBase(BaseModel):
            name: Optional[str] = None
            num_process: int = 3
            device: Optional[str] = "cuda"
            parallel_backend: Optional[str] = "multiprocessing"
            seed: Optional[int] = 0
            preprocessing: Optional[str] = None

        log = logging.getLogger("toolbox-stub")

        class ToolboxRegistry:
            info = staticmethod(log.info)
            warning = staticmethod(log.warning)
            debug = staticmethod(log.debug)

        pt = types.ModuleType("pogema_toolbox")
        ac = types.ModuleType("pogema_toolbox.algorithm_config")
        rg = types.ModuleType("pogema_toolbox.registry")
        ac.AlgoBase, rg.ToolboxRegistry = AlgoBase, ToolboxRegistry
        pt.algorithm_config, pt.registry = ac, rg
        sys.modules.update({"pogema_toolbox": pt, "pogema_toolbox.algorithm_config": ac, "pogema_toolbox.registry": rg})
    if "cppimport" not in sys.modules:
        try:
            import cppimport  # noqa: F401
        except Exception:
            ci = types.ModuleType("cppimport")
            ih = types.ModuleType("cppimport.import_hook")
            ci.import_hook = ih
            sys.modules.update({"cppimport": ci, "cppimport.import_hook": ih})
    try:
        import loguru  # noqa: F401
    except Exception:
        lg = types.ModuleType("loguru")
        lg.logger = logging.getLogger("loguru-stub")
        sys.modules["loguru"] = lg


def load():
    """-> the reference's mapf_gpt.inference module (MAPFGPTInference, MAPFGPTInferenceConfig), or None."""
    global _MOD
    if _MOD is None:
        if not available():
            return None
        _install_stubs()
        if str(REF_DIR) not in sys.path:
            sys.path.insert(0, str(REF_DIR))
        import importlib
        if "observation_generator" in sys.modules:      # oracle.load_ref_module() got there first: same binary, reuse it
            sys.modules.setdefault("mapf_gpt.observation_generator", sys.modules["observation_generator"])
        _MOD = importlib.import_module("mapf_gpt.inference")
        assert Path(_MOD.__file__).resolve().is_relative_to(REF_DIR.resolve()), _MOD.__file__
    return _MOD


def save_reference_checkpoint(sd: dict, cfg, path: Path) -> None:
    """The reference .pt layout (inference.py:72-78): {'model': state_dict, 'model_args': GPTConfig kwargs}."""
    import torch
    args = dict(block_size=cfg.block_size, vocab_size=cfg.vocab_size, n_layer=cfg.n_layer, n_head=cfg.n_head,
                n_embd=cfg.n_embd, dropout=0.0, bias=False)
    torch.save({"model": {k: v.clone() for k, v in sd.items()}, "model_args": args}, path)


class ReferenceRollout:
    """E envs x n agents driven through the reference's MAPFGPTInference exactly as pogema_toolbox.run_episode does
    (obs dicts in, action lists out), with oracle/pogema_oracle.c as the environment.

    mode "act":       one act(obs) call per env per step (what example.py:65 and every Dask worker do)
    mode "act_batch": one act_batch(all envs) call per step (the reference's own batched entry point, inference.py:151)
    """

    def __init__(self, grid, starts, goals, sd, cfg, device="cpu", mode="act_batch", torch_threads=None, autocast_bf16=False):
        import torch
        mod = load()
        if mod is None:
            raise RuntimeError("baseline/_ref is missing: run `make -C oracle baseline_ref` where /root/reference exists")
        self.torch = torch
        self.grid = np.ascontiguousarray(grid, dtype=np.int64)
        self.pos = np.array(starts, dtype=np.int32).copy()
        self.goals = np.array(goals, dtype=np.int32)
        self.E, self.n = self.pos.shape[:2]
        self.mode, self.device, self.autocast = mode, device, autocast_bf16
        self._tmp = tempfile.TemporaryDirectory()
        ck = Path(self._tmp.name) / f"rand-{cfg.n_layer}x{cfg.n_embd}.pt"      # not one of the HF names: no download attempt
        save_reference_checkpoint(sd, cfg, ck)
        self.algo = mod.MAPFGPTInference(mod.MAPFGPTInferenceConfig(path_to_weights=str(ck), device=device))
        self.algo.reset_states()
        self.torch_threads = torch_threads
        self._goal_t = [[tuple(int(v) for v in g) for g in self.goals[e]] for e in range(self.E)]

    def _obs(self, e):
        g = self.grid
        return [{"global_obstacles": g, "global_xy": tuple(int(v) for v in self.pos[e, i]), "global_target_xy": self._goal_t[e][i]}
                for i in range(self.n)]

    def step(self):
        import oracle
        torch = self.torch
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if self.autocast else _null()
        with ctx:
            if self.mode == "act":
                acts = [self.algo.act(self._obs(e)) for e in range(self.E)]
            else:
                acts = self.algo.act_batch([self._obs(e) for e in range(self.E)])
        if self.torch_threads:          # the generator's ctor pinned OpenMP to one thread (observation_generator.h:115)
            torch.set_num_threads(self.torch_threads)
        for e in range(self.E):
            self.pos[e], _ = oracle.pogema_step_soft(self.grid, self.pos[e], np.asarray(acts[e], np.int32))
        return acts


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def time_rollout(r: ReferenceRollout, steps: int, warmup: int = 1):
    import torch
    for _ in range(warmup):
        r.step()
    if r.device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        r.step()
    if r.device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return r.E * r.n * steps / dt, dt


# ------------------------------------------------------------------------------------------------ "ref-fair": the reference's own
# scaling model -- `num_process` worker processes (eval_configs/*/0*.yaml: parallel_backend balanced_dask, num_process), each
# running whole episodes through one act() per step on ONE thread.  Plain subprocesses with hard
# timeouts (a worker that fails to start must not hang the bench): every worker warms up, reports ready, waits for the go file.
def _fair_worker_main(argv):
    import torch
    worker, model, map_name, agents, steps, sync_dir = int(argv[0]), argv[1], argv[2], int(argv[3]), int(argv[4]), Path(argv[5])
    warmup = int(argv[6]) if len(argv) > 6 else 1
    sys.path.insert(0, str(ROOT))
    from mapf_gpt_b200 import maps, weights as W
    torch.set_num_threads(1)
    cfg = W.model_config(model)
    sd = W.random_init(cfg, 1234)
    m = maps.load_map(map_name)
    st, gl = maps.sample_instance(m, agents, 0, worker)
    r = ReferenceRollout(m["grid"], st[None], gl[None], sd, cfg, device="cpu", mode="act")
    for _ in range(max(warmup, 1)):            # warm-up (also builds the generator, which pins OpenMP to one thread)
        r.step()
    (sync_dir / f"ready_{worker}").touch()
    deadline = time.time() + 300
    while not (sync_dir / "go").exists():
        if time.time() > deadline:
            raise SystemExit(3)
        time.sleep(0.002)
    t0 = time.perf_counter()
    for _ in range(steps):
        r.step()
    print(f"FAIR_SECONDS {time.perf_counter() - t0:.6f}", flush=True)


def time_fair_processes(model: str, map_name: str, agents: int, procs: int, steps: int, timeout_s: float = 240.0, warmup: int = 1):
    """-> (agent-steps/s summed over `procs` concurrent single-thread workers, slowest worker's seconds)."""
    import subprocess
    with tempfile.TemporaryDirectory() as d:
        ps = [subprocess.Popen([sys.executable, str(Path(__file__).resolve()), "--fair-worker", str(i), model, map_name, str(agents),
                                str(steps), d, str(warmup)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, cwd=str(ROOT))
              for i in range(procs)]
        try:
            t_end = time.time() + timeout_s
            while sum((Path(d) / f"ready_{i}").exists() for i in range(procs)) < procs:
                if time.time() > t_end or any(p.poll() not in (None, 0) for p in ps):
                    raise RuntimeError("a fair-baseline worker did not start")
                time.sleep(0.05)
            (Path(d) / "go").touch()
            secs = []
            for p in ps:
                out, _ = p.communicate(timeout=max(1.0, t_end - time.time()))
                secs.append(float([l for l in out.splitlines() if l.startswith("FAIR_SECONDS")][-1].split()[1]))
        finally:
            for p in ps:
                if p.poll() is None:
                    p.kill()
    return procs * agents * steps / max(secs), max(secs)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--fair-worker":
        sys.path.insert(0, str(ROOT))
        _fair_worker_main(sys.argv[2:])
