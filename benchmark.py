#!/usr/bin/env python
"""benchmark.py -- the reference's benchmark sweep (benchmark.py:20-50) on the B200 engine, driven by the same YAML files.

For every set folder (01-random ... 05-puzzles) it reads `<eval_configs>/<set>/<set>.yaml` (and `<set>/maps.yaml` when the
directory holds one; otherwise the packaged map store), expands the `environment` grid_search axes
(num_agents x map_name x seed; eval_configs/01-random/01-random.yaml:1-14), runs every algorithm of the `algorithms`
section, writes `<out>/<set>/<algorithm>.json` with one record per episode in the toolbox's shape
({"metrics": {...}, "env_grid_search": {...}, "algorithm": ...}) and renders the tabular `results_views`
(`:150-160`: mean of every metric over the dropped keys).

What differs from pogema_toolbox.evaluation (not installable here): all episodes of one agent count are env slots of ONE
engine and run as a device-resident rollout instead of Dask workers calling act() per step; under torchrun the episodes
are sharded over the ranks (mapf_gpt_b200.parallel) and gathered with one all-reduce.  Starts/goals come from
maps.sample_instance (our sampler and seeds, not POGEMA's), `runtime` is the batch's wall time divided by its episodes,
and without pretrained checkpoints under weights/ the policy is a seeded random init (stated in the output).

  python benchmark.py                                   # all five sets, the algorithms the YAMLs name (2M, 6M)
  python benchmark.py --eval_configs /path/to/eval_configs --sets 02-mazes --algorithms MAPF-GPT-2M
  python benchmark.py --add_85M                         # also MAPF-GPT-85M (weights/MAPF-GPT-85M.pt)
  torchrun --nproc-per-node 8 benchmark.py              # episodes sharded over 8 GPUs
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import re
import time
from pathlib import Path

import numpy as np
import yaml

FOLDERS = ["01-random", "02-mazes", "03-warehouse", "04-movingai", "05-puzzles"]      # benchmark.py:28-34
METRICS = ["CSR", "ISR", "SoC", "makespan", "ep_length", "avg_agents_density", "runtime"]
MAX_DISTINCT_LARGE_MAPS = 64      # engine capacity for per-map precompute tables (maps wider than 64 cells only)


def expand_grid_search(env_cfg: dict):
    """-> (fixed keys, [axis names], [one dict per combination]) of the `environment` section."""
    axes = {k: v["grid_search"] for k, v in env_cfg.items() if isinstance(v, dict) and "grid_search" in v}
    fixed = {k: v for k, v in env_cfg.items() if k not in axes}
    names = list(axes)
    combos = [dict(zip(names, c)) for c in itertools.product(*(axes[k] for k in names))]
    return fixed, names, combos


def load_policy(algo_name: str, algo_cfg: dict, cache: dict):
    from mapf_gpt_b200 import weights as W
    path = Path(algo_cfg.get("path_to_weights", f"weights/{algo_name}.pt"))
    if path in cache:
        return cache[path]
    if path.exists():
        sd, cfg = W.load_checkpoint(path)
        src = str(path)
    else:
        m = re.search(r"(2M|6M|85M)", path.name + algo_name)
        if not m:
            raise FileNotFoundError(f"{path} not found and the model size cannot be inferred from {algo_name!r}")
        cfg = W.model_config(m.group(1))
        sd = W.random_init(cfg, 1234)
        src = f"seeded random init of the {m.group(1)} architecture ({path} not found; pretrained weights need the network)"
    cache[path] = (sd, cfg, src)
    return cache[path]


def run_batch(episodes, sd, cfg, horizon, device, seed0):
    """episodes: [(combo dict, map dict, starts, goals)] with one agent count -> metrics rows [len, MG_METRIC_COLS], seconds."""
    from mapf_gpt_b200 import engine as E
    n = len(episodes[0][2])
    H = max(ep[1]["grid"].shape[0] for ep in episodes)
    Wd = max(ep[1]["grid"].shape[1] for ep in episodes)
    grids = np.ones((len(episodes), H, Wd), np.uint8)              # smaller maps: pad with obstacles, origin unchanged
    for i, ep in enumerate(episodes):
        g = ep[1]["grid"]
        grids[i, :g.shape[0], :g.shape[1]] = g
    t0 = time.perf_counter()
    eng = E.RolloutEngine(len(episodes), n, H, Wd, device=device)
    eng.load_model(sd, cfg)
    eng.set_seed(seed0)
    eng.set_max_episode_steps(horizon)
    eng.reset(0, grids, np.stack([ep[2] for ep in episodes]), np.stack([ep[3] for ep in episodes]))
    eng.rollout(horizon, E.MODE_PHILOX)
    met = eng.metrics()
    eng.close()
    return met, time.perf_counter() - t0


def tabular_view(records, view: dict, axis_names):
    """pogema_toolbox tabular view: mean of the metrics over `drop_keys`, grouped by what is left."""
    drop = set(view.get("drop_keys", []))
    keys = [k for k in axis_names if k not in drop]
    cols = [m for m in METRICS if m not in drop]
    groups = {}
    for r in records:
        gk = tuple([r["algorithm"]] + [r["env_grid_search"][k] for k in keys])
        groups.setdefault(gk, []).append(r["metrics"])
    rows = []
    for gk in sorted(groups, key=lambda t: tuple(str(x) if not isinstance(x, (int, float)) else x for x in t)):
        ms = groups[gk]
        rows.append(list(gk) + [float(np.mean([m[c] for m in ms])) for c in cols] + [len(ms)])
    return ["algorithm"] + keys + cols + ["episodes"], rows


def print_table(header, rows, title):
    try:
        from tabulate import tabulate
        print(f"\n{title}\n" + tabulate(rows, headers=header, floatfmt=".3f"))
    except Exception:
        print(f"\n{title}\n" + " | ".join(header))
        for r in rows:
            print(" | ".join(f"{v:.3f}" if isinstance(v, float) else str(v) for v in r))


def main():
    from mapf_gpt_b200 import maps, parallel
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--eval_configs", default=None, help="reference-layout directory; default ./eval_configs, else the packaged copy")
    p.add_argument("--sets", nargs="*", default=FOLDERS)
    p.add_argument("--algorithms", nargs="*", default=None, help="subset of the YAML's algorithm names")
    p.add_argument("--add_85M", action="store_true", help="also run MAPF-GPT-85M (the YAMLs list 2M and 6M)")
    p.add_argument("--out", default="eval_results", help="results go to <out>/<set>/<algorithm>.json")
    p.add_argument("--limit", type=int, default=None, help="smoke runs: at most this many (map, seed) episodes per agent count")
    p.add_argument("--device", type=int, default=None)
    args = p.parse_args()

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = args.device if args.device is not None else int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{device}"))
    base = Path(args.eval_configs) if args.eval_configs else (Path("eval_configs") if Path("eval_configs").is_dir() else maps.EVAL_CONFIGS)
    policies = {}
    summary = []
    for folder in args.sets:
        cfg_path = base / folder / f"{Path(folder).name}.yaml"
        with open(cfg_path) as f:
            ecfg = yaml.safe_load(f)
        if (base / folder / "maps.yaml").exists():                       # benchmark.py:36-40
            with open(base / folder / "maps.yaml") as f:
                maps.register_maps(yaml.safe_load(f), folder)
        fixed, axis_names, combos = expand_grid_search(ecfg["environment"])
        if fixed.get("collision_system", "soft") != "soft" or fixed.get("on_target", "nothing") != "nothing":
            raise SystemExit(f"{cfg_path}: only collision_system=soft / on_target=nothing episodes are implemented (the eval sets use these)")
        horizon = int(fixed.get("max_episode_steps", 128))
        algos = dict(ecfg["algorithms"])
        if args.add_85M:
            algos["MAPF-GPT-85M"] = {"name": "MAPF-GPT", "path_to_weights": "weights/MAPF-GPT-85M.pt"}
        if args.algorithms:
            algos = {k: v for k, v in algos.items() if k in args.algorithms}
        by_n = {}
        for c in combos:
            full = {**fixed, **c}
            by_n.setdefault(int(full["num_agents"]), []).append(c)
        map_cache = {}
        for algo_name, algo_cfg in algos.items():
            sd, gcfg, src = load_policy(algo_name, algo_cfg, policies)
            records = []
            t_set = time.perf_counter()
            for n, cs in sorted(by_n.items()):
                if args.limit:
                    cs = cs[:args.limit]
                episodes = []
                for c in cs:
                    full = {**fixed, **c}
                    name = full["map_name"]
                    if name not in map_cache:
                        map_cache[name] = maps.load_map(name)
                    try:
                        st, gl = maps.sample_instance(map_cache[name], n, int(full.get("seed", 0)))
                    except ValueError as ex:                                   # more agents than free cells
                        if rank == 0:
                            print(f"# skipped {c}: {ex}")
                        continue
                    episodes.append((c, map_cache[name], st, gl))
                if not episodes:
                    continue
                first, cnt = parallel.shard_range(len(episodes), rank, world)
                mine = episodes[first:first + cnt]
                rows = np.zeros((0, 11))
                large = max(max(ep[1]["grid"].shape) for ep in episodes) > 74
                i = 0
                while i < len(mine):          # large maps: at most MAX_DISTINCT_LARGE_MAPS distinct grids per engine
                    j, seen = i, set()
                    while j < len(mine) and (not large or len(seen | {mine[j][1]["name"]}) <= MAX_DISTINCT_LARGE_MAPS):
                        seen.add(mine[j][1]["name"])
                        j += 1
                    met, secs = run_batch(mine[i:j], sd, gcfg, horizon, device, int(fixed["seed"]) if isinstance(fixed.get("seed"), int) else 0)
                    rows = np.concatenate([rows, np.concatenate([met, np.full((len(met), 1), secs / len(met))], 1)])
                    i = j
                table = parallel.gather_rows(rows, len(episodes), first, device=f"cuda:{device}" if world > 1 else None)
                for ep, r in zip(episodes, table):
                    records.append({"metrics": {"CSR": r[1], "ISR": r[2], "SoC": r[3], "makespan": r[4], "ep_length": r[0],
                                                "avg_agents_density": r[8], "runtime": r[10]},
                                    "env_grid_search": {k: ep[0][k] for k in axis_names}, "algorithm": algo_name})
            if rank == 0:
                out = Path(args.out) / folder
                out.mkdir(parents=True, exist_ok=True)
                with open(out / f"{algo_name}.json", "w") as f:
                    json.dump(records, f)
                secs = time.perf_counter() - t_set
                steps = sum(r["metrics"]["ep_length"] * r["env_grid_search"].get("num_agents", fixed.get("num_agents", 0)) for r in records)
                print(f"# {folder} / {algo_name}: {len(records)} episodes in {secs:.1f} s ({steps / max(secs, 1e-9):,.0f} agent-steps/s incl. "
                      f"engine setup) on {world} GPU(s); weights: {src}")
                for vname, view in (ecfg.get("results_views") or {}).items():
                    if view.get("type") == "tabular":
                        header, trows = tabular_view(records, view, axis_names)
                        if view.get("print_results", True):
                            print_table(header, trows, f"{folder} / {vname}")
                summary.append({"set": folder, "algorithm": algo_name, "episodes": len(records), "seconds": secs})
    if rank == 0:
        print(json.dumps({"benchmark": summary, "total_episodes": sum(s["episodes"] for s in summary)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
