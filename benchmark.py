#!/usr/bin/env python
"""benchmark.py -- sweep shaped like the reference's benchmark.py:20-50 / eval_configs/*/0*.yaml grid_search axes
(num_agents x map_name x seed), run as batched device-resident episodes: every (map, seed) instance of one agent count is
one env slot of the same engine, so a whole row of the reference's result table is ONE rollout.

Without pogema the instances come from mapf_gpt_b200.maps.sample_instance (our seeds, not POGEMA's), and without the
pretrained checkpoints the success rates are those of whatever weights are given; the table columns are the toolbox's
(CSR, ISR, SoC, makespan, ep_length; eval_configs/01-random/01-random.yaml:156-160).
"""
import argparse
import json
from pathlib import Path

import numpy as np

SETS = {  # reference sweeps: (maps, agent counts, horizon) -- eval_configs/0{1..5}-*/0*.yaml
    "01-random": ("validation-random-seed-", [8, 16, 24, 32, 48, 64], 128),
    "02-mazes": ("validation-mazes-seed-", [8, 16, 24, 32, 48, 64], 128),
    "03-warehouse": ("wfi_warehouse", [32, 64, 96, 128, 160, 192], 128),
    "04-movingai": ("Berlin_1_256_", [64, 128, 192, 256], 256),
    "05-puzzles": ("puzzle-", [2, 3, 4], 128),
}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model", choices=["2M", "6M", "85M"], default="2M")
    p.add_argument("--sets", nargs="*", default=list(SETS))
    p.add_argument("--seeds", type=int, default=8, help="instances per map")
    p.add_argument("--device", type=int, default=0)
    args = p.parse_args()
    from mapf_gpt_b200 import engine as E, maps, weights as W
    path = Path(f"weights/MAPF-GPT-{args.model}.pt")
    if path.exists():
        sd, cfg = W.load_checkpoint(path)
    else:
        cfg = W.model_config(args.model)
        sd = W.random_init(cfg, 1234)
        print(f"# no {path}: seeded random-init weights (success rates are not the paper's)")
    for name in args.sets:
        prefix, agent_counts, horizon = SETS[name]
        names = [m for m in maps.map_names() if m.startswith(prefix)]
        by_shape = {}
        for mn in names:
            m = maps.load_map(mn)
            by_shape.setdefault(m["grid"].shape, []).append(m)
        for n in agent_counts:
            rows = []
            for shape, ms in by_shape.items():
                insts = []
                for m in ms:
                    for s in range(args.seeds):
                        try:
                            insts.append((m["grid"],) + maps.sample_instance(m, n, s))
                        except ValueError:
                            pass
                if not insts:
                    continue
                eng = E.RolloutEngine(len(insts), n, *shape, device=args.device)
                eng.load_model(sd, cfg)
                eng.set_max_episode_steps(horizon)
                eng.reset(0, np.stack([g for g, _, _ in insts]), np.stack([s for _, s, _ in insts]),
                          np.stack([g for _, _, g in insts]))
                eng.rollout(horizon, E.MODE_PHILOX)
                rows.append(eng.metrics())
                eng.close()
            if rows:
                met = np.concatenate(rows)
                print(json.dumps({"set": name, "num_agents": n, "episodes": int(met.shape[0]), "CSR": met[:, 1].mean(),
                                  "ISR": met[:, 2].mean(), "SoC": met[:, 3].mean(), "makespan": met[:, 4].mean(),
                                  "ep_length": met[:, 0].mean()}))


if __name__ == "__main__":
    main()
