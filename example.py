#!/usr/bin/env python
"""example.py -- single-episode entry point, shaped like the reference's example.py:14-72
(same flags: --map_name --num_agents --seed --device --max_episode_steps --model --show_map_names).

POGEMA / pogema_toolbox are not used: the episode runs on the B200 engine.  Two modes:
  default     device-resident rollout (mg_engine_rollout): tokenizer, policy, sampling and the soft step on the GPU
  --via-act   the harness loop of pogema_toolbox.run_episode: obs dicts -> MAPFGPTInference.act(obs)
              -> env step, which exercises the drop-in adapter exactly as the reference's example.py does
Weights: weights/MAPF-GPT-<model>.pt when present (reference .pt layout), else seeded random init (stated in the output).
"""
import argparse
import json
import time
from pathlib import Path

import numpy as np


def main():
    p = argparse.ArgumentParser(description="MAPF-GPT inference on the B200 engine")
    p.add_argument("--num_agents", type=int, default=32)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--map_name", type=str, default="validation-random-seed-001")
    p.add_argument("--device", type=str, default=None, help="cuda or cuda:<i> (there is no CPU path)")
    p.add_argument("--max_episode_steps", type=int, default=128)
    p.add_argument("--show_map_names", action="store_true")
    p.add_argument("--model", type=str, choices=["2M", "6M", "85M"], default="2M")
    p.add_argument("--via-act", action="store_true", help="drive the episode through MAPFGPTInference.act(obs dicts)")
    args = p.parse_args()

    from mapf_gpt_b200 import engine as E, maps, weights as W
    if args.show_map_names:
        print("\n".join(maps.map_names()))
        return
    m = maps.load_map(args.map_name)
    grid = m["grid"]
    starts, goals = maps.sample_instance(m, args.num_agents, args.seed)
    path = Path(f"weights/MAPF-GPT-{args.model}.pt")
    if path.exists():
        sd, cfg = W.load_checkpoint(path)
        weights = str(path)
    else:
        cfg = W.model_config(args.model)
        sd = W.random_init(cfg, 1234)
        weights = "seeded random init (no checkpoint under weights/; pretrained weights need the network)"
    dev = 0 if args.device in (None, "cuda") else int(args.device.split(":")[1])
    n = args.num_agents
    t0 = time.perf_counter()
    if args.via_act:
        from mapf_gpt_b200.inference import MAPFGPTInference, MAPFGPTInferenceConfig
        algo = MAPFGPTInference(MAPFGPTInferenceConfig(device=args.device), net=(sd, cfg))
        algo.reset_states()
        env = E.RolloutEngine(1, n, *grid.shape, device=dev)      # plays the POGEMA env of the harness
        env.set_max_episode_steps(args.max_episode_steps)
        env.reset(0, grid, starts, goals)
        pos = starts.copy()
        for _ in range(args.max_episode_steps):
            obs = [{"global_obstacles": grid, "global_xy": tuple(int(v) for v in pos[i]),
                    "global_target_xy": tuple(int(v) for v in goals[i])} for i in range(n)]
            actions = algo.act(obs)
            pos = env.env_step(np.asarray(actions, np.int32)[None])[0]
            if env.metrics()[0, 1] == 1.0:
                break
        met = env.metrics()[0]
    else:
        eng = E.RolloutEngine(1, n, *grid.shape, device=dev)
        eng.load_model(sd, cfg)
        eng.set_seed(args.seed)
        eng.set_max_episode_steps(args.max_episode_steps)
        eng.reset(0, grid, starts, goals)
        eng.rollout(args.max_episode_steps, E.MODE_PHILOX)
        met = eng.metrics()[0]
    dt = time.perf_counter() - t0
    print(json.dumps({"map_name": args.map_name, "num_agents": n, "seed": args.seed, "model": args.model, "weights": weights,
                      "ep_length": int(met[0]), "CSR": met[1], "ISR": met[2], "SoC": met[3], "makespan": met[4],
                      "runtime": dt}))


if __name__ == "__main__":
    main()
