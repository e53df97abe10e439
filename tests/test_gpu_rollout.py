"""GPU: the whole path (tokens -> forward -> sample -> step) against the CPU oracle, the drop-in
adapter, shard invariance and BASELINE-size properties."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
LOGIT_TOL = 2e-2   # bf16 operands, fp32 accumulate/residual vs the reference fp32; measured max 1.2e-2 (DESIGN.md "Tolerance")


def instances(name, n, envs, seed=0, first=0):
    from mapf_gpt_b200 import maps
    m = maps.load_map(name)
    st = np.stack([maps.sample_instance(m, n, seed, first + e)[0] for e in range(envs)])
    gl = np.stack([maps.sample_instance(m, n, seed, first + e)[1] for e in range(envs)])
    return m["grid"], st, gl


def sharp_model(name="2M"):
    from mapf_gpt_b200 import weights as W
    cfg = W.model_config(name)
    return cfg, W.scale_weights(W.perturb_layernorm(W.random_init(cfg)), 3.0)


def test_teacher_forced_rollout_vs_cpu_oracle(built):
    """Config C1 shape (random-000, 32 agents, 1 env): every step the device state is compared with the
    CPU path (compiled-reference/C tokenizer + torch-fp32 forward + C soft step) and then FORCED to it."""
    from mapf_gpt_b200 import engine as E
    from oracle import cpu_rollout
    cfg, sd = sharp_model()
    grid, st, gl = instances("validation-random-seed-000", 32, 1)
    cpu = cpu_rollout.CpuRollout(grid, st, gl, sd, cfg.n_layer, cfg.n_head)
    eng = E.RolloutEngine(1, 32, *grid.shape)
    eng.load_model(sd, cfg)
    eng.reset(0, grid, st, gl)
    rng = np.random.default_rng(0)
    flips = total = 0
    for t in range(10):
        q67 = torch.from_numpy(rng.exponential(size=(32, 67)).astype(np.float32))
        pos_before = cpu.pos.copy()
        last = cpu.last.copy()
        toks, ref_acts = cpu.step(q=q67)
        eng.update_agents(pos_before, None, last)
        got = eng.generate_observations()
        assert (got[0].astype(np.int64) == toks[0]).all(), f"tokens differ at step {t}"
        acts, logits = eng.act(E.MODE_SUPPLIED_Q, q67[:, :5].numpy()[None], want_logits=True)
        from oracle import gpt_oracle as G
        ref_logits = G.forward_logits(sd, cfg.n_layer, cfg.n_head, torch.from_numpy(toks[0]))[:, :5].numpy()
        assert np.abs(logits[0] - ref_logits).max() < LOGIT_TOL
        p = torch.softmax(torch.from_numpy(ref_logits), -1).numpy() / q67[:, :5].numpy()
        srt = np.sort(p, -1)
        dec = (srt[:, -1] - srt[:, -2]) > 0.15 * srt[:, -1]
        assert (acts[0][dec] == ref_acts[0][dec]).all()
        flips += int((acts[0] != ref_acts[0]).sum()); total += 32
        new = eng.env_step(ref_acts)                       # execute the ORACLE's actions (teacher forcing)
        assert (new[0] == cpu.pos[0]).all(), f"positions differ at step {t}"
    assert flips <= 0.03 * total                            # flip rate inside the tolerance band
    eng.close()


def test_device_rollout_equals_stepwise_api_and_act_host(built):
    from mapf_gpt_b200 import engine as E
    cfg, sd = sharp_model()
    grid, st, gl = instances("validation-mazes-seed-000", 24, 6)
    outs = []
    for mode in ("rollout", "stepwise", "host"):
        eng = E.RolloutEngine(6, 24, *grid.shape)
        eng.load_model(sd, cfg)
        eng.set_seed(3)
        eng.reset(0, grid, st, gl)
        if mode == "rollout":
            eng.rollout(5, E.MODE_PHILOX)
        elif mode == "stepwise":
            for _ in range(5):
                eng.update_agents()
                eng.generate_observations(fetch=False)
                eng.act(E.MODE_PHILOX)
                eng.env_step(None, fetch=False)
        else:
            pos = st.copy()
            for _ in range(5):
                eng.act_host(pos, gl, E.MODE_PHILOX)
                pos = eng.env_step(None)
        outs.append((eng.positions().copy(), eng.tokens().copy(), eng.metrics().copy()))
        eng.close()
    for o in outs[1:]:
        assert (o[0] == outs[0][0]).all() and (o[1] == outs[0][1]).all() and np.array_equal(o[2], outs[0][2])


def test_env_shard_invariance(built):
    """Env e behaves the same whichever engine (GPU) hosts it: Philox streams follow the global env id."""
    from mapf_gpt_b200 import engine as E
    cfg, sd = sharp_model()
    grid, st, gl = instances("validation-mazes-seed-000", 16, 4)
    whole = E.RolloutEngine(4, 16, *grid.shape)
    whole.load_model(sd, cfg); whole.set_seed(9); whole.reset(0, grid, st, gl); whole.rollout(6, E.MODE_PHILOX)
    ref = whole.positions().copy()
    whole.close()
    for first in (0, 2):
        part = E.RolloutEngine(2, 16, *grid.shape)
        part.load_model(sd, cfg); part.set_seed(9); part.set_env_offset(first)
        part.reset(0, grid, st[first:first + 2], gl[first:first + 2]); part.rollout(6, E.MODE_PHILOX)
        assert (part.positions() == ref[first:first + 2]).all()
        part.close()


def test_ragged_slots_and_edge_cases(built):
    import oracle
    from mapf_gpt_b200 import engine as E, maps
    cfg, sd = sharp_model()
    m = maps.load_map("validation-random-seed-002")
    grid = m["grid"]
    eng = E.RolloutEngine(3, 12, *grid.shape)
    eng.load_model(sd, cfg)
    ns = [12, 1, 5]
    sts, gls, orcs = [], [], []
    for e, n in enumerate(ns):
        st, gl = maps.sample_instance(m, n, 7, e)
        if e == 2:
            gl[0] = st[0]                                    # start == goal
        eng.reset(e, grid, st, gl)
        o = oracle.ObsOracle(grid); o.create_agents(st, gl)
        sts.append(st); gls.append(gl); orcs.append(o)
    assert eng.num_envs == 3
    eng.update_agents()
    toks = eng.generate_observations()
    for e, n in enumerate(ns):
        orcs[e].update_agents(sts[e], gls[e], np.full(n, -1, np.int32))
        assert (toks[e, :n] == orcs[e].generate_observations()).all()
    acts = eng.act(E.MODE_GREEDY)
    assert acts.shape == (3, 12)
    assert all(((acts[e, :n] >= 0) & (acts[e, :n] <= 4)).all() and (acts[e, n:] == -1).all() for e, n in enumerate(ns))
    eng.env_step(None)
    met = eng.metrics()
    assert met[:, 7].tolist() == [12, 1, 5] and met[2, 5] >= 0
    from mapf_gpt_b200 import _lib
    with pytest.raises(_lib.MgError):
        eng.reset(0, grid, np.array([[[0, 0]]], np.int32), np.array([[[6, 6]]], np.int32))   # outside the padding contract
    with pytest.raises(_lib.MgError):
        E.RolloutEngine(1, 4, 600, 600)                      # beyond the engine's grid capacity
    eng.close()


def test_dropin_adapter_matches_reference_call_pattern(built, tmp_path):
    """MAPFGPTInference.act on POGEMA-style observation dicts: tokens exact vs the oracle, actions equal
    argmax(softmax(logits)/q) with q drawn like torch.multinomial does (chunks of batch_size x 67 from a
    CUDA generator seeded 0 at reset_states)."""
    import oracle
    from mapf_gpt_b200 import weights as W
    from mapf_gpt_b200.inference import MAPFGPTInference, MAPFGPTInferenceConfig
    cfg, sd = sharp_model()
    path = tmp_path / "rand-2M.pt"
    W.save_checkpoint(path, sd, cfg)
    grid, st, gl = instances("validation-random-seed-000", 32, 1)
    algo = MAPFGPTInference(MAPFGPTInferenceConfig(path_to_weights=str(path), device="cuda", batch_size=20))
    algo.reset_states()
    o = oracle.ObsOracle(grid); o.create_agents(st[0], gl[0])
    gen = torch.Generator(device="cuda").manual_seed(0)
    from mapf_gpt_b200 import engine as E
    side = E.RolloutEngine(1, 1, 11, 11)
    side.load_model(sd, cfg)
    pos, last = st[0].copy(), np.full(32, -1, np.int32)
    for t in range(4):
        obs = [{"global_obstacles": grid, "global_xy": tuple(int(v) for v in pos[i]),
                "global_target_xy": tuple(int(v) for v in gl[0][i])} for i in range(32)]
        acts = algo.act(obs)
        o.update_agents(pos, gl[0], last)
        assert (algo._engine.tokens()[0] == o.generate_observations()).all()
        q = torch.cat([torch.empty((b, 67), device="cuda").exponential_(1, generator=gen)[:, :5] for b in (20, 12)])
        lg = torch.from_numpy(side.forward_tokens(algo._engine.tokens()[0])).cuda()      # rows are batch-invariant
        want = (torch.softmax(lg, -1) / q).argmax(-1).cpu().numpy()
        assert (np.asarray(acts) == want).all()
        last = np.asarray(acts, np.int32)
        pos, _ = oracle.pogema_step_soft(grid, pos, last)
    # pre-tokenized rows are accepted too (inference.py:146)
    rows = o.generate_observations().tolist()
    assert len(algo.act(rows)) == 32
    algo.reset_states()
    assert algo._engine is not None and algo._engine.num_envs == 0   # O(1) reset: the engine and its model stay


@pytest.mark.parametrize("name,n,envs,model", [("validation-mazes-seed-000", 64, 1024, "2M"), ("wfi_warehouse", 192, 64, "6M"),
                                               ("Berlin_1_256_05", 256, 8, "85M")])
def test_baseline_size_properties(built, name, n, envs, model):
    """BASELINE.json sizes (C2 full; C3/C4 at reduced env counts): size-independent invariants plus exact
    token parity on sampled envs."""
    import oracle
    from mapf_gpt_b200 import engine as E
    cfg, sd = sharp_model(model)
    grid, st, gl = instances(name, n, envs)
    eng = E.RolloutEngine(envs, n, *grid.shape)
    eng.load_model(sd, cfg)
    eng.reset(0, grid, st, gl)
    eng.rollout(3, E.MODE_PHILOX)
    pos = eng.positions()
    assert (grid[pos[..., 0], pos[..., 1]] == 0).all()                               # nobody inside an obstacle
    flat = pos[..., 0].astype(np.int64) * 1000 + pos[..., 1]
    assert all(len(np.unique(flat[e])) == n for e in range(envs))                    # no vertex conflicts
    assert (np.abs(pos - st).sum(-1) <= 3).all()                                     # at most one cell per step
    tok = eng.tokens().astype(np.int32)
    assert tok.min() >= 0 and tok.max() <= 66 and (tok[..., 251:] == 66).all()
    assert (tok[..., 121] == 20).all() and (tok[..., 122] == 20).all() and (tok[..., 60] == 20).all()
    met = eng.metrics()
    assert (met[:, 0] == 3).all() and met[:, 6].sum() == 3 * n * envs
    # the tokens on the device are those of the LAST observe (before the last move): rebuild them on the CPU
    eng2 = E.RolloutEngine(envs, n, *grid.shape)
    eng2.load_model(sd, cfg); eng2.reset(0, grid, st, gl); eng2.rollout(2, E.MODE_PHILOX)
    pos2 = eng2.positions()
    for e in np.linspace(0, envs - 1, 4).astype(int):
        o = oracle.ObsOracle(grid); o.create_agents(st[e], gl[e])
        hist = tok[e, :, 125:130]                               # own slot: a1..a5 as the device holds them
        # replay: positions before the third observe are pos2; history comes from the device's own actions
        o2 = oracle.ObsOracle(grid); o2.create_agents(pos2[e], gl[e])
        o2.update_agents(pos2[e], gl[e], np.full(n, -1, np.int32))
        ref = o2.generate_observations()
        a = tok[e].copy(); b = ref.copy()
        for s in range(13):                                      # mask the action-history tokens (replay has none)
            a[:, 125 + 10 * s:130 + 10 * s] = 0; b[:, 125 + 10 * s:130 + 10 * s] = 0
        assert (a == b).all()
    eng.close(); eng2.close()


def test_reference_shaped_entry_points(built):
    """example.py (reference flags) in both modes: device-resident rollout and the run_episode-style act() loop."""
    import json, subprocess, sys
    root = Path(__file__).resolve().parents[1]
    for extra in ([], ["--via-act"]):
        out = subprocess.run([sys.executable, str(root / "example.py"), "--map_name", "validation-random-seed-000",
                              "--num_agents", "8", "--max_episode_steps", "6", "--seed", "1"] + extra,
                             capture_output=True, text=True, timeout=600, cwd=root)
        assert out.returncode == 0, out.stderr[-2000:]
        res = json.loads(out.stdout.strip().splitlines()[-1])
        assert res["num_agents"] == 8 and 1 <= res["ep_length"] <= 6 and 0.0 <= res["ISR"] <= 1.0
    names = subprocess.run([sys.executable, str(root / "example.py"), "--show_map_names"], capture_output=True, text=True,
                           timeout=120, cwd=root).stdout.split()
    assert "validation-mazes-seed-000" in names and "wfi_warehouse" in names


def test_act_batch_multi_slot_keys_and_partial_calls(built):
    """act_batch(observations_list, positions): persistent slot keys, ragged agent counts, a slot missing from a call keeps
    its history (inference.py:151-172)."""
    import oracle
    from mapf_gpt_b200 import maps
    from mapf_gpt_b200.inference import MAPFGPTInference, MAPFGPTInferenceConfig
    cfg, sd = sharp_model()
    m = maps.load_map("validation-random-seed-003")
    grid = m["grid"]
    ns = {"a": 9, "b": 5, 7: 12}
    inst = {k: maps.sample_instance(m, n, 3, i) for i, (k, n) in enumerate(ns.items())}
    algo = MAPFGPTInference(MAPFGPTInferenceConfig(device="cuda"), net=(sd, cfg), do_sample=False)
    algo.reset_states()
    orc, pos, last = {}, {}, {}
    for k, (st, gl) in inst.items():
        orc[k] = oracle.ObsOracle(grid); orc[k].create_agents(st, gl)
        pos[k], last[k] = st.copy(), np.full(len(st), -1, np.int32)

    def obs_of(k):
        return [{"global_obstacles": grid, "global_xy": tuple(int(v) for v in pos[k][i]),
                 "global_target_xy": tuple(int(v) for v in inst[k][1][i])} for i in range(len(pos[k]))]

    for step, keys in enumerate([["a", "b", 7], ["a", 7], ["b", "a", 7], ["b"]]):
        res = algo.act_batch([obs_of(k) for k in keys], positions=keys)
        toks = algo._engine.tokens()
        for k, acts in zip(keys, res):
            assert len(acts) == ns[k] and all(0 <= a <= 4 for a in acts)
            orc[k].update_agents(pos[k], inst[k][1], last[k])
            e = algo._slots[k]
            assert (toks[e, :ns[k]] == orc[k].generate_observations()).all(), (step, k)
            last[k] = np.asarray(acts, np.int32)
            pos[k], _ = oracle.pogema_step_soft(grid, pos[k], last[k])
    assert algo.act_batch([]) == []
