"""GPU: each sm_100a kernel against its checker, through the C ABI."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD_OBS = np.load(Path(__file__).parent / "golden" / "obs_golden.npz")
GOLD_GPT = np.load(Path(__file__).parent / "golden" / "gpt_golden.npz")
GOLDEN = Path(__file__).parent / "golden"
N_SCEN = len([k for k in GOLD_OBS.files if k.endswith("_tokens")])
LOGIT_TOL = 2e-2   # bf16 operands, fp32 accumulate/residual vs the reference fp32; measured max 1.2e-2 (DESIGN.md "Tolerance")


@pytest.fixture(scope="module")
def dev(built):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("cfg,M,N,K", [(0, 256, 160, 160), (0, 384, 480, 160), (0, 128, 160, 640), (1, 256, 768, 256),
                                       (1, 128, 256, 1024), (1, 256, 2304, 768), (2, 256, 384, 192), (0, 128 * 151, 640, 160)])
def test_tcgen05_gemm(dev, cfg, M, N, K):
    from mapf_gpt_b200 import engine as E
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    A = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev, generator=g) * 0.5).bfloat16()
    C = E.test_gemm(A, B, cfg)
    ref = A.double() @ B.double().t()
    assert float((C.double() - ref).abs().max()) < 1e-3 * float(ref.abs().max()) + 1e-4   # fp32 accumulation order only


@pytest.mark.parametrize("hs,n_seq,n_head,scale", [(32, 2, 5, 1.0), (32, 3, 8, 0.2), (64, 2, 12, 1.0), (32, 1, 1, 3.0)])
def test_tcgen05_attention(dev, hs, n_seq, n_head, scale):
    from mapf_gpt_b200 import engine as E
    g = torch.Generator(device=dev).manual_seed(hs + n_head)
    q = (torch.randn(n_seq, n_head, 256, hs, device=dev, generator=g) * scale).bfloat16()
    k = (torch.randn(n_seq, n_head, 256, hs, device=dev, generator=g) * scale).bfloat16()
    v = torch.randn(n_seq, n_head, 256, hs, device=dev, generator=g).bfloat16()
    o = E.test_attention(q, k, v)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())   # non-causal, model.py:58-60
    assert float((o.float() - ref).abs().max()) < 2e-2      # P and O are rounded to bf16


@pytest.mark.parametrize("k", range(N_SCEN))
def test_engine_verbs_vs_golden_tokens(dev, k):
    """update_agents + generate_observations (and the reset-time BFS) against tokens from the compiled reference."""
    import oracle
    from mapf_gpt_b200 import engine as E
    grid, goals = GOLD_OBS[f"s{k}_grid"], GOLD_OBS[f"s{k}_goals"]
    pos, act, tok = GOLD_OBS[f"s{k}_pos"], GOLD_OBS[f"s{k}_act"], GOLD_OBS[f"s{k}_tokens"]
    n = pos.shape[1]
    eng = E.RolloutEngine(2, n, *grid.shape)
    eng.reset(0, grid, np.stack([pos[0], pos[0]]), np.stack([goals, goals]))
    o = oracle.ObsOracle(grid)
    o.create_agents(pos[0], goals)
    for a in range(0, n, max(1, n // 5)):                                  # BFS field, bit-exact
        assert (eng.cost2go(1, a) == o.partial(a)[1]).all()
    for t in range(len(pos)):
        eng.update_agents(np.stack([pos[t]] * 2), None, np.stack([act[t]] * 2))
        got = eng.generate_observations()
        assert (got[0] == tok[t]).all() and (got[1] == tok[t]).all(), f"scenario {k} step {t}"
    eng.close()


@pytest.mark.parametrize("k", [0, 4])
def test_single_env_c_abi_twin_of_pybind_module(dev, k, built):
    """mg_gen_* == ObservationGenerator verbs (observation_generator.cpp:548-563)."""
    import ctypes as C
    grid, goals = GOLD_OBS[f"s{k}_grid"], GOLD_OBS[f"s{k}_goals"]
    pos, act, tok = GOLD_OBS[f"s{k}_pos"], GOLD_OBS[f"s{k}_act"], GOLD_OBS[f"s{k}_tokens"]
    n = pos.shape[1]
    g32 = np.ascontiguousarray(grid, np.int32)
    p = lambda a: np.ascontiguousarray(a, np.int32).ctypes.data_as(C.c_void_p)
    h = built.mg_gen_create(p(g32), grid.shape[0], grid.shape[1], None)
    assert h
    gl = np.ascontiguousarray(goals, np.int32)
    assert built.mg_gen_create_agents(h, p(pos[0]), p(gl), n) == 0
    out = np.empty((n, 256), np.int32)
    for t in range(len(pos)):
        assert built.mg_gen_update_agents(h, p(pos[t]), p(gl), p(act[t]), n) == 0
        assert built.mg_gen_generate_observations(h, out.ctypes.data_as(C.c_void_p)) == 0
        assert (out == tok[t]).all()
    built.mg_gen_destroy(h)


def test_goal_change_recomputes_field(dev):
    import oracle
    from mapf_gpt_b200 import engine as E, maps
    m = maps.load_map("validation-mazes-seed-001")
    grid = m["grid"]
    st, gl = maps.sample_instance(m, 20, 3)
    _, gl2 = maps.sample_instance(m, 20, 4)
    eng = E.RolloutEngine(1, 20, *grid.shape)
    eng.reset(0, grid, st, gl)
    o = oracle.ObsOracle(grid)
    o.create_agents(st, gl)
    act = np.full(20, -1, np.int32)
    for goals in (gl, gl2, gl2, gl):
        eng.update_agents(st[None], goals[None], act[None])
        o.update_agents(st, goals, act)
        assert (eng.generate_observations()[0] == o.generate_observations()).all()
        act = np.arange(20, dtype=np.int32) % 5
    eng.close()


@pytest.mark.parametrize("density,seed", [(0.2, 0), (0.6, 1), (0.95, 2)])
def test_soft_step_kernel_vs_oracle(dev, density, seed):
    import oracle
    from mapf_gpt_b200 import engine as E
    rng = np.random.default_rng(seed)
    H, Wd, envs = 24, 27, 48
    grid = np.ones((H, Wd), np.uint8)
    grid[5:-5, 5:-5] = rng.random((H - 10, Wd - 10)) < 0.15
    free = np.argwhere(grid == 0)
    n = max(2, int(density * len(free)))
    pos = np.stack([free[rng.permutation(len(free))[:n]] for _ in range(envs)]).astype(np.int32)
    eng = E.RolloutEngine(envs, n, H, Wd)
    eng.reset(0, grid, pos, pos[:, ::-1].copy())
    for t in range(6):
        act = rng.integers(-1, 6, (envs, n)).astype(np.int32)
        new = eng.env_step(act)
        for e in range(envs):
            pos[e], _ = oracle.pogema_step_soft(grid, pos[e], act[e])
        assert (new == pos).all(), f"step {t}"
    met = eng.metrics()
    assert (met[:, 0] == 6).all() and (met[:, 6] == 6 * n).all()
    eng.close()


@pytest.mark.parametrize("name,tag", [("2M", "init"), ("2M", "sharp"), ("6M", "init"), ("6M", "sharp"), ("85M", "sharp")])
def test_forward_logits_vs_reference_golden(dev, name, tag):
    """Logits of the reference model.py (fp32, CPU) on committed tokens; tolerance LOGIT_TOL abs."""
    from mapf_gpt_b200 import engine as E, weights as W
    cfg = W.model_config(name)
    sd = W.random_init(cfg, 1234)
    if tag == "sharp":
        sd = W.scale_weights(W.perturb_layernorm(sd), 3.0)
    assert W.state_dict_digest(sd) == bytes(GOLD_GPT[f"{name}_{tag}_digest"]).hex()
    eng = E.RolloutEngine(1, 1, 11, 11)
    eng.load_model(sd, cfg)
    lg = eng.forward_tokens(GOLD_GPT["tokens"])
    ref = GOLD_GPT[f"{name}_{tag}_logits"][:, :5]
    err = np.abs(lg - ref).max()
    assert err < LOGIT_TOL, err
    # greedy actions agree wherever the reference's top-2 margin exceeds twice the tolerance
    srt = np.sort(ref, -1)
    dec = (srt[:, -1] - srt[:, -2]) > 2 * LOGIT_TOL
    assert (lg.argmax(-1)[dec] == GOLD_GPT[f"{name}_{tag}_greedy"][dec]).all()
    eng.close()


def test_forward_many_rows_matches_oracle_and_is_batch_invariant(dev):
    from mapf_gpt_b200 import engine as E, weights as W
    from oracle import gpt_oracle as G
    cfg = W.model_config("2M")
    sd = W.scale_weights(W.perturb_layernorm(W.random_init(cfg)), 3.0)
    rng = np.random.default_rng(1)
    toks = rng.integers(0, 67, (300, 256)).astype(np.int8)
    eng = E.RolloutEngine(1, 1, 11, 11)
    eng.load_model(sd, cfg)
    a = eng.forward_tokens(toks)
    b = eng.forward_tokens(toks[:7])
    assert np.array_equal(a[:7], b)                                     # rows are independent, bitwise
    sdd = {k: v.to(dev) for k, v in sd.items()}
    ref = G.forward_logits(sdd, cfg.n_layer, cfg.n_head, torch.from_numpy(toks.astype(np.int64)).to(dev))[:, :5].cpu().numpy()
    assert np.abs(a - ref).max() < LOGIT_TOL
    eng.close()


def test_sampler_identical_p_and_q_give_identical_actions(dev):
    """GPT.act tail (model.py:249-257) == argmax(softmax(logits[:5]) / q)."""
    from mapf_gpt_b200 import engine as E, maps, weights as W
    m = maps.load_map("validation-random-seed-000")
    n, envs = 32, 8
    st = np.stack([maps.sample_instance(m, n, 2, e)[0] for e in range(envs)])
    gl = np.stack([maps.sample_instance(m, n, 2, e)[1] for e in range(envs)])
    cfg = W.model_config("2M")
    eng = E.RolloutEngine(envs, n, *m["grid"].shape)
    eng.load_model(W.scale_weights(W.random_init(cfg), 4.0), cfg)
    eng.reset(0, m["grid"], st, gl)
    eng.update_agents()
    eng.generate_observations(fetch=False)
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.empty((envs * n, 67), device=dev).exponential_(1, generator=g)[:, :5]
    acts, logits = eng.act(E.MODE_SUPPLIED_Q, q.cpu().numpy().reshape(envs, n, 5), want_logits=True)
    p = torch.softmax(torch.from_numpy(logits).reshape(-1, 5).to(dev), -1)
    want = (p / q).argmax(-1).cpu().numpy().reshape(envs, n)
    assert (acts == want).all()
    greedy = eng.act(E.MODE_GREEDY)
    assert (greedy == logits.argmax(-1)).all()
    # philox mode: deterministic in (seed, env, agent, step), different across seeds
    eng.set_seed(5); a1 = eng.act(E.MODE_PHILOX)
    eng.set_seed(5); a2 = eng.act(E.MODE_PHILOX)
    eng.set_seed(6); a3 = eng.act(E.MODE_PHILOX)
    assert (a1 == a2).all() and (a1 != a3).any()
    eng.close()


@pytest.mark.parametrize("name", ["2M", "6M"])
def test_fused_and_generic_paths_agree(dev, name, monkeypatch):
    """C in {160, 256}: post_attn_kernel (fused proj+LN2+MLP+next LN1) vs the five separate kernels."""
    from mapf_gpt_b200 import engine as E, weights as W
    cfg = W.model_config(name)
    sd = W.scale_weights(W.perturb_layernorm(W.random_init(cfg)), 3.0)
    toks = np.random.default_rng(3).integers(0, 67, (70, 256)).astype(np.int8)
    outs = []
    for generic in ("0", "1"):
        monkeypatch.setenv("MAPF_GPT_B200_GENERIC", generic)
        eng = E.RolloutEngine(1, 1, 11, 11)
        eng.load_model(sd, cfg)
        outs.append(eng.forward_tokens(toks))
        eng.close()
    assert np.abs(outs[0] - outs[1]).max() < 2e-2       # same bf16 operands; GELU polynomial vs erff, LN pass order
    from oracle import gpt_oracle as G
    ref = G.forward_logits(sd, cfg.n_layer, cfg.n_head, torch.from_numpy(toks.astype(np.int64)))[:, :5].numpy()
    assert np.abs(outs[0] - ref).max() < LOGIT_TOL and np.abs(outs[1] - ref).max() < LOGIT_TOL


@pytest.mark.parametrize("name", ["2M", "6M"])
def test_block0_lookup_table_is_bit_identical_to_the_computed_block0(dev, name, monkeypatch):
    """Block 0's embedding + ln_1 + c_attn depend only on (token, position): the engine tabulates them at model load through
    the same kernels; logits must not change by a single bit, and every (token, position) pair is exercised."""
    from mapf_gpt_b200 import engine as E, weights as W
    cfg = W.model_config(name)
    sd = W.scale_weights(W.perturb_layernorm(W.random_init(cfg)), 3.0)
    rng = np.random.default_rng(5)
    toks = np.concatenate([np.repeat(np.arange(67, dtype=np.int8)[:, None], 256, 1),      # every (token, position) pair
                           rng.integers(0, 67, (61, 256)).astype(np.int8)])
    outs = []
    # default: x is gathered from the table by the first post_attn, q/k/v by the first attention launch (no block-0 kernel);
    # then the lookup kernel writing q/k/v to HBM; then no table at all (embedding + ln_1 + c_attn computed per step)
    for env in ({}, {"MAPF_GPT_B200_NO_BLOCK0_GATHER": "1"}, {"MAPF_GPT_B200_NO_BLOCK0_TABLE": "1"}):
        for k in ("MAPF_GPT_B200_NO_BLOCK0_GATHER", "MAPF_GPT_B200_NO_BLOCK0_TABLE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = E.RolloutEngine(1, 1, 11, 11)
        eng.load_model(sd, cfg)
        outs.append(eng.forward_tokens(toks))
        eng.close()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_generic_path_kernel_variants_agree(dev, monkeypatch):
    """C = 768 (85M shape, 2 layers to keep it short): persistent CTA-pair GEMMs + TMEM-resident probabilities (defaults) against the
    single-CTA GEMM and the smem-P attention kernel.  Same bf16 operands and fp32 accumulation order per output element; the
    attention variants differ in where the row sum is taken (fp32 threads vs bf16 tensor core), hence a tolerance, not equality."""
    from mapf_gpt_b200 import engine as E, weights as W
    cfg = W.GPTConfig(block_size=256, vocab_size=67, n_layer=2, n_head=12, n_embd=768, dropout=0.0, bias=False)
    sd = W.scale_weights(W.perturb_layernorm(W.random_init(cfg)), 3.0)
    toks = np.random.default_rng(9).integers(0, 67, (66, 256)).astype(np.int8)
    outs = {}
    for name, env in (("default", {}), ("ln_kernel", {"MAPF_GPT_B200_LN_FUSED": "0"}), ("single_cta_gemm", {"MAPF_GPT_B200_GEMM_PAIR": "0"}),
                      ("smem_p_attention", {"MAPF_GPT_B200_ATTN_CLASSIC": "1", "MAPF_GPT_B200_LN_FUSED": "0"}),
                      ("no_prune", {"MAPF_GPT_B200_NO_PRUNE": "1"})):
        for k in ("MAPF_GPT_B200_GEMM_PAIR", "MAPF_GPT_B200_ATTN_CLASSIC", "MAPF_GPT_B200_LN_FUSED", "MAPF_GPT_B200_NO_PRUNE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = E.RolloutEngine(1, 1, 11, 11)
        eng.load_model(sd, cfg)
        outs[name] = eng.forward_tokens(toks)
        eng.close()
    # LayerNorm folded into the pair GEMMs (default) vs the LayerNorm kernel: the operand is bf16(x) instead of bf16(LN(x))
    # (two independent bf16 roundings, each within LOGIT_TOL of the oracle below: their distance may reach twice that)
    assert np.abs(outs["default"] - outs["ln_kernel"]).max() < 2 * LOGIT_TOL
    assert np.abs(outs["ln_kernel"] - outs["single_cta_gemm"]).max() < 1e-3      # same arithmetic, different tiling
    assert np.abs(outs["ln_kernel"] - outs["smem_p_attention"]).max() < 2e-2
    from oracle import gpt_oracle as G
    ref = G.forward_logits(sd, cfg.n_layer, cfg.n_head, torch.from_numpy(toks.astype(np.int64)))[:, :5].numpy()
    assert np.abs(outs["default"] - ref).max() < LOGIT_TOL and np.abs(outs["ln_kernel"] - ref).max() < LOGIT_TOL
    # last block in full (every token) vs pruned to token 255 (default): same function, different kernels on the last block
    assert np.abs(outs["no_prune"] - ref).max() < LOGIT_TOL and np.abs(outs["default"] - outs["no_prune"]).max() < 2 * LOGIT_TOL


# ------------------------------------------------------------------------------------------------ large maps
def _big_grid(seed=0, h=150, w=170, p=0.18):
    from mapf_gpt_b200 import maps
    rng = np.random.default_rng(seed)
    return maps.pad_grid((rng.random((h, w)) < p).astype(np.uint8))


def test_large_map_known_answer_main_scenario(dev, built):
    """The reference's int main() (observation_generator.cpp:530-544): 256x256 free grid, agent (120,120), goal (20,200),
    through the single-env C-ABI twin.  Exercises precompute tables + windowed partial fields on the GPU."""
    import ctypes as C, hashlib
    p = lambda a: np.ascontiguousarray(a, np.int32).ctypes.data_as(C.c_void_p)
    grid = np.zeros((256, 256), np.int32)
    h = built.mg_gen_create(p(grid), 256, 256, None)
    assert h, built.mg_last_error()
    pos, goal, act = np.array([[120, 120]], np.int32), np.array([[20, 200]], np.int32), np.array([0], np.int32)
    assert built.mg_gen_create_agents(h, p(pos), p(goal), 1) == 0, built.mg_last_error()
    assert built.mg_gen_update_agents(h, p(pos), p(goal), p(act), 1) == 0
    out = np.empty((1, 256), np.int32)
    assert built.mg_gen_generate_observations(h, out.ctypes.data_as(C.c_void_p)) == 0
    built.mg_gen_destroy(h)
    assert hashlib.sha256(out[0].astype(np.int8).tobytes()).hexdigest() == \
        "896eb85aa89a369759917e5903f237dc28387e6b7d431fbd6703f302a97585e1"
    assert (out[0] == GOLD_OBS["main_row"]).all()


def test_large_map_windows_tokens_and_recompute_triggers(dev):
    """160x180 random grid: partial fields (bounds + values) and tokens vs the C oracle while agents move across window
    boundaries and goals change (cpp:200-286, 464-481)."""
    import oracle
    from mapf_gpt_b200 import engine as E, maps
    grid = _big_grid()
    m = {"name": "big", "grid": grid, "starts": np.zeros(grid.shape, bool), "goals": np.zeros(grid.shape, bool)}
    n, envs = 48, 2
    inst = [maps.sample_instance(m, n, 5, e) for e in range(envs)]
    st, gl = np.stack([i[0] for i in inst]), np.stack([i[1] for i in inst])
    eng = E.RolloutEngine(envs, n, *grid.shape)
    eng.reset(0, grid, st, gl)
    orc = [oracle.ObsOracle(grid) for _ in range(envs)]
    for e in range(envs):
        orc[e].create_agents(st[e], gl[e])
        for a in range(0, n, 7):
            b, f = eng.partial(e, a)
            ob, of = orc[e].partial(a)
            assert b == ob and (f == of).all(), (e, a, b, ob)
    rng = np.random.default_rng(1)
    free = np.argwhere(maps.largest_component(grid))
    pos, act = st.copy(), np.full((envs, n), -1, np.int32)
    for t in range(40):
        if t % 9 == 4:
            gl = gl.copy()
            idx = rng.integers(0, n, 6)
            gl[:, idx] = free[rng.integers(0, len(free), (envs, 6))]
        eng.update_agents(pos, gl, act)
        toks = eng.generate_observations()
        for e in range(envs):
            orc[e].update_agents(pos[e], gl[e], act[e])
            assert (toks[e] == orc[e].generate_observations()).all(), f"step {t} env {e}"
        # drift everybody the same way so that FOVs cross the 64-cell window lines
        act = np.where(rng.random((envs, n)) < 0.7, 4 if t < 20 else 2, rng.integers(0, 5, (envs, n))).astype(np.int32)
        newpos = eng.env_step(act)
        for e in range(envs):
            pos[e], _ = oracle.pogema_step_soft(grid, pos[e], act[e])
        assert (newpos == pos).all()
    for e in range(envs):
        for a in range(0, n, 5):
            b, f = eng.partial(e, a)
            ob, of = orc[e].partial(a)
            assert b == ob and (f == of).all()
    eng.close()


def test_large_map_device_rollout(dev):
    """Device-resident rollout on a large grid: the per-step window trigger + partial recompute keep tokens exact."""
    import oracle
    from mapf_gpt_b200 import engine as E, maps, weights as W
    grid = _big_grid(3, 120, 200)
    m = {"name": "big", "grid": grid, "starts": np.zeros(grid.shape, bool), "goals": np.zeros(grid.shape, bool)}
    n = 40
    st, gl = maps.sample_instance(m, n, 2)
    cfg = W.model_config("2M")
    eng = E.RolloutEngine(1, n, *grid.shape)
    eng.load_model(W.scale_weights(W.random_init(cfg), 3.0), cfg)
    eng.reset(0, grid, st, gl)
    eng.rollout(12, E.MODE_PHILOX)
    pos = eng.positions()[0]
    assert (grid[pos[:, 0], pos[:, 1]] == 0).all() and len({tuple(p) for p in pos.tolist()}) == n
    # one more observe on the device, rebuilt on the CPU from the device's own positions (history masked)
    eng.update_agents()
    tok = eng.generate_observations()[0].astype(np.int32)
    o = oracle.ObsOracle(grid)
    o.create_agents(pos, gl)
    o.update_agents(pos, gl, np.full(n, -1, np.int32))
    ref = o.generate_observations()
    for s_ in range(13):
        tok[:, 125 + 10 * s_:130 + 10 * s_] = 0
        ref[:, 125 + 10 * s_:130 + 10 * s_] = 0
    assert (tok == ref).all()
    eng.close()


def test_cost2go_cache_file_matches_reference_and_reloads(dev, tmp_path, monkeypatch):
    """save_cost2go (cpp:62-80,114-131): the engine writes precomputed_cost2go.bin byte-identical to the file the unmodified
    reference writes for the same map (tests/golden/make_cache_golden.py), reloads it instead of recomputing (tokens stay
    exact vs the oracle), and refuses a table of the wrong shape instead of using it."""
    import hashlib, json, importlib.util
    import oracle
    from mapf_gpt_b200 import engine as E, maps, _lib
    gold = json.loads((GOLDEN / "cost2go_cache_golden.json").read_text())
    spec = importlib.util.spec_from_file_location("make_cache_golden", GOLDEN / "make_cache_golden.py")
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    grid = mk.cache_grid()
    m = {"name": "cache", "grid": grid, "starts": np.zeros(grid.shape, bool), "goals": np.zeros(grid.shape, bool)}
    n = 24
    st, gl = maps.sample_instance(m, n, 3)
    monkeypatch.chdir(tmp_path)
    f = tmp_path / "precomputed_cost2go.bin"

    def tokens(params):
        eng = E.RolloutEngine(1, n, *grid.shape, params=params)
        eng.reset(0, grid, st[None], gl[None])
        eng.update_agents(None, None, None)
        t = eng.generate_observations()[0]
        eng.close()
        return t

    orc = oracle.ObsOracle(grid)
    orc.create_agents(st, gl)
    orc.update_agents(st, gl, np.full(n, -1, np.int32))
    want = orc.generate_observations()
    assert (tokens(None) == want).all() and not f.exists()            # default: no file is touched
    assert (tokens({"save_cost2go": 1}) == want).all()
    raw = f.read_bytes()
    assert len(raw) == gold["bytes"] and hashlib.sha256(raw).hexdigest() == gold["sha256"]
    stamp = f.stat().st_mtime_ns
    assert (tokens({"save_cost2go": 1}) == want).all() and f.stat().st_mtime_ns == stamp   # loaded, not rewritten
    f.write_bytes(np.array([3, 3], np.uint64).tobytes() + bytes(18))
    with pytest.raises(_lib.MgError, match="stale cache"):
        tokens({"save_cost2go": 1})
