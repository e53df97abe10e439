"""CPU: the C restatement of the observation generator (oracle/obs_oracle.c) against
(a) the reference's known-answer scenario, (b) committed golden tokens produced by the compiled
reference, (c) the compiled reference itself when oracle/_ref is present."""
import hashlib

import numpy as np
import pytest

import oracle
from mapf_gpt_b200 import maps
from pathlib import Path

GOLD = np.load(Path(__file__).parent / "golden" / "obs_golden.npz")
SCENARIOS = [k for k in GOLD.files if k.endswith("_tokens")]   # written by tests/golden/make_golden.py


def test_known_answer_main(built):
    """int main() of observation_generator.cpp:530-544 (256x256 free grid, >64 code path)."""
    o = oracle.ObsOracle(np.zeros((256, 256), np.int32))
    o.create_agents([(120, 120)], [(20, 200)])
    o.update_agents([(120, 120)], [(20, 200)], [0])
    row = o.generate_observations()[0]
    assert hashlib.sha256(row.astype(np.int8).tobytes()).hexdigest() == \
        "896eb85aa89a369759917e5903f237dc28387e6b7d431fbd6703f302a97585e1"
    assert (row == GOLD["main_row"]).all()
    # structure stated in SURVEY section 4: anti-diagonal ramp, self slot, padding
    w = row[:121].reshape(11, 11)
    i, j = np.indices((11, 11))
    assert (w == 20 + i - j).all()
    assert row[121:131].tolist() == [20, 20, 0, 40, 44, 44, 44, 44, 45, 59]
    assert (row[131:] == 66).all()


@pytest.mark.parametrize("k", range(len(SCENARIOS)))
def test_golden_tokens(built, k):
    grid, goals = GOLD[f"s{k}_grid"], GOLD[f"s{k}_goals"]
    pos, act, tok = GOLD[f"s{k}_pos"], GOLD[f"s{k}_act"], GOLD[f"s{k}_tokens"]
    o = oracle.ObsOracle(grid)
    o.create_agents(pos[0], goals)
    for t in range(len(pos)):
        o.update_agents(pos[t], goals, act[t])
        assert (o.generate_observations() == tok[t].astype(np.int32)).all(), f"scenario {k} step {t}"
    assert tok.min() >= 0 and tok.max() <= 66


def _rollout_pair(ref, grid, n, steps, seed, change_goals=False):
    m = {"name": "x", "grid": grid, "starts": np.zeros(grid.shape, bool), "goals": np.zeros(grid.shape, bool)}
    st, gl = maps.sample_instance(m, n, seed)
    P = ref.InputParameters(20, 13, 5, 256, 5, 5, 64, False)
    o, g = oracle.ObsOracle(grid), ref.ObservationGenerator(grid.astype(int).tolist(), P)
    tup = lambda a: [tuple(x) for x in a.tolist()]
    o.create_agents(st, gl)
    g.create_agents(tup(st), tup(gl))
    rng = np.random.default_rng(seed)
    pos, act, bad = st.copy(), np.full(n, -1, np.int32), 0
    free = np.argwhere(maps.largest_component(grid))
    for t in range(steps):
        if change_goals and t % 7 == 3:          # lifelong-style goal changes -> recompute trigger (cpp:464-468)
            idx = rng.integers(0, n, max(1, n // 8))
            gl = gl.copy()
            gl[idx] = free[rng.integers(0, len(free), len(idx))]
        o.update_agents(pos, gl, act)
        g.update_agents(tup(pos), tup(gl), act.tolist())
        bad += int((o.generate_observations() != np.asarray(g.generate_observations())).sum())
        act = rng.integers(-1, 6, n).astype(np.int32)        # includes out-of-range -> "n"
        pos, _ = oracle.pogema_step_soft(grid, pos, np.clip(act, 0, 4))
    return bad


def test_live_reference_small_maps(built):
    ref = oracle.load_ref_module()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for name, n in [("validation-random-seed-001", 24), ("validation-mazes-seed-002", 48), ("puzzle-03", 3)]:
        grid = maps.load_map(name, solid_padding=False)["grid"]
        assert _rollout_pair(ref, grid, n, 25, 5, change_goals=True) == 0


def test_live_reference_baseline_maps(built):
    """The other BASELINE.json map families at their agent counts (warehouse 192, a Berlin tile 256, a puzzle at 4), goals changing
    on the way: the C restatement against the unmodified reference, every token of every agent."""
    ref = oracle.load_ref_module()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for name, n, steps in [("wfi_warehouse", 192, 12), ("Berlin_1_256_07", 256, 10), ("puzzle-11", 4, 30)]:
        grid = maps.load_map(name, solid_padding=False)["grid"]
        assert _rollout_pair(ref, grid, n, steps, 3, change_goals=True) == 0, name


def test_live_reference_large_map_windows(built):
    """> 128 cells: the windowed multi-source BFS (cpp:200-286) and the recompute trigger."""
    ref = oracle.load_ref_module()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(0)
    grid = maps.pad_grid((rng.random((140, 150)) < 0.18).astype(np.uint8))
    assert _rollout_pair(ref, grid, 60, 80, 9, change_goals=True) == 0


def test_vocab_and_layout_properties(built):
    """Size-independent properties of the token rows (App. A)."""
    for k in range(len(SCENARIOS)):
        tok = GOLD[f"s{k}_tokens"].astype(np.int32)
        c2g, slots, tail = tok[..., :121], tok[..., 121:251].reshape(tok.shape[:-1] + (13, 10)), tok[..., 251:]
        assert (tail == 66).all()
        assert (c2g <= 43).all() and (c2g[..., 60] == 20).all()          # centre cell: own cost - own cost = 0
        assert (slots[..., 0, 0] == 20).all() and (slots[..., 0, 1] == 20).all()   # slot 0 = the agent itself
        used = slots[..., 0] != 66
        assert ((slots[..., 4:9][used] >= 44) & (slots[..., 4:9][used] <= 49)).all()
        assert ((slots[..., 9][used] >= 50) & (slots[..., 9][used] <= 65)).all()
        assert (slots[~used] == 66).all()
        d = np.abs(slots[..., 0] - 20) + np.abs(slots[..., 1] - 20)       # nearest first
        d = np.where(used, d, 99)
        assert (np.diff(d, axis=-1) >= 0).all()
