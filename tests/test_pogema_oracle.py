"""CPU: the POGEMA `soft` step restatement (oracle/pogema_oracle.c, PARITY UNPINNED).
Scenario table from SURVEY App. C.3, invariants, and equivalence of the procedural form with the
order-independent fixed point the CUDA kernel implements."""
import numpy as np
import pytest

import oracle

MOVES = np.array([(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)])
W0, UP, DOWN, LEFT, RIGHT = range(5)


def room(h=9, w=9):
    g = np.ones((h, w), np.uint8)
    g[1:-1, 1:-1] = 0
    return g


def fixed_point_step(grid, pos, act):
    """Declarative form: least set W of waiting agents."""
    n = len(pos)
    pos = np.asarray(pos)
    act = np.where((np.asarray(act) >= 0) & (np.asarray(act) <= 4), act, 0)
    tgt = pos + MOVES[act]
    occ = {tuple(p): i for i, p in enumerate(pos)}
    wait = [(act[i] == 0) or grid[tuple(tgt[i])] != 0 for i in range(n)]
    swap = []
    for i in range(n):
        if wait[i]:
            continue
        j = occ.get(tuple(tgt[i]))
        if j is not None and j != i and act[j] != 0 and tuple(tgt[j]) == tuple(pos[i]):
            swap.append(i)
    for i in swap:
        wait[i] = True
    claim = {}
    for i in range(n):
        if not wait[i]:
            claim.setdefault(tuple(tgt[i]), i)            # lowest index first
    for i in range(n):
        if not wait[i] and claim[tuple(tgt[i])] != i:
            wait[i] = True
    changed = True
    while changed:
        changed = False
        for i in range(n):
            if wait[i]:
                continue
            k = occ.get(tuple(tgt[i]))
            if k is not None and wait[k]:
                wait[i] = True
                changed = True
    return np.where(np.array(wait)[:, None], pos, tgt)


SCENARIOS = [
    # (positions, actions, expected positions)
    ("free move", [(2, 2)], [RIGHT], [(2, 3)]),
    ("into wall", [(1, 1)], [UP], [(1, 1)]),
    ("swap -> both wait", [(2, 2), (2, 3)], [RIGHT, LEFT], [(2, 2), (2, 3)]),
    ("vertex conflict: lowest index wins", [(2, 2), (2, 4)], [RIGHT, LEFT], [(2, 3), (2, 4)]),
    ("vertex conflict, three movers", [(2, 3), (3, 2), (3, 4)], [DOWN, RIGHT, LEFT], [(3, 3), (3, 2), (3, 4)]),
    ("follow into vacated cell", [(2, 2), (2, 3)], [RIGHT, RIGHT], [(2, 3), (2, 4)]),
    ("follower of a waiting agent waits", [(2, 2), (2, 3)], [RIGHT, W0], [(2, 2), (2, 3)]),
    ("chain behind a blocked leader", [(1, 3), (2, 3), (3, 3)], [UP, UP, UP], [(1, 3), (2, 3), (3, 3)]),
    ("rotation of four is legal", [(2, 2), (2, 3), (3, 3), (3, 2)], [RIGHT, DOWN, LEFT, UP],
     [(2, 3), (3, 3), (3, 2), (2, 2)]),
    ("cascade: loser blocks its follower", [(2, 2), (2, 4), (2, 5)], [RIGHT, LEFT, LEFT], [(2, 3), (2, 4), (2, 5)]),
    ("out-of-range action is wait", [(2, 2)], [7], [(2, 2)]),
]


@pytest.mark.parametrize("name,pos,act,want", SCENARIOS, ids=[s[0] for s in SCENARIOS])
def test_scenarios(built, name, pos, act, want):
    grid = room()
    got, _ = oracle.pogema_step_soft(grid, np.array(pos, np.int32), np.array(act, np.int32))
    assert got.tolist() == [list(w) for w in want]
    assert fixed_point_step(grid, pos, act).tolist() == [list(w) for w in want]


@pytest.mark.parametrize("density,seed", [(0.15, 0), (0.4, 1), (0.7, 2), (0.9, 3)])
def test_invariants_and_fixed_point_equivalence(built, density, seed):
    rng = np.random.default_rng(seed)
    grid = room(14, 15)
    grid[1:-1, 1:-1] = rng.random((12, 13)) < 0.15
    free = np.argwhere(grid == 0)
    n = max(2, int(density * len(free)))
    for trial in range(150):
        pos = free[rng.permutation(len(free))[:n]].astype(np.int32)
        act = rng.integers(0, 5, n).astype(np.int32)
        new, moved = oracle.pogema_step_soft(grid, pos, act)
        again, _ = oracle.pogema_step_soft(grid, pos, act)
        assert (new == again).all()                                            # deterministic
        assert len({tuple(p) for p in new.tolist()}) == n                      # no vertex conflict
        assert (grid[new[:, 0], new[:, 1]] == 0).all()                         # no obstacle entry
        step = np.abs(new - pos).sum(1)
        assert (step <= 1).all()
        assert ((step == 1) == (moved == 1)).all()
        assert (new[act == 0] == pos[act == 0]).all()                          # waiting agents keep their cell
        was = {tuple(p): i for i, p in enumerate(pos.tolist())}
        for i in np.flatnonzero(moved):                                        # no edge swap
            j = was.get(tuple(new[i].tolist()))
            if j is not None:
                assert tuple(new[j].tolist()) != tuple(pos[i].tolist())
        assert (new[moved == 1] == (pos + MOVES[act])[moved == 1]).all()
        assert (fixed_point_step(grid, pos, act) == new).all()                 # the form the CUDA kernel uses
