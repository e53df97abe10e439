"""Benchmark maps and the synthetic start/goal sampler.

The reference gets maps from eval_configs/<set>/maps.yaml through
ToolboxRegistry.register_maps (example.py:29-32, benchmark.py:37-40) and starts/goals
from POGEMA's own seeded sampler, which is not available offline.  This module
holds every map the five eval sets name (401, extracted once by tools/extract_maps.py
into data/maps.json.gz), accepts more through register_maps(), and has a
deterministic sampler of our own; seeds are ours, not POGEMA's.
"""
from __future__ import annotations

import gzip
import json
from collections import deque
from functools import lru_cache
from pathlib import Path

import numpy as np

OBS_RADIUS = 5  # POGEMA pads every side by obs_radius (example.py:48)
_DATA = Path(__file__).resolve().parent / "data" / "maps.json.gz"
EVAL_CONFIGS = Path(__file__).resolve().parent / "data" / "eval_configs"   # packaged sweep descriptions (benchmark.py)


@lru_cache(maxsize=1)
def _maps() -> dict:
    with gzip.open(_DATA, "rt") as f:
        return json.load(f)


def register_maps(maps: dict, set_name: str = "user") -> None:
    """ToolboxRegistry.register_maps (benchmark.py:37-40): {name: multi-line map string} as maps.yaml holds them."""
    for n, txt in maps.items():
        _maps()[n] = {"set": set_name, "rows": str(txt).split("\n")}


def map_names() -> list[str]:
    return list(_maps())


def parse_map(rows: list[str]):
    """rows of '#', '.', '@', '$', '!' -> (obstacles uint8 HxW, start_mask, goal_mask)."""
    h, w = len(rows), max(len(r) for r in rows)
    obst = np.ones((h, w), dtype=np.uint8)
    starts = np.zeros((h, w), dtype=bool)
    goals = np.zeros((h, w), dtype=bool)
    for i, r in enumerate(rows):
        for j, c in enumerate(r):
            if c != "#":
                obst[i, j] = 0
            if c == "@":
                starts[i, j] = True
            elif c == "$":
                goals[i, j] = True
    return obst, starts, goals


def pad_grid(obst: np.ndarray, r: int = OBS_RADIUS, solid: bool = True) -> np.ndarray:
    """Pad by r cells.  solid=True: all padding is obstacle.  solid=False: POGEMA style,
    a one-cell obstacle ring around the map and free-but-unreachable cells outside it.
    Both give identical tokens."""
    h, w = obst.shape
    out = np.ones((h + 2 * r, w + 2 * r), dtype=np.uint8) if solid else np.zeros((h + 2 * r, w + 2 * r), dtype=np.uint8)
    if not solid:
        out[r - 1:r + h + 1, r - 1:r + w + 1] = 1
    out[r:r + h, r:r + w] = obst
    return out


def load_map(name: str, solid_padding: bool = True):
    """-> dict(name, grid (padded uint8), starts, goals (padded bool masks, may be empty))."""
    m = _maps()[name]
    obst, st, gl = parse_map(m["rows"])
    r = OBS_RADIUS
    grid = pad_grid(obst, r, solid_padding)
    pst = np.zeros(grid.shape, dtype=bool)
    pgl = np.zeros(grid.shape, dtype=bool)
    pst[r:-r, r:-r] = st
    pgl[r:-r, r:-r] = gl
    return {"name": name, "set": m["set"], "grid": grid, "starts": pst, "goals": pgl}


def largest_component(grid: np.ndarray, inner_only: bool = True) -> np.ndarray:
    """Bool mask of the largest 4-connected free component (inside the padding)."""
    h, w = grid.shape
    r = OBS_RADIUS if inner_only else 0
    lab = np.zeros((h, w), dtype=np.int32)
    best, best_n, cur = 0, 0, 0
    for i in range(r, h - r):
        for j in range(r, w - r):
            if grid[i, j] or lab[i, j]:
                continue
            cur += 1
            lab[i, j] = cur
            q = deque([(i, j)])
            n = 0
            while q:
                a, b = q.popleft()
                n += 1
                for da, db in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                    x, y = a + da, b + db
                    if r <= x < h - r and r <= y < w - r and not grid[x, y] and not lab[x, y]:
                        lab[x, y] = cur
                        q.append((x, y))
            if n > best_n:
                best, best_n = cur, n
    return lab == best if best else np.zeros((h, w), dtype=bool)


def sample_instance(m: dict, num_agents: int, seed: int, env: int = 0):
    """Distinct starts, distinct goals, start != goal per agent, all in the largest
    component; warehouse-style maps draw starts from '@' and goals from '$' cells.
    rng = default_rng(seed * 1_000_003 + env)  (SURVEY 8d)."""
    rng = np.random.default_rng(seed * 1_000_003 + env)
    comp = largest_component(m["grid"]) if "_comp" not in m else m["_comp"]
    m["_comp"] = comp
    s_mask = comp & m["starts"] if m["starts"].any() else comp
    g_mask = comp & m["goals"] if m["goals"].any() else comp
    s_cells = np.argwhere(s_mask)
    g_cells = np.argwhere(g_mask)
    if len(s_cells) < num_agents or len(g_cells) < num_agents:
        raise ValueError(f"map {m['name']}: {num_agents} agents do not fit "
                         f"({len(s_cells)} start cells, {len(g_cells)} goal cells)")
    for _ in range(64):
        starts = s_cells[rng.permutation(len(s_cells))[:num_agents]]
        goals = g_cells[rng.permutation(len(g_cells))[:num_agents]]
        same = (starts == goals).all(axis=1)
        if not same.any():
            break
        # rotate the goals of the colliding agents among themselves / with a neighbour
        idx = np.flatnonzero(same)
        for i in idx:
            j = (i + 1) % num_agents
            goals[[i, j]] = goals[[j, i]]
        if not (starts == goals).all(axis=1).any():
            break
    else:  # pragma: no cover
        raise RuntimeError("could not sample start != goal")
    return starts.astype(np.int32), goals.astype(np.int32)
