"""Test-only checkers (CPU).  NOT product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (mapf_gpt_b200/) never does.

  * ObsOracle            ctypes front of oracle/obs_oracle.c   (our C restatement of
                         mapf_gpt/observation_generator.{h,cpp})
  * pogema_step_soft     ctypes front of oracle/pogema_oracle.c (PARITY UNPINNED, see file)
  * load_ref_module()    the reference's own observation generator compiled into
                         oracle/_ref/ by oracle/Makefile (None when it was never built)
  * gpt_oracle           torch-fp32 restatement of mapf_gpt/model.py (oracle/gpt_oracle.py)
"""
from __future__ import annotations

import ctypes
import importlib.util
import os
import subprocess
import sysconfig
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(quiet: bool = True) -> None:
    """Compile liboracle.so and, when /root/reference exists, oracle/_ref (see Makefile)."""
    subprocess.run(["make", "-C", str(_HERE), "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "liboracle.so"
        if not so.exists():
            build()
        lib = ctypes.CDLL(str(so))
        i32p = ctypes.POINTER(ctypes.c_int32)
        lib.og_create.restype = ctypes.c_void_p
        lib.og_create.argtypes = [i32p] + [ctypes.c_int] * 9
        lib.og_destroy.argtypes = [ctypes.c_void_p]
        lib.og_create_agents.argtypes = [ctypes.c_void_p, i32p, i32p, ctypes.c_int]
        lib.og_update_agents.argtypes = [ctypes.c_void_p, i32p, i32p, i32p, ctypes.c_int]
        lib.og_generate_observations.argtypes = [ctypes.c_void_p, i32p]
        lib.og_generate_observations.restype = ctypes.c_int
        lib.og_get_partial.argtypes = [ctypes.c_void_p, ctypes.c_int, i32p,
                                       ctypes.POINTER(ctypes.c_uint16), ctypes.c_int]
        lib.og_get_partial.restype = ctypes.c_int
        lib.pg_step_soft.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, i32p, i32p, i32p]
        _LIB = lib
    return _LIB


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


class ObsOracle:
    """Same verbs as the reference pybind class (observation_generator.cpp:548-563)."""

    def __init__(self, grid, cost2go_value_limit=20, num_agents=13, num_previous_actions=5,
                 context_size=256, obs_radius=5, agents_radius=5, grid_step=64):
        g, gp = _i32(grid)
        assert g.ndim == 2
        self.H, self.W = g.shape
        self._h = _lib().og_create(gp, self.H, self.W, cost2go_value_limit, num_agents,
                                   num_previous_actions, context_size, obs_radius, agents_radius,
                                   grid_step)
        self.n = 0

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().og_destroy(self._h)
            self._h = None

    def create_agents(self, positions, goals):
        p, pp = _i32(positions)
        g, gp = _i32(goals)
        self.n = len(p)
        _lib().og_create_agents(self._h, pp, gp, self.n)

    def update_agents(self, positions, goals, actions):
        p, pp = _i32(positions)
        g, gp = _i32(goals)
        a, ap = _i32(actions)
        _lib().og_update_agents(self._h, pp, gp, ap, self.n)

    def generate_observations(self) -> np.ndarray:
        out = np.empty((self.n, 256), dtype=np.int32)
        rc = _lib().og_generate_observations(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        if rc != 0:
            raise IndexError("token outside vocabulary (int_vocab.at would throw, cpp:357-361)")
        return out

    def partial(self, agent: int):
        b = np.zeros(4, dtype=np.int32)
        n = _lib().og_get_partial(self._h, agent, b.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), None, 0)
        buf = np.empty(n, dtype=np.uint16)
        _lib().og_get_partial(self._h, agent, b.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                              buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), n)
        left, right, top, bottom = (int(v) for v in b)
        rows = n // (bottom - top + 1) if (right - left + 1) * (bottom - top + 1) == n else self.H
        return (left, right, top, bottom), buf.reshape(rows, -1)


def pogema_step_soft(obstacles, positions, actions):
    """One `soft` step.  Returns (new_positions, moved_flags)."""
    ob = np.ascontiguousarray(obstacles != 0, dtype=np.uint8)
    H, W = ob.shape
    pos = np.array(positions, dtype=np.int32, order="C").copy()
    act, ap = _i32(actions)
    moved = np.zeros(len(pos), dtype=np.int32)
    _lib().pg_step_soft(ob.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), H, W, len(pos),
                        pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ap,
                        moved.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return pos, moved


def load_ref_module():
    """The reference's pybind module built into oracle/_ref, or None."""
    import sys
    for name in ("observation_generator", "mapf_gpt.observation_generator"):   # a pybind module registers its types once per
        if name in sys.modules:                                                 # process: reuse whichever copy is loaded
            return sys.modules[name]
    so = _HERE / "_ref" / ("observation_generator" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not so.exists():
        return None
    spec = importlib.util.spec_from_file_location("observation_generator", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["observation_generator"] = mod
    return mod
