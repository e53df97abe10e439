"""ctypes binding of libmapf_gpt_b200.so (the C ABI in include/mapf_gpt_b200.h).

The product path fails loudly when the CUDA library is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
# MAPF_GPT_B200_LIB_PATH: developer switch for A/B runs of compile-time kernel variants (csrc/Makefile OUT= / EXTRA=)
LIB_PATH = Path(os.environ.get("MAPF_GPT_B200_LIB_PATH") or _HERE / "libmapf_gpt_b200.so")
_lib = None

MG_OK, MG_ERR_ARG, MG_ERR_CUDA, MG_ERR_STATE, MG_ERR_VOCAB, MG_ERR_NUMERIC = 0, -1, -2, -3, -4, -5
MG_METRIC_COLS = 10


class MgParams(C.Structure):  # mg_params == InputParameters (observation_generator.h:22-40)
    _fields_ = [(n, C.c_int32) for n in ("cost2go_value_limit", "num_agents", "num_previous_actions",
                                         "context_size", "obs_radius", "agents_radius", "grid_step",
                                         "save_cost2go")]


class MgModelConfig(C.Structure):  # GPTConfig (model.py:107-115)
    _fields_ = [(n, C.c_int32) for n in ("block_size", "vocab_size", "n_layer", "n_head", "n_embd")]


class MgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmapf_gpt_b200: {msg} (code {code})")
        self.code = code


def build(verbose: bool = False) -> Path:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (csrc/Makefile); in-tree output."""
    subprocess.run(["make", "-C", str(_HERE / "csrc")], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_SIGS = {
    # name: (restype, argtypes)
    "mg_version": (C.c_int, []),
    "mg_last_error": (C.c_char_p, []),
    "mg_default_params": (None, [C.POINTER(MgParams)]),
    "mg_device_count": (C.c_int, []),
    "mg_engine_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(MgParams)]),
    "mg_engine_destroy": (None, [C.c_void_p]),
    "mg_engine_load_model": (C.c_int, [C.c_void_p, C.POINTER(MgModelConfig), C.c_void_p, C.c_size_t]),
    "mg_model_num_floats": (C.c_size_t, [C.POINTER(MgModelConfig)]),
    "mg_engine_reset": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_engine_num_envs": (C.c_int, [C.c_void_p]),
    "mg_engine_clear": (C.c_int, [C.c_void_p]),
    "mg_engine_update_agents": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_engine_generate_observations": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mg_engine_act": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_engine_set_active": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mg_engine_set_seed": (C.c_int, [C.c_void_p, C.c_uint64]),
    "mg_engine_set_env_offset": (C.c_int, [C.c_void_p, C.c_int]),
    "mg_engine_set_max_episode_steps": (C.c_int, [C.c_void_p, C.c_int]),
    "mg_engine_forward_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "mg_engine_eval_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mg_engine_env_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mg_engine_rollout": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mg_engine_act_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mg_engine_get_positions": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mg_engine_get_tokens": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mg_engine_get_cost2go": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mg_engine_get_partial": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "mg_engine_get_metrics": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mg_engine_synchronize": (C.c_int, [C.c_void_p]),
    "mg_engine_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "mg_engine_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "mg_engine_launch_count": (C.c_longlong, [C.c_void_p]),
    "mg_engine_num_lanes": (C.c_int, [C.c_void_p]),
    "mg_engine_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "mg_gen_create": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.POINTER(MgParams)]),
    "mg_gen_create_agents": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "mg_gen_update_agents": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "mg_gen_generate_observations": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mg_gen_destroy": (None, [C.c_void_p]),
    "mg_test_timeline": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mg_test_gemm_time": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "mg_test_umma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mg_test_gemm": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mg_test_attention": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "mg_test_attention_ex": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mg_set_precision": (C.c_int, [C.c_int]),
}

EXPORTED = sorted(_SIGS)


def lib():
    """Load the shared library (never builds implicitly on a GPU box: the .so ships in-tree)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {_HERE / 'csrc'}`.  mapf_gpt_b200 has no CPU fallback.")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MgError(rc, lib().mg_last_error().decode(errors="replace"))
