"""CPU: host-side logic and the C-ABI library's load/export contract (no compute without a GPU)."""
import ctypes
import os
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(built):
    hdr = (ROOT / "include" / "mapf_gpt_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 35
    lib = ctypes.CDLL(str(ROOT / "mapf_gpt_b200" / "libmapf_gpt_b200.so"))
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    from mapf_gpt_b200 import _lib
    assert sorted(_lib.EXPORTED) == declared, "ctypes binding and header disagree"
    assert built.mg_version() >= 100


def test_no_silent_cpu_fallback(built):
    """Without a GPU every compute entry fails loudly; with one this test is vacuous."""
    from mapf_gpt_b200 import _lib, engine
    if built.mg_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.MgError, match="no CUDA device"):
        engine.RolloutEngine(1, 4, 31, 31)
    from mapf_gpt_b200.inference import MAPFGPTInference, MAPFGPTInferenceConfig
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MAPFGPTInference(MAPFGPTInferenceConfig(), net=({}, None))


def test_product_never_imports_oracle():
    for p in (ROOT / "mapf_gpt_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h"):
            txt = p.read_text()
            assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, p


def test_config_fields_match_reference():
    """Field names and defaults of mapf_gpt/inference.py:13-31."""
    from mapf_gpt_b200.inference import MAPFGPTInferenceConfig
    want = dict(name="MAPF-GPT", num_agents=13, num_previous_actions=5, cost2go_value_limit=20, agents_radius=5,
                cost2go_radius=5, path_to_weights="weights/MAPF-GPT-2M.pt", device=None, context_size=256,
                mask_actions_history=False, mask_goal=False, mask_cost2go=False, mask_greed_action=False,
                repo_id="aandreychuk/MAPF-GPT", grid_step=64, save_cost2go=False, batch_size=2048, num_process=8)
    cfg = MAPFGPTInferenceConfig()
    for k, v in want.items():
        assert getattr(cfg, k) == v, k
    with pytest.raises(Exception):
        MAPFGPTInferenceConfig(not_a_field=1)          # extra=forbid


def test_params_struct_matches_input_parameters(built):
    from mapf_gpt_b200 import _lib
    p = _lib.MgParams()
    built.mg_default_params(ctypes.byref(p))
    # InputParameters defaults, observation_generator.h:24
    assert [getattr(p, f) for f, _ in _lib.MgParams._fields_] == [20, 13, 5, 256, 5, 5, 64, 0]


def test_weight_flattening_and_checkpoint_roundtrip(built, tmp_path):
    import torch
    from mapf_gpt_b200 import _lib, engine, weights as W
    for name in ("2M", "6M"):
        cfg = W.model_config(name)
        sd = W.random_init(cfg)
        flat = engine.flatten_weights(sd, cfg)
        mc = _lib.MgModelConfig(cfg.block_size, cfg.vocab_size, cfg.n_layer, cfg.n_head, cfg.n_embd)
        assert flat.size == built.mg_model_num_floats(ctypes.byref(mc))
    cfg = W.model_config("2M")
    sd = W.random_init(cfg)
    path = tmp_path / "ck.pt"
    # reference layout, with torch.compile's prefix (inference.py:33-44)
    torch.save({"model": {"_orig_mod." + k: v for k, v in sd.items()}, "model_args": cfg.__dict__}, path)
    sd2, cfg2 = W.load_checkpoint(path)
    assert cfg2 == cfg and set(sd2) == set(sd)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert W.state_dict_digest(W.random_init(cfg)) == W.state_dict_digest(sd)     # seeded init is reproducible


def test_maps_and_sampler():
    from mapf_gpt_b200 import maps
    m = maps.load_map("validation-mazes-seed-000")
    assert m["grid"].shape == (31, 31) and int((m["grid"] == 0).sum()) == 305
    assert maps.load_map("wfi_warehouse")["grid"].shape == (43, 56)
    assert maps.load_map("Berlin_1_256_00")["grid"].shape == (74, 74)
    a = maps.load_map("validation-random-seed-000", solid_padding=True)["grid"]
    b = maps.load_map("validation-random-seed-000", solid_padding=False)["grid"]
    assert a.shape == b.shape == (30, 31) and (a[5:-5, 5:-5] == b[5:-5, 5:-5]).all()
    assert (b[4, 4:-4] == 1).all() and (b[0] == 0).all()                          # wall ring, free outside
    st, gl = maps.sample_instance(m, 64, 0, 3)
    st2, gl2 = maps.sample_instance(m, 64, 0, 3)
    assert (st == st2).all() and (gl == gl2).all()
    assert len({tuple(x) for x in st.tolist()}) == 64 and len({tuple(x) for x in gl.tolist()}) == 64
    assert not (st == gl).all(1).any()
    comp = maps.largest_component(m["grid"])
    assert comp[st[:, 0], st[:, 1]].all() and comp[gl[:, 0], gl[:, 1]].all()
    w = maps.load_map("wfi_warehouse")
    st, gl = maps.sample_instance(w, 192, 0)
    assert w["starts"][st[:, 0], st[:, 1]].all() and w["goals"][gl[:, 0], gl[:, 1]].all()
    with pytest.raises(ValueError):
        maps.sample_instance(maps.load_map("puzzle-00"), 64, 0)


def test_bench_flops_formula():
    import bench
    # SURVEY 8d
    assert bench.flops_per_agent_step(5, 160) == 996_168_640
    assert bench.flops_per_agent_step(8, 256) == 3_758_130_688
    assert bench.flops_per_agent_step(12, 768) == 45_902_565_888


def test_cost2go_cache_golden_is_self_consistent():
    """precomputed_cost2go.bin (cpp:114-131): size_t rows, size_t cols, rows x cols uint16 -- the golden record written by
    tests/golden/make_cache_golden.py from the unmodified reference must describe exactly that layout."""
    import json
    g = json.loads((Path(__file__).parent / "golden" / "cost2go_cache_golden.json").read_text())
    assert g["rows"] == g["cols"] > 0
    assert g["bytes"] == 16 + 2 * g["rows"] * g["cols"]
    assert len(g["sha256"]) == 64


def test_benchmark_yaml_axes_and_tabular_view():
    """benchmark.py reads the sweep from eval_configs/<set>/<set>.yaml (reference benchmark.py:28-50): the packaged copies
    expand to the reference's 3 296 episodes per model, every named map is in the store, and the tabular view averages over
    the dropped keys."""
    import importlib.util
    import yaml
    from mapf_gpt_b200 import maps
    spec = importlib.util.spec_from_file_location("benchmark", ROOT / "benchmark.py")
    bm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bm)
    total, names = 0, set(maps.map_names())
    want = {"01-random": 768, "02-mazes": 768, "03-warehouse": 768, "04-movingai": 512, "05-puzzles": 480}
    for folder in bm.FOLDERS:
        cfg = yaml.safe_load(open(maps.EVAL_CONFIGS / folder / f"{folder}.yaml"))
        fixed, axes, combos = bm.expand_grid_search(cfg["environment"])
        assert len(combos) == want[folder] and "num_agents" in axes
        assert fixed["collision_system"] == "soft" and fixed["max_episode_steps"] in (128, 256)
        assert {({**fixed, **c})["map_name"] for c in combos} <= names
        assert set(cfg["algorithms"]) == {"MAPF-GPT-2M", "MAPF-GPT-6M"}
        total += len(combos)
    assert total == 3296
    recs = [{"algorithm": "A", "env_grid_search": {"num_agents": n, "map_name": f"m{i}"},
             "metrics": {m: float(n + i) for m in bm.METRICS}} for n in (8, 16) for i in range(3)]
    header, rows = bm.tabular_view(recs, {"type": "tabular", "drop_keys": ["seed", "map_name", "runtime"]}, ["num_agents", "map_name"])
    assert header == ["algorithm", "num_agents", "CSR", "ISR", "SoC", "makespan", "ep_length", "avg_agents_density", "episodes"]
    assert rows == [["A", 8] + [9.0] * 6 + [3], ["A", 16] + [17.0] * 6 + [3]]


def test_arrow_shard_reader_mirrors_the_reference_loader(tmp_path, monkeypatch):
    """mapf_gpt_b200.dataset (first slice): shards in the reference's schema (generate_dataset.py:188-191) read back the
    way dataset/fast_data_loader.py reads them; batches carry -1 targets except at the last position; `train` folders are split
    over LOCAL_RANK / WORLD_SIZE."""
    import pyarrow as pa
    from mapf_gpt_b200 import dataset as D
    rng = np.random.default_rng(0)
    (tmp_path / "train").mkdir()
    shards = []
    for i in range(4):
        x = rng.integers(0, 67, (50, 256)).astype(np.int8)
        y = rng.integers(0, 5, 50).astype(np.int8)
        D.write_shard(tmp_path / "train" / f"chunk_part_{i}.arrow", x, y)
        shards.append((x, y))
    with pa.memory_map(str(tmp_path / "train" / "chunk_part_2.arrow")) as src:      # the reference's own reading code (:38-43)
        table = pa.ipc.open_file(src).read_all()
        assert table.schema.names == ["input_tensors", "gt_actions"]
        assert (np.stack(table["input_tensors"].to_numpy(zero_copy_only=False)) == shards[2][0]).all()
        assert (table["gt_actions"].to_numpy(zero_copy_only=False) == shards[2][1]).all()
    ds = D.MapfArrowDataset(tmp_path / "train", device="cuda", batch_size=32, seed=0)
    assert ds.get_full_dataset_size() == 200 and ds.get_shard_size() == 200
    it = iter(ds)
    xb, tb = next(it)
    assert xb.shape == (32, 256) and tb.shape == (32, 256) and xb.dtype == np.int8 and (tb[:, :-1] == -1).all()
    rows = {bytes(r) for r in shards[0][0]}
    assert all(bytes(r) in rows for r in xb)                                       # first file, shuffled inside the file
    monkeypatch.setenv("LOCAL_RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    part = D.MapfArrowDataset(tmp_path / "train", batch_size=64)
    assert [os.path.basename(p) for p in part.file_paths] == ["chunk_part_2.arrow", "chunk_part_3.arrow"]
    assert part.get_shard_size() == 100 and part.get_full_dataset_size() == 200
