"""Golden vector for the precomputed_cost2go.bin cache (observation_generator.cpp:62-80,114-131).

Run in the build container (needs oracle/_ref, i.e. /root/reference):  python tests/golden/make_cache_golden.py
The UNMODIFIED reference generator is constructed with save_cost2go=True in an empty working directory; the file it writes is
recorded by size and sha256 (the table itself is ~90 KB, the map is regenerated from its seed by the test)."""
import hashlib, json, os, sys, tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle
from mapf_gpt_b200 import maps

SEED, H, W, P_OBST = 11, 84, 96, 0.15


def cache_grid():
    rng = np.random.default_rng(SEED)
    return maps.pad_grid((rng.random((H, W)) < P_OBST).astype(np.uint8))


if __name__ == "__main__":
    oracle.build()
    ref = oracle.load_ref_module()
    assert ref is not None, "oracle/_ref is not built"
    grid = cache_grid()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            ref.ObservationGenerator(grid.astype(int).tolist(), ref.InputParameters(20, 13, 5, 256, 5, 5, 64, True))
            raw = Path("precomputed_cost2go.bin").read_bytes()
        finally:
            os.chdir(cwd)
    rows, cols = np.frombuffer(raw[:16], np.uint64)
    out = {"seed": SEED, "H": H, "W": W, "p_obst": P_OBST, "rows": int(rows), "cols": int(cols), "bytes": len(raw),
           "sha256": hashlib.sha256(raw).hexdigest()}
    (Path(__file__).parent / "cost2go_cache_golden.json").write_text(json.dumps(out, indent=1) + "\n")
    print(out)
