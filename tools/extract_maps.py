"""Extract the benchmark maps named in BASELINE.json from the reference's eval_configs
(read-only input data, eval_configs/<set>/maps.yaml) into mapf_gpt_b200/data/maps.json.

Run once in the build container (needs /root/reference); the JSON travels with the repo.
Map strings: '#' obstacle; '.', '@' (start cells), '$' (goal cells), '!' free.
"""
import json
import sys
from pathlib import Path

import yaml

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parents[1] / "mapf_gpt_b200" / "data" / "maps.json"

WANT = {
    "01-random": ["validation-random-seed-%03d" % i for i in range(4)],
    "02-mazes": ["validation-mazes-seed-%03d" % i for i in range(4)],
    "03-warehouse": ["wfi_warehouse"],
    "04-movingai": ["Berlin_1_256_%02d" % i for i in range(16)],
    "05-puzzles": ["puzzle-%02d" % i for i in range(16)],
}

out = {}
for folder, names in WANT.items():
    maps = yaml.safe_load(open(REF / "eval_configs" / folder / "maps.yaml"))
    for n in names:
        if n in maps:
            out[n] = {"set": folder, "rows": maps[n].split("\n")}
OUT.parent.mkdir(parents=True, exist_ok=True)
json.dump(out, open(OUT, "w"), separators=(",", ":"))
print(f"wrote {len(out)} maps to {OUT} ({OUT.stat().st_size} bytes)")
