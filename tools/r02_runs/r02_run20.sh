#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r02t; mkdir -p $O
timeout 300 python bench.py --quick --steps 8 --warmup 3 > $O/b2M.json 2>$O/b2M.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02t/b2M.json').read().strip().splitlines()[-1]); print(round(d['value']), d['clocks'])
PY
tail -3 $O/b2M.err
