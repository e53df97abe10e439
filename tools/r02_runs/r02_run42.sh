#!/bin/bash
# 85M path: embed_tile_kernel (coalesced reads, slab transposed through shared memory) vs embed_kernel
cd "$(dirname "$0")/../.."
O=gpurun_out/r02ax; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py -m gpu -x -q -k "embed_tile or 85M or golden" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 3 --warmup 3 --model 85M --map Berlin_1_256_05 --agents 256 --envs 32 > $O/b_$name.json 2>$O/b_$name.err
  python - <<PY
import json
f="$O/b_$name.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("$name", round(d['value']), d['roofline']['whole_step_frac'], round(d['ms_per_step'],2), {k:v['avg_ms'] for k,v in d['kernels'].items() if v['share']>0.01}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-600:])
PY
}
for rep in 1 2; do
run rows_$rep MAPF_GPT_B200_EMBED_ROWS=1
run tile_$rep X=1
done
