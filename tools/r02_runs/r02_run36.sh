#!/bin/bash
# 24-bit residual stream on the 85M (generic, CTA-pair GEMM) path: tests, flip rate, A/B on the C4 shard
cd "$(dirname "$0")/../.."
O=gpurun_out/r02ao; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py tests/test_gpu_rollout.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/tests.log
python tools/flip_rate.py 85M > $O/flip_85M.txt 2>&1; cut -c1-300 $O/flip_85M.txt | tail -3; MAPF_GPT_B200_X24=0 python tools/flip_rate.py 85M > $O/flip_85M_fp32.txt 2>&1; cut -c1-300 $O/flip_85M_fp32.txt | tail -3
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
run fp32_$rep MAPF_GPT_B200_X24=0
run x24_$rep X=1
done
