#!/bin/bash
# tail-aware stores (the block before the pruned last one stores x' and q for token 255 only) + successor L2 prefetch: tests and A/B
cd "$(dirname "$0")/../.."
O=gpurun_out/r02am; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py tests/test_gpu_rollout.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 8 --warmup 3 $ARGS > $O/b_$name.json 2>$O/b_$name.err
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
ARGS=""
run old_$rep MAPF_GPT_B200_POST_PF=0 MAPF_GPT_B200_FULL_TAIL_STORES=1
run pf_$rep MAPF_GPT_B200_FULL_TAIL_STORES=1
run tail_$rep MAPF_GPT_B200_POST_PF=0
run new_$rep X=1
ARGS="--model 6M --map wfi_warehouse --agents 192 --envs 512 --steps 4"
run 6M_old_$rep MAPF_GPT_B200_FULL_TAIL_STORES=1
run 6M_new_$rep X=1
done
