#!/bin/bash
# A/B of the persistent post_attn launch: stagger spread and grid size
cd "$(dirname "$0")/../.."
O=gpurun_out/r02e; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rollout.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 8 --warmup 3 > $O/b2M_$name.json 2>$O/b2M_$name.err
  env "$@" timeout 300 python bench.py --quick --steps 4 --warmup 3 --model 6M --map wfi_warehouse --agents 192 --envs 512 > $O/b6M_$name.json 2>$O/b6M_$name.err
  python - <<PY
import json
for f in ("$O/b2M_$name.json","$O/b6M_$name.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['roofline']['whole_step_frac'], d['kernels']['attention']['avg_ms'], d['kernels']['post_attn_fused']['avg_ms'], d['clocks']['sm_mhz'])
    except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
}
run oneshot MAPF_GPT_B200_POST_PERSIST=0
run default X=1
run stag0 MAPF_GPT_B200_POST_STAGGER_NS=0
run stag11 MAPF_GPT_B200_POST_STAGGER_NS=11000
run stag44 MAPF_GPT_B200_POST_STAGGER_NS=44000
run stag88 MAPF_GPT_B200_POST_STAGGER_NS=88000
run default2 X=1
