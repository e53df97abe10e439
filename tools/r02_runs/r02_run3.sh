#!/bin/bash
# A/B of the persistent post_attn launch (MAPF_GPT_B200_POST_PERSIST=0 = one CTA per tile group, as before)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02c; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rollout.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
for mode in persist oneshot persist oneshot; do
  if [ $mode = oneshot ]; then export MAPF_GPT_B200_POST_PERSIST=0; else unset MAPF_GPT_B200_POST_PERSIST; fi
  timeout 300 python bench.py --quick --steps 8 --warmup 3 > $O/b2M_$mode.json 2>$O/b2M_$mode.err
  timeout 300 python bench.py --quick --steps 4 --warmup 3 --model 6M --map wfi_warehouse --agents 192 --envs 512 > $O/b6M_$mode.json 2>$O/b6M_$mode.err
  python - <<PY
import json
for f in ("$O/b2M_$mode.json","$O/b6M_$mode.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['roofline']['whole_step_frac'], d['kernels']['attention']['avg_ms'], d['kernels']['post_attn_fused']['avg_ms'], d['clocks']['sm_mhz'])
    except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
done
