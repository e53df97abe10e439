#!/bin/bash
# A/B: key halves of S / P pipelined against the softmax in the max-free attention kernels (libnext.so = new)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02k; mkdir -p $O
NEW=$PWD/mapf_gpt_b200/libnext.so
MAPF_GPT_B200_LIB_PATH=$NEW timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py tests/test_gpu_rollout.py -m gpu -x -q > $O/tests_new.log 2>&1; echo "tests(new) rc=$?"; tail -4 $O/tests_new.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 8 --warmup 3 > $O/b2M_$name.json 2>$O/b2M_$name.err
  env "$@" timeout 300 python bench.py --quick --steps 4 --warmup 3 --model 6M --map wfi_warehouse --agents 192 --envs 512 > $O/b6M_$name.json 2>$O/b6M_$name.err
  env "$@" timeout 300 python bench.py --quick --steps 3 --warmup 3 --model 85M --map Berlin_1_256_05 --agents 256 --envs 32 > $O/b85M_$name.json 2>$O/b85M_$name.err
  python - <<PY
import json
for f in ("$O/b2M_$name.json","$O/b6M_$name.json","$O/b85M_$name.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['roofline']['whole_step_frac'], d['kernels']['attention']['avg_ms'], d['clocks']['sm_mhz'])
    except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
}
run old X=1
run new MAPF_GPT_B200_LIB_PATH=$NEW
run old2 X=1
run new2 MAPF_GPT_B200_LIB_PATH=$NEW
