#!/bin/bash
# A/B: MUFU.TANH GELU vs the FMA-only polynomial; default persistent policy (C=256 only)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02g; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rollout.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
MAPF_GPT_B200_LIB_PATH=$PWD/mapf_gpt_b200/libvar_gelu_tanh.so timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -m gpu -q -k "forward or logit or fused" > $O/tests_tanh.log 2>&1; echo "tanh tests rc=$?"; tail -3 $O/tests_tanh.log
python tools/flip_rate.py 2M 6M > $O/flip_base.txt 2>&1; cat $O/flip_base.txt
MAPF_GPT_B200_LIB_PATH=$PWD/mapf_gpt_b200/libvar_gelu_tanh.so python tools/flip_rate.py 2M 6M > $O/flip_tanh.txt 2>&1; cat $O/flip_tanh.txt
run() { name=$1; shift
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
run base X=1
run tanh MAPF_GPT_B200_LIB_PATH=$PWD/mapf_gpt_b200/libvar_gelu_tanh.so
run base2 X=1
run tanh2 MAPF_GPT_B200_LIB_PATH=$PWD/mapf_gpt_b200/libvar_gelu_tanh.so
