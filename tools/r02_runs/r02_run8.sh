#!/bin/bash
# GELU forms A/B (logistic default, tanh, FMA-only polynomial) + whole GPU suite + flip rates + the full YAML benchmark
cd "$(dirname "$0")/../.."
O=gpurun_out/r02h; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
python tools/flip_rate.py > $O/flip_rate.txt 2>&1; cat $O/flip_rate.txt | cut -c1-250
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
run logistic X=1
run poly MAPF_GPT_B200_LIB_PATH=$PWD/mapf_gpt_b200/libvar_gelu_poly.so
run tanh MAPF_GPT_B200_LIB_PATH=$PWD/mapf_gpt_b200/libvar_gelu_tanh.so
run logistic2 X=1
( time timeout 1500 python benchmark.py --out $O/eval_results ) > $O/benchmark_full.log 2>&1; echo "benchmark rc=$?"; grep "^# \|total_episodes\|real" $O/benchmark_full.log | cut -c1-200
