#!/bin/bash
# round-2 first GPU pass: tests, then A/B of the max-free softmax on the three model sizes
mkdir -p gpurun_out/r02a
cd "$(dirname "$0")/../.."
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02a/tests.log
tail -15 gpurun_out/r02a/tests.log
for mode in fast safe; do
  if [ $mode = safe ]; then export MAPF_GPT_B200_SAFE_SOFTMAX=1; else unset MAPF_GPT_B200_SAFE_SOFTMAX; fi
  timeout 300 python bench.py --quick --steps 8 --warmup 3 > gpurun_out/r02a/bench_2M_$mode.json 2> gpurun_out/r02a/bench_2M_$mode.err
  timeout 300 python bench.py --quick --steps 4 --warmup 3 --model 6M --map wfi_warehouse --agents 192 --envs 512 > gpurun_out/r02a/bench_6M_$mode.json 2> gpurun_out/r02a/bench_6M_$mode.err
  timeout 300 python bench.py --quick --steps 3 --warmup 3 --model 85M --map Berlin_1_256_05 --agents 256 --envs 32 > gpurun_out/r02a/bench_85M_$mode.json 2> gpurun_out/r02a/bench_85M_$mode.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02a/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f, round(d['value']), d['roofline']['whole_step_frac'], {n:(v['avg_ms'],v['share']) for n,v in k.items() if v['share']>0.02}, d['clocks']['sm_mhz'])
    except Exception as ex:
        print(f, 'ERR', ex)
PY
