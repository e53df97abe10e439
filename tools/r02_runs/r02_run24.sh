#!/bin/bash
# last-block pruning on the generic (85M) path: tests + A/B (MAPF_GPT_B200_NO_PRUNE=1 = every block in full)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02aa; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py tests/test_gpu_rollout.py -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log
python tools/flip_rate.py 85M > $O/flip_85M.txt 2>&1; cut -c1-260 $O/flip_85M.txt
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 3 --warmup 3 --model 85M --map Berlin_1_256_05 --agents 256 --envs 32 > $O/b85M_$name.json 2>$O/b85M_$name.err
  python - <<PY
import json
for f in ("$O/b85M_$name.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['roofline']['whole_step_frac'], {k:(v['avg_ms'],v['launches']) for k,v in d['kernels'].items() if v['share']>0.005}, d['clocks']['sm_mhz'])
    except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
}
run pruned X=1
run full MAPF_GPT_B200_NO_PRUNE=1
run pruned2 X=1
run full2 MAPF_GPT_B200_NO_PRUNE=1
