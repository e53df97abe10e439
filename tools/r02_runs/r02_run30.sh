#!/bin/bash
# 2-GPU pass: bench.py under torchrun exactly as the driver launches it (+ the reference arm), the two-devices-in-one-process test,
# benchmark.py sharded over 2 ranks
cd "$(dirname "$0")/../.."
O=gpurun_out/r02ai; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "two_devices or pre_tokenized or dropin or lanes" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
timeout 600 python -m pytest tests/test_gpu_rollout.py -m gpu -q -k "dropin" >> $O/tests.log 2>&1; echo "tests2 rc=$?"; tail -3 $O/tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 8 --warmup 3 ) > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -5 $O/bench_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 ) > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?"
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 benchmark.py --limit 32 --sets 02-mazes 04-movingai --algorithms MAPF-GPT-2M --out $O/eval_results ) > $O/benchmark_n2.log 2>&1; echo "benchmark n2 rc=$?"; grep "^# \|total" $O/benchmark_n2.log | cut -c1-160
python - <<'PY'
import json
for f in ('gpurun_out/r02ai/bench_n2.json','gpurun_out/r02ai/bench_ref_n2.json'):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, d.get('impl'), round(d['value']), d['n_gpus'], d.get('e2e',{}).get('value'))
        for k,v in d.get('other_configs',{}).items():
            print('  ',k, {kk:(round(vv) if isinstance(vv,float) and vv>100 else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')})
    except Exception as ex: print(f,'ERR',ex)
PY
