#!/bin/bash
# 8-GPU pass, launched as the driver does: the default bench (C2 headline + every other config; its C4 entry IS BASELINE's C4:
# Berlin tile, 256 agents x 32 envs per GPU x 8 GPUs, MAPF-GPT-85M) and the reference arm
cd "$(dirname "$0")/../.."
O=gpurun_out/r02as; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 8 --warmup 3 ) > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"; tail -4 $O/bench_n8.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 4 --warmup 1 ) > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err; echo "ref n8 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r02as/bench_n8.json','gpurun_out/r02as/bench_ref_n8.json'):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, d.get('impl'), round(d['value']), d['n_gpus'], d.get('e2e',{}).get('value'), d.get('clocks'))
        for k,v in d.get('other_configs',{}).items():
            print('  ',k, {kk:(round(vv) if isinstance(vv,float) and vv>100 else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')}, v.get('roofline',{}).get('whole_step_frac'))
    except Exception as ex: print(f,'ERR',ex)
PY
