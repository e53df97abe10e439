#!/bin/bash
# last evidence pass of the round (HEAD): whole GPU suite, smoke, default bench (timed), reference arm
cd "$(dirname "$0")/../.."
O=gpurun_out/r02ay; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time timeout 900 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -4 $O/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 8 --warmup 3 ) > $O/bench_ref.json 2> $O/bench_ref.err; tail -4 $O/bench_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ay/bench_default.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'frac',d['roofline']['whole_step_frac'], d['roofline']['kernel'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])
for k,v in d.get('other_configs',{}).items():
    print(k, {kk:(round(vv) if isinstance(vv,float) and vv>100 else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')}, v.get('roofline',{}).get('whole_step_frac'), v.get('e2e',{}) and round(v['e2e']['value']))
print(json.dumps(d['other_configs'].get('C1_act_latency_random_32x1_2M'))[:400])
for k in ('cpu_baseline','cpu_as_shipped','stock_gpu'): print(k, json.dumps(d.get(k))[:200])
PY
