#!/bin/bash
# round-2 second GPU pass: all GPU tests, the full default bench (every config + comparators), the reference arm,
# the YAML benchmark smoke, and the attention polynomial-share A/B
cd "$(dirname "$0")/../.."
O=gpurun_out/r02b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" | tee -a $O/tests.log
tail -5 $O/tests.log
( time timeout 900 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
( time timeout 600 python bench.py --impl reference --steps 8 --warmup 3 ) > $O/bench_ref.json 2> $O/bench_ref.err
( time timeout 600 python benchmark.py --limit 16 --out $O/eval_results ) > $O/benchmark_limit16.log 2>&1; echo "benchmark rc=$?"
for v in 0x00 0x88 base 0xAA 0xEE; do
  if [ $v = base ]; then unset MAPF_GPT_B200_LIB_PATH; else export MAPF_GPT_B200_LIB_PATH=$PWD/mapf_gpt_b200/libvar_poly_$v.so; fi
  timeout 300 python bench.py --quick --steps 8 --warmup 3 > $O/poly_$v.json 2>/dev/null
done
unset MAPF_GPT_B200_LIB_PATH
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b/poly_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['kernels']['attention']['avg_ms'], d['kernels']['post_attn_fused']['avg_ms'], d['clocks']['sm_mhz'])
    except Exception as ex: print(f,'ERR',ex)
d=json.loads(open('gpurun_out/r02b/bench_default.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'frac',d['roofline']['whole_step_frac'])
for k,v in d.get('other_configs',{}).items():
    print(k, {kk:(round(vv) if isinstance(vv,float) and vv>100 else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','error')}, v.get('roofline',{}).get('whole_step_frac'), v.get('e2e',{}) and round(v['e2e']['value']))
print(json.dumps(d['other_configs'].get('C1_act_latency_random_32x1_2M'))[:900])
for k in ('cpu_baseline','cpu_as_shipped','stock_gpu'): print(k, json.dumps(d.get(k))[:700])
PY
tail -3 $O/bench_default.err; tail -30 $O/benchmark_limit16.log
