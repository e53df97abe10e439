#!/bin/bash
# A/B: sequences per forward chunk
cd "$(dirname "$0")/../.."
O=gpurun_out/r02q; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 8 --warmup 3 > $O/b2M_$name.json 2>$O/b2M_$name.err
  env "$@" timeout 300 python bench.py --quick --steps 4 --warmup 3 --model 6M --map wfi_warehouse --agents 192 --envs 512 > $O/b6M_$name.json 2>$O/b6M_$name.err
  python - <<PY
import json
for f in ("$O/b2M_$name.json","$O/b6M_$name.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['roofline']['whole_step_frac'], {k:(v['avg_ms'],v['launches']) for k,v in d['kernels'].items() if v['share']>0.015}, d['clocks']['sm_mhz'])
    except Exception as ex: print(f,'ERR',ex, open(f.replace('.json','.err')).read()[-800:])
PY
}
run c8192 X=1
run c16384 MAPF_GPT_B200_CHUNK_SEQS=16384
run c32768 MAPF_GPT_B200_CHUNK_SEQS=32768
run c4096 MAPF_GPT_B200_CHUNK_SEQS=4096
run c8192b X=1
