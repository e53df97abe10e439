#!/bin/bash
# A/B on the GPU box: post_attn_kernel as CTA pairs (MAPF_GPT_B200_POST_CL=2) vs single CTAs: parity first, then bench --quick.
set -u
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith('{"metric"'):
        d = json.loads(ln)
        print(sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "clk", d["clocks"].get("sm_mhz"),
              {k: v["avg_ms"] for k, v in d["kernels"].items() if v["share"] > 0.01})
PY
}
for cl in ${CLS_TEST:-1 2}; do
  MAPF_GPT_B200_POST_CL=$cl timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu \
     -k "forward or fused_and_generic" > gpurun_out/ab_cl${cl}_tests.log 2>&1
  echo "CL=$cl parity exit $?"; tail -5 gpurun_out/ab_cl${cl}_tests.log
done
for cl in ${CLS_BENCH:-1 2 1 2}; do
  MAPF_GPT_B200_POST_CL=$cl timeout 300 python bench.py --quick --steps 4 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/ab_cl${cl}_bench.log 2>&1
  echo "CL=$cl bench exit $?"; summ gpurun_out/ab_cl${cl}_bench.log
done
