#!/bin/bash
# generic A/B on the GPU box: each argument is "NAME=VALUE[,NAME=VALUE...]" (or "base"); runs bench.py --quick per setting
set -u
mkdir -p gpurun_out
summ() { python - "$1" "$2" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith('{"metric"'):
        d = json.loads(ln)
        print(sys.argv[2], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "clk", d["clocks"].get("sm_mhz"),
              {k: v["avg_ms"] for k, v in d["kernels"].items() if v["share"] > 0.01})
PY
}
i=0
for setting in "$@"; do
  i=$((i+1))
  envs=""
  if [ "$setting" != "base" ]; then envs=$(echo "$setting" | tr ',' ' '); fi
  env $envs timeout 300 python bench.py --quick --steps 4 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/ab_env_$i.log 2>&1
  rc=$?
  if [ $rc -ne 0 ]; then echo "$setting: exit $rc"; tail -5 gpurun_out/ab_env_$i.log; else summ gpurun_out/ab_env_$i.log "$setting"; fi
done
