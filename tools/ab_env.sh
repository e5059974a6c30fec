#!/bin/bash
# A/B of environment knobs on ONE box (run under gpurun), bench.py only (tools/ab_bench.sh also checks the accuracy of conv knobs):
#   gpurun --timeout 600 -- 'tools/ab_env.sh 2 PE_BENCH_HANDLES=2 PE_BENCH_HANDLES=3'      # 2 alternating repetitions
reps=$1; shift
mkdir -p gpurun_out
for r in $(seq 1 $reps); do
for v in "$@"; do
  echo "== $v (rep $r)"
  env ${v//,/ } timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | tee "gpurun_out/abenv_${v//[^A-Za-z0-9_=]/_}_${r}.json" | python -c '
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d["roofline"]
print("fps %.1f  e2e %.1f  ms/step %.3f  conv %.3f ms/step  frac %.3f  sm %s MHz %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"],
      r["kernel_ms_per_step"], r["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))'
done
done
