#!/bin/bash
# A/B of two BUILDS of libposeengine.so on one box (run under gpurun): bench.py fps for the baseline build given as $1
# (e.g. caffe_rtpose_b200/build/ab_base/libposeengine.so, built from another commit) and for the in-tree build, alternating.
#   gpurun --timeout 600 -- 'tools/ab_lib.sh caffe_rtpose_b200/build/ab_base/libposeengine.so 2'
base=$1; reps=${2:-1}
mkdir -p gpurun_out
for r in $(seq 1 $reps); do
for v in base new; do
  if [ $v = base ]; then export PE_LIB=$PWD/$base; else unset PE_LIB; fi
  echo "== $v (rep $r)"
  timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | tee "gpurun_out/ablib_${v}_${r}.json" | python -c '
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d["roofline"]
print("fps %.1f  e2e %.1f  ms/step %.3f  conv %.3f ms/step  frac %.3f  sm %s MHz %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"],
      r["kernel_ms_per_step"], r["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))'
done
done
