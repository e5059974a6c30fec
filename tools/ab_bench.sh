#!/bin/bash
# A/B of conv-kernel knobs on ONE box (run under gpurun): accuracy at 160x96 (tools/gpu_diag.py tc) and bench.py fps for
# every "ENV=VALUE[,ENV2=VALUE2...]" variant given on the command line.  Example (≈2 GPU-minutes per variant):
#   gpurun --timeout 600 -- 'tools/ab_bench.sh PE_TC_CHUNK_MUL=1 PE_TC_CHUNK=0 PE_TC_CHUNK_MUL=2'
# Knobs: PE_TC_CHUNK (0 = one TMEM chain per tile, n = n steps per chunk), PE_TC_CHUNK_MUL (multiplier of the default
# chunk sizes), PE_TC_PAIR (1 = CTA-pair kernel, cta_group::2), PE_TC_PDL (0 = no programmatic dependent launch), PE_TC_NARROW (0 = never pick 64/32-wide tiles), PE_TC_VARIANT (0 = baseline one-tile-per-tap kernel),
# PE_GRAPH (0 = no CUDA graphs).
export DIAG_PRECS=${DIAG_PRECS:-2}
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"
  env ${v//,/ } timeout 120 python tools/gpu_diag.py tc 2>&1 | grep -E "stride-8"
  env ${v//,/ } timeout 150 python bench.py --no-cpu-baseline 2>/dev/null | tee "gpurun_out/ab_${v//[^A-Za-z0-9_=]/_}.json" | python -c '
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d["roofline"]
print("fps %.1f  e2e %.1f  conv %.3f ms/step  %.1f TFLOP/s  frac %.3f  sm %s MHz %s" % (d["value"], d["e2e"]["value"],
      r["kernel_ms_per_step"], r["achieved"], r["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))'
done
