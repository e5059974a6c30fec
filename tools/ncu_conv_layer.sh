#!/bin/bash
# One `ncu --set full` capture (source-level, needs the -lineinfo build) of ONE conv launch of a bench step, run under
# gpurun on one GPU.  $1 = index of the conv launch inside the second forward (0 = conv1_1 ... 91 = Mconv7_stage6_L2 in
# execution order = prototxt order, L1 / L2 interleaved; conv3_2 = 5, Mconv1_stage2_L1 = 22, Mconv2_stage2_L1 = 24), default 5.
# (tools/ncu_round.sh takes the round's whole evidence set with a smaller target.)  Graphs are disabled so that -s counts kernel launches.
#   gpurun --timeout 400 -- 'tools/ncu_conv_layer.sh 5'
# Read back here with:  python tools/summarize_ncu.py full gpurun_out/prof_conv_$1.ncu-rep profiles/<round>_conv_full.md
#                       ncu -i gpurun_out/prof_conv_$1.ncu-rep --page source --csv > gpurun_out/prof_conv_$1_source.csv
idx=${1:-5}
mkdir -p gpurun_out
PE_GRAPH=0 PE_TWO_LANES=0 timeout 380 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s $((92 * 5 + idx)) -c 1 \
  -f -o gpurun_out/prof_conv_${idx} python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_conv_${idx}.log 2>&1
tail -3 gpurun_out/prof_conv_${idx}.log
