#!/bin/bash
# Round evidence under ncu (run under gpurun, one GPU): $1 = tag (e.g. r2k).
#   1. launch list of one full bench step (all kernels, gpu__time_duration)        -> gpurun_out/<tag>_launches.csv
#   2. dram bytes / duration / tensor-pipe % of the 92 conv launches of one forward -> gpurun_out/<tag>_conv_traffic.csv
#   3. --set full + source of one 7x7 128->128 layer (Mconv2_stage2_L1, conv #24) and one 3x3 256->256 layer (conv3_2, conv #5)
tag=${1:-r2}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
NT_FORWARDS=3 timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:conv_tc -s 184 -c 92 --csv --log-file gpurun_out/${tag}_conv_traffic.csv python tools/ncu_target.py > gpurun_out/${tag}_traffic.log 2>&1
echo "traffic rc=$?"
for idx in 24 5; do
  NT_FORWARDS=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s $((184 + idx)) -c 1 -f -o gpurun_out/${tag}_conv_${idx} \
      python tools/ncu_target.py > gpurun_out/${tag}_full_${idx}.log 2>&1
  echo "full capture conv #$idx rc=$?"
done
