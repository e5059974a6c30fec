#!/bin/bash
# Round-end validation of the committed build on one box (run under gpurun): $1 = tag (e.g. r2r).
#   full GPU test suite, smoke(), bench.py (+ the reference arm), ncu launch list of a bench step, dram bytes / tensor-pipe % of the
#   92 conv launches of one forward, --set full capture of nms_flags_kernel.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/${tag}_smoke.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
NT_FORWARDS=3 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:conv_tc -s 184 -c 92 --csv --log-file gpurun_out/${tag}_conv_traffic.csv python tools/ncu_target.py > gpurun_out/${tag}_traffic.log 2>&1; echo "traffic rc=$?"
tools/ncu_nms.sh $tag
