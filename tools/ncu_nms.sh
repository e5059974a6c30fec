#!/bin/bash
# --set full capture of one nms_flags_kernel launch (resize x8 + multi-scale average + 3x3 NMS flags, post.cu) of a 9-frame C2 forward.
tag=${1:-r2}
mkdir -p gpurun_out
NT_FORWARDS=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:nms_flags -s 1 -c 1 -f -o gpurun_out/${tag}_nms_flags \
    python tools/ncu_target.py > gpurun_out/${tag}_nms_full.log 2>&1
echo "nms full capture rc=$?"
