#!/bin/bash
# End-to-end run of rtpose.bin on the GPU box with REAL decode: writes N synthetic 1280x720 JPEGs (cv2, quality 90) to /tmp, then
# processes them with the reference's command line.  $1 = number of GPUs (default 1), $2 = frames (default 720), $3 = producers.
#   gpurun --timeout 600 -- 'tools/e2e_cli.sh 1 720 32'
ng=${1:-1}; n=${2:-720}; prod=${3:-32}
dir=/tmp/e2e_frames
python - <<PY
import os, sys, cv2
sys.path.insert(0, ".")
from caffe_rtpose_b200 import synth
os.makedirs("$dir", exist_ok=True)
base = [synth.make_frame(i) for i in range(24)]
for i in range($n):
    cv2.imwrite("$dir/f%05d.jpg" % i, base[i % 24], [cv2.IMWRITE_JPEG_QUALITY, 90])
print("wrote $n jpgs")
PY
for mode in "--decode_bench" ""; do
  echo "== rtpose.bin $mode num_gpu=$ng producers=$prod"
  timeout 300 caffe_rtpose_b200/rtpose.bin --image_dir $dir --random_init he --model COCO --no_display --no_frame_drops --num_gpu $ng \
      --num_producers $prod --write_json /tmp/e2e_json $mode 2>&1 | grep -E "FPS|frames/s|Done|decoded|ready|broadcast|rror" | tail -6
done
