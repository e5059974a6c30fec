#!/bin/bash
# Build-container helper: run one gpurun call, retrying while the pod answers "busy" (exit code 3, nothing charged).
#   tools/gpurun_retry.sh <log file> <timeout seconds> '<command>'
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)" >> "$log"; exit $rc; fi
  sleep 90
done
