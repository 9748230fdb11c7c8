#!/bin/bash
# development aid: repeat a gpurun call until the pod has a slot (exit code 3 = nothing charged, retry)
# usage: scripts/gpu_retry.sh <timeout-seconds> '<command>'
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
