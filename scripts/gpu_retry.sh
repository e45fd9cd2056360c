#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout_s> <command string> ; retries while the pod answers "busy" (exit 3)
t=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $t -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 90
done
tail -120 /tmp/gpurun_last.log
exit $rc
