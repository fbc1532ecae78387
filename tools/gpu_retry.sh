#!/bin/bash
# tools/gpu_retry.sh <timeout-seconds> <command...>: gpurun with retries while the pod answers "busy" (exit 3)
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
