#!/bin/bash
# usage: scratch/gpu_retry.sh <timeout> <command...> — retries gpurun while the pod answers busy (exit code 3)
t=$1; shift
for i in $(seq 1 30); do
  gpurun --timeout $t -- "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
