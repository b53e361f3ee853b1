#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'
# Retries while gpurun answers 3 (no box / slot free right now, nothing charged).
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
