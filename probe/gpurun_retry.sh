#!/bin/bash
# usage: probe/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers "busy", exit code 3)
t=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $t "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
