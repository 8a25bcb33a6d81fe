#!/bin/bash
# tools/gpurun_retry.sh <timeout-seconds> '<command>' — retries while the pod answers "no box free" (exit code 3)
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
