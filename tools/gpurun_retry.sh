#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-s> '<command>'   -- retries while the pod answers "busy" (exit code 3)
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
