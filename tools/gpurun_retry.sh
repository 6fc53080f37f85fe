#!/bin/bash
# usage: [GPUS=2] tools/gpurun_retry.sh <timeout-s> '<command>'   -- retries while the pod answers "busy" (exit code 3)
extra=""
if [ -n "$GPUS" ] && [ "$GPUS" != "1" ]; then extra="--gpus $GPUS"; fi
for attempt in $(seq 1 60); do
  /usr/local/graft/bin/gpurun $extra --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
