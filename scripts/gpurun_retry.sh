#!/bin/bash
# gpurun with retries on "no box / slot free right now" (exit code 3: nothing charged).
#   scripts/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun args...]
t=$1; shift; cmd=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@" -- "$cmd"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
