#!/bin/bash
# gpurun with retries on "no box / slot free right now" (exit 3): tools/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
