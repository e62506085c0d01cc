#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <logfile> <command...>   -- retries while the pod answers busy (rc 3)
to=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
