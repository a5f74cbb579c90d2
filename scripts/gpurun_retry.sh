#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <log> <command...>   — retries while gpurun answers "busy" (exit 3)
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
