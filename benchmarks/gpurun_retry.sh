#!/bin/bash
# usage: benchmarks/gpurun_retry.sh <timeout_s> <log> <command...>   (retries while the pod answers "busy", rc 3)
T=$1; LOG=$2; shift 2
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
