#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gpurun_retry.sh <timeout_s> '<command>' [extra gpurun args]
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
