#!/bin/bash
# usage: gpurun_retry.sh LOGFILE [gpurun args...]   -- retries while the pod has no free slot (exit code 3)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
