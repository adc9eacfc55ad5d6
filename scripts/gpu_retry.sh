#!/bin/bash
# usage: scripts/gpu_retry.sh LOGFILE TIMEOUT 'command' [gpurun extra args]  -- retries while the pool answers busy (exit 3)
log=$1; to=$2; cmd=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout $to -- "$cmd" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 45
done
exit 3
