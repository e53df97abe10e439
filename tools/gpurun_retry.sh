#!/bin/bash
# usage: gpurun_retry.sh <logfile> <timeout> <command...>
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
