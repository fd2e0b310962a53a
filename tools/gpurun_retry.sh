#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> <command...>   -- retries while the pod answers "busy" (exit code 3)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc" >> $LOG; exit $rc; fi
  sleep 45
done
echo "gave up" >> $LOG
