#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod has no free slot (nothing is charged then)
LOG=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > $LOG 2>&1
  rc=$?
  if ! grep -q "status=transient" $LOG && [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
