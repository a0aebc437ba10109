#!/bin/bash
# gpurun with retries on "no box free right now" (exit code 3 / transient): usage gpurun_retry.sh <timeout> '<command>' <logfile>
t=$1; cmd=$2; log=$3
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $t -- "$cmd" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
exit $rc
