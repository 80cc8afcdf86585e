#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout> <logfile> <command...>  -- retries while the pod answers "busy"
# (transient) or while an earlier call of this repo is still registered as running
T=$1; LOG=$2; shift 2
for i in $(seq 1 200); do
  /usr/local/graft/bin/gpurun --gpus ${GPUS:-1} --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|already running" $LOG; then sleep 20; continue; fi
  exit $rc
done
exit 3
