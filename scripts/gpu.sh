#!/bin/bash
# usage: scripts/gpu.sh <timeout_s> <log> <command...>   - retries while the pod has no free slot (rc 3)
# or while an earlier call of this repo is still draining
T=$1; LOG=$2; shift 2
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -eq 3 ]; then sleep 60; continue; fi
  if [ $rc -eq 2 ] && grep -q "already running" $LOG; then sleep 30; continue; fi
  exit $rc
done
exit 3
