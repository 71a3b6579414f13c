#!/bin/bash
# usage: profiles/grun.sh <logfile> <timeout_s> '<command>'   -- gpurun with retry while the pod has no free slot
log=$1; to=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|no box\|busy" $log; then sleep 45; continue; fi
  break
done
