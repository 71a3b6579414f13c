#!/bin/bash
# usage: profiles/grun2.sh <gpus> <logfile> <timeout_s> '<command>'
g=$1; log=$2; to=$3; shift 3
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --gpus $g --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|no box\|busy" $log; then sleep 60; continue; fi
  break
done
