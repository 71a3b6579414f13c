#!/bin/bash
# Profiling recipe (run under gpurun, 1 GPU). Outputs land in gpurun_out/.
#   bash profiles/run_profile.sh <tag> [kernel-regex]
set -u
TAG=${1:-r01}
KREGEX=${2:-k_render_tc}
mkdir -p gpurun_out
# 1) every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
# 2) the dominant kernel, full set, one launch after warm-up
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 3 -c 1 -o gpurun_out/${TAG}_render_tc \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full_bench.log 2>&1
ls -la gpurun_out | tail -5
