#!/bin/bash
# Profiling recipe (run under gpurun, 1 GPU). Outputs land in gpurun_out/; summaries are then copied into profiles/.
#   bash profiles/run_profile.sh <tag> [kernel-regex]
set -u
TAG=${1:-r01}
KREGEX=${2:-k_render_tc3}
mkdir -p gpurun_out
# 1) every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
# 2) the dominant kernel, full set, one launch after warm-up, with source-level sampling
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 3 -c 1 -f -o gpurun_out/${TAG}_render \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full_bench.log 2>&1
ncu -i gpurun_out/${TAG}_render.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_render.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
python profiles/ncu_source_summary.py gpurun_out/${TAG}_source.csv > gpurun_out/${TAG}_stalls.txt 2>&1
ls -la gpurun_out | tail -8
