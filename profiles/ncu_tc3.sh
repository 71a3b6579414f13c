#!/bin/bash
# ncu --set full capture of one k_render_tc3 launch (after warm-up) + source-level stall sampling
mkdir -p gpurun_out
TAG=${1:-r01_tc3}
ncu --set full --clock-control none --import-source on -k regex:k_render_tc3 -s 3 -c 1 -f -o gpurun_out/${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ls -la gpurun_out | tail -6
