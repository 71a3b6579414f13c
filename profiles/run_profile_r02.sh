#!/bin/bash
# Round-2 profiling recipe (run under gpurun, 1 GPU): launch list of the bench incl. the training leg, then `ncu --set full` of
# one launch each of the render kernel and of the two tcgen05 training kernels.  Outputs in gpurun_out/; summaries -> profiles/.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-eager-gpu > gpurun_out/${TAG}_launches_bench.log 2>&1
for K in k_render_tc3 k_bwd_chain k_bwd_dw; do
  SKIP=3; [ $K != k_render_tc3 ] && SKIP=2
  ncu --set full --clock-control none --import-source on -k regex:${K} -s $SKIP -c 1 -f -o gpurun_out/${TAG}_${K} \
      python profiles/train_breakdown.py --no-profiler --render > gpurun_out/${TAG}_${K}_run.log 2>&1
  ncu -i gpurun_out/${TAG}_${K}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${K}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_${K}.ncu-rep --page source --csv > gpurun_out/${TAG}_${K}_source.csv 2>/dev/null
  python profiles/ncu_raw_to_summary.py gpurun_out/${TAG}_${K}_raw.csv > gpurun_out/${TAG}_${K}_ncu_full_summary.csv 2>/dev/null
  python profiles/ncu_source_summary.py gpurun_out/${TAG}_${K}_source.csv > gpurun_out/${TAG}_${K}_stalls.txt 2>&1
done
ls -la gpurun_out | grep ${TAG} | tail -20
