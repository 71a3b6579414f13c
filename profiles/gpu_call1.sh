#!/bin/bash
# GPU call: per-pipeline parity (separate processes: a trapped launch poisons its CUDA context), full suite, variant timing.
mkdir -p gpurun_out
for v in pipe3_ring6 pipe3_ring3 pipe2 pipe1; do
  timeout 300 python -m pytest tests -m gpu -q -x --tb=short -k "variants and $v" > gpurun_out/pytest_$v.log 2>&1
  echo "$v exit $?" >> gpurun_out/summary.txt
done
timeout 600 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_all.log 2>&1
echo "all exit $?" >> gpurun_out/summary.txt
timeout 300 python profiles/perf_variants.py --steps 4 > gpurun_out/variants.json 2> gpurun_out/variants.err
echo "variants exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/pytest_all.log; cat gpurun_out/variants.json
