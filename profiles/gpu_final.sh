#!/bin/bash
# Round-end style validation in one GPU call: gpu tests, smoke, the bench line, the reference arm, profiles.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short > gpurun_out/final_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/final_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --torch-eager-gpu > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"
cat gpurun_out/final_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null; cat gpurun_out/final_bench_reference.json
bash profiles/run_profile.sh ${1:-r01b} > gpurun_out/final_profile.log 2>&1; tail -3 gpurun_out/final_profile.log
