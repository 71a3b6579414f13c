timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 600 gpurun_out/r02b_bench.json
python bench.py --impl reference > gpurun_out/r02b_bench_reference.json 2>/dev/null
bash profiles/run_profile_r02.sh r02b > gpurun_out/r02b_profile.log 2>&1; tail -5 gpurun_out/r02b_profile.log
