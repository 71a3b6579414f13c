python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 200 gpurun_out/r02c_bench.json
python bench.py --impl reference > gpurun_out/r02c_bench_reference.json 2>/dev/null
bash profiles/run_profile_r02.sh r02c > gpurun_out/r02c_profile.log 2>&1; tail -3 gpurun_out/r02c_profile.log
