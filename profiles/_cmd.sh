timeout 900 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 600 gpurun_out/r02d_bench.err
python __graft_entry__.py smoke 2>&1 | tail -3 || true
