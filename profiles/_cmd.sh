timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 200 gpurun_out/r02d_bench.err
