python profiles/bw_diff.py 400 2>&1 | tail -2 | tee gpurun_out/r02c_bw_diff.json
