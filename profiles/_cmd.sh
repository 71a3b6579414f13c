timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-eager-gpu > gpurun_out/r02d_launches_bench.log 2>&1
tail -c 300 gpurun_out/r02d_launches_bench.log; wc -l gpurun_out/r02d_launches.csv
