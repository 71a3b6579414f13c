mkdir -p gpurun_out
python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 300 gpurun_out/r02c_bench.json
python bench.py --impl reference > gpurun_out/r02c_bench_reference.json 2>/dev/null
for c in 2 3 4 5; do python bench.py --config $c --steps 5 --warmup 3 2>/dev/null | tail -1; done > gpurun_out/r02c_configs.json; cut -c1-160 gpurun_out/r02c_configs.json
python profiles/configs_bench.py 2>/dev/null > gpurun_out/r02c_configs_bench.json; grep -E "Positional|Mip" gpurun_out/r02c_configs_bench.json | cut -c1-160
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-eager-gpu > /dev/null 2>&1
for c in 3; do
  ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_elapsed.avg --clock-control none -k regex:k_render_tc3 -s 3 -c 1 --csv python bench.py --config $c --steps 2 --warmup 3 2>/dev/null | grep -E '^"[0-9]' | awk -F'","' -v c=$c '{gsub(/"/,"",$NF); print "config" c "," $5 "," $(NF-2) "," $(NF-1) "," $NF}'
done > gpurun_out/r02c_mip_ncu.csv
ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_elapsed.avg --clock-control none -k regex:k_render_tc3 -s 2 -c 1 --csv python profiles/mip_run.py pos 2>/dev/null | grep -E '^"[0-9]' | awk -F'","' '{gsub(/"/,"",$NF); print "positional," $5 "," $(NF-2) "," $(NF-1) "," $NF}' >> gpurun_out/r02c_mip_ncu.csv
cat gpurun_out/r02c_mip_ncu.csv | cut -c1-200
