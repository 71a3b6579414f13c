mkdir -p gpurun_out
for c in 2 3 4 5; do
  ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_elapsed.avg,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_render_tc3 -s 3 -c 2 --csv python bench.py --config $c --steps 2 --warmup 3 2>/dev/null | grep -E '^"[0-9]' | awk -F'","' -v c=$c '{gsub(/"/,"",$NF); print "config" c "," $5 "," $(NF-2) "," $(NF-1) "," $NF}' 
done > gpurun_out/r02b_configs_ncu.csv
cat gpurun_out/r02b_configs_ncu.csv | cut -c1-220
