python profiles/ab_time.py libnerf_b200.so libnerf_b200_loop.so libnerf_b200.so
NF_LIB=libnerf_b200_stats.so timeout 300 python profiles/stats_run.py 3 2>&1 | tail -8
ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_elapsed.avg --clock-control none -k regex:k_render_tc3 -s 2 -c 1 --csv python profiles/ab_time.py --one 2 /tmp/x.pt 2>&1 | tail -3 | cut -d, -f13-
