NF_LIB=libnerf_b200_stats.so timeout 300 python profiles/stats_run.py 3 2>&1 | tail -14
timeout 200 python profiles/ab_time.py libnerf_b200.so 2>&1 | tail -2
