timeout 300 python profiles/ab_time.py libnerf_b200.so libnerf_b200_s4.so libnerf_b200.so libnerf_b200_s4.so 2>&1 | tail -5
NF_LIB=libnerf_b200_s4.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "reference_golden or ragged" 2>&1 | tail -4
