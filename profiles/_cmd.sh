for lib in libnerf_b200.so libnerf_b200_nst2.so libnerf_b200.so libnerf_b200_nst2.so; do NF_LIB=$lib timeout 120 python profiles/mip_run.py 2>&1 | tail -1; done
NF_LIB=libnerf_b200_nst2.so timeout 600 python -m pytest tests -m gpu -x -q -k "mip" 2>&1 | tail -2
