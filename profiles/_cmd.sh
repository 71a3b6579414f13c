timeout 600 python -m pytest tests -m gpu -x -q -k "mip" 2>&1 | tail -6
for lib in libnerf_b200.so libnerf_b200_nowb.so libnerf_b200.so; do NF_LIB=$lib timeout 120 python profiles/mip_run.py 2>&1 | tail -1; done
