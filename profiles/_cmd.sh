timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in libnerf_b200.so libnerf_b200_posold.so libnerf_b200.so; do NF_LIB=$lib python profiles/mip_run.py pos 2>&1 | tail -1; done
