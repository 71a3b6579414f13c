timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
NF_LIB=libnerf_b200_exp.so timeout 600 python -m pytest tests -m gpu -x -q -k "variants or dnerf" 2>&1 | tail -3
