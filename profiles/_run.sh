timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -E "nf_tc|FAILED|passed|failed|Error|assert" | cut -c1-200 | head -12
