timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -E "nf_tc|FAILED|passed|failed|Error|assert" | cut -c1-200 | head -12
CLUSTERS="1" DEBUGS="0 1 2 3" bash profiles/dbg_modes.sh
