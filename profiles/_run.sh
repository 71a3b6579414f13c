timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | grep -E "FAILED|passed|failed|Error|assert" | cut -c1-200 | head -12
PAIRED="1" DEBUGS="0 64" bash profiles/dbg_modes.sh
