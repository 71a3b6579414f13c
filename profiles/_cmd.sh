timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E|passed|failed|Error" | head -20
