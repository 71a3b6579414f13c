timeout 900 python -m pytest tests -m gpu -x -q -k "sdf" 2>&1 | tail -15
