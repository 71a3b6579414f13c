timeout 900 python -m pytest tests -m gpu -x -q -k "ragged_grids" 2>&1 | tail -8
