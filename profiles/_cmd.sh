timeout 900 python -m pytest tests -m gpu -x -q -k "normals or sdf" 2>&1 | tail -12
