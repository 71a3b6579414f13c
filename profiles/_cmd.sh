timeout 900 python -m pytest tests -m gpu -x -q -k "positional" 2>&1 | tail -8
python profiles/configs_bench.py 2>/dev/null | tail -12 | cut -c1-230
