timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python profiles/wide_ab.py 2>&1 | tail -3
python bench.py --config 3 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-200
