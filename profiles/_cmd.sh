timeout 180 python __graft_entry__.py --smoke 2>&1 | tail -5; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -3
