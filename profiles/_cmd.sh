timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-eager-gpu 2>&1 | tail -1 | cut -c1-400
python bench.py --config 5 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-300
python bench.py --config 4 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-300
