timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
