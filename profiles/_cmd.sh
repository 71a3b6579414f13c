timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python profiles/ab_time.py libnerf_b200.so libnerf_b200_nowc.so libnerf_b200.so libnerf_b200_nowc.so
