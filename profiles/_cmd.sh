timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -2
for lib in libnerf_b200.so libnerf_b200_prev.so libnerf_b200.so libnerf_b200_prev.so; do NF_LIB=$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-torch-eager-gpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', d['train']['ms_per_step'])"; done
