timeout 400 python profiles/train_kinds.py 2>/dev/null | tee gpurun_out/r02d_train_kinds.jsonl
