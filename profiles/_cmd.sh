timeout 900 python bench.py --no-cpu-baseline --no-torch-eager-gpu --steps 5 > gpurun_out/r02d_bench_quick.json 2> gpurun_out/err.txt; tail -c 300 gpurun_out/err.txt
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r02d_bench_quick.json') if l.startswith('{')][-1]
print(d['value'], d['ms_per_step'], d['train'])"
