python profiles/sdf_run.py
python profiles/sdf_run.py bisect
python profiles/configs_bench.py 2>/dev/null | grep -E "f-4" | cut -c1-200
