for args in "cylinder 256 9 9" "cylinder 256 3 2" "cylinder 192 14 19" "cone 160 9 9"; do timeout 60 python profiles/_dbg.py $args 2>&1 | grep -v "^$" | tail -1 | cut -c1-120; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python profiles/mip_run.py | tail -1
