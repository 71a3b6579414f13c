# timing experiments: NF_TC_PAIRED (1 = cta_group::2 two-tile pipeline, 0 = single-CTA pipeline) x NF_TC_DEBUG
# (1 = epilogue idle, 2 = no MMA issued, 3 = weight streaming + barriers only)
for c in ${PAIRED:-1 0}; do for d in ${DEBUGS:-0 3}; do
  echo -n "paired=$c debug=$d ms_per_step="; NF_TC_PAIRED=$c NF_TC_DEBUG=$d python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"; done; done
