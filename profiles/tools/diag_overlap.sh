# Runs the default bench line a few times on one box (value only) -- used to check the box-to-box variance
# of the overlapped batch.
nvidia-smi --query-gpu=serial,power.limit,clocks.max.sm,clocks.sm,temperature.gpu --format=csv,noheader
run() { python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['e2e']['value'])"; }
for i in 1 2 3; do run plain; done
