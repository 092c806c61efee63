# Same-box A/B of library builds: bash profiles/tools/ab.sh libA.so libB.so ...  (paths relative to the repo root)
run() { SLOTH_B200_LIB=$PWD/$1 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'us/frame', round(d['ms_per_step']*1e3,1), 'k_geom3', round(d['roofline']['avg_launch_ms']*1e3,1))"; }
nvidia-smi --query-gpu=serial --format=csv,noheader
for rep in 1 2; do for lib in "$@"; do run $lib; done; done
