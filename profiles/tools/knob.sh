# bash profiles/tools/knob.sh VAR v1 v2 ...   -- the default bench line for several values of one SLOTH_* knob
var=$1; shift
nvidia-smi --query-gpu=serial --format=csv,noheader
for v in "$@"; do env $var=$v python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$var=$v', round(d['value']), 'us/frame', round(d['ms_per_step']*1e3,1), 'k_geom3', round(d['roofline']['avg_launch_ms']*1e3,1))"; done
