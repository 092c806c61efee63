# bash profiles/tools/knob2.sh VAR v1 v2 ...   -- kernel times of the bench workload for several values of one SLOTH_* knob
var=$1; shift
nvidia-smi --query-gpu=serial --format=csv,noheader
for v in "$@"; do env $var=$v python bench.py --steps 24 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms']; print('$var=$v', round(d['value']), 'us/frame', round(d['ms_per_step']*1e3,1), {a: round(b*1e3,1) for a,b in k.items()})"; done
