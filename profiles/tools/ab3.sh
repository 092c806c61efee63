# Same-box A/B of library builds on the bench workload, one pass: bash profiles/tools/ab3.sh libA.so libB.so ...
run() { SLOTH_B200_LIB=$PWD/$1 python bench.py --steps 48 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms']; print('$1', round(d['value']), 'us/frame', round(d['ms_per_step']*1e3,1), {a: round(b*1e3,1) for a,b in k.items()}, 'ok' if d['post_check']['last_timed_frame_equals_single_render'] else 'MISMATCH')"; }
nvidia-smi --query-gpu=serial --format=csv,noheader
for lib in "$@"; do run $lib; done
