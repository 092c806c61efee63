"""BASELINE config 4: `sloth models/Pikachu.obj image -w 1920 -h 1080 -j 360` -- the 360-frame JS export,
frames sharded over the GPUs of the box (rank r renders frames r, r+N, ...; no collective).

    python profiles/turntable_export.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/turntable_export.py

Per rank: render + device-side Context::flush (<span> text, ~31 B/cell) + device->host copy, pipelined
(sloth_render_text_batch).  Rank 0 prints one JSON line; `cpu_reference_s_per_frame` is the CPU oracle
(render, one thread) plus a C-speed estimate of the per-cell formatting measured with numpy on one frame."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rust_sloth_b200 as rs
from rust_sloth_b200 import turntable as tt
import scenes as S

W, H, N_FRAMES = 1920, 1080, 360
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
xyz, rgb, s0 = S.soup("pikachu")
ctx = rs.Context.blank(True, device=local)
ctx.set_scene(xyz, rgb, s0)
ctx.resize(W, H)
rots = tt.turntable_rotations(0.0, 0.0, 0.0, N_FRAMES)
mine = tt.frame_shard(len(rots), rank, world)
L = rs.load_library()
import ctypes as C
cap = int(L.sloth_text_capacity(ctx._h, 2)); stride = (cap + 63) & ~63
chunk = 8
pin = rs.PinnedBuffer((chunk * stride + 3) // 4)
lens = (C.c_size_t * chunk)()
def run(frames):
    total = 0
    for i in range(0, len(frames), chunk):
        r = np.ascontiguousarray(rots[frames[i:i + chunk]])
        rs._check(L.sloth_render_text_batch(ctx._h, rs._fp(r), len(r), 2, C.c_void_p(pin.array.ctypes.data), stride, lens))
        total += sum(lens[k] for k in range(len(r)))
    return total
run(mine[:chunk])
if dist: dist.barrier()
t0 = time.perf_counter(); nbytes = run(mine); dt = time.perf_counter() - t0
if dist:
    import torch
    t = torch.tensor([dt, float(nbytes)], dtype=torch.float64, device=f"cuda:{local}")
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dt, nbytes = float(mx[0]), float(t[1])
if rank == 0:
    import oracle
    t1 = time.perf_counter(); cells, _, _ = oracle.render(xyz, rgb, s0, W, H, rots[0], mode=0); t_render = time.perf_counter() - t1
    t1 = time.perf_counter(); ref_txt = rs.flush_bytes(cells[:200000], True, True, True); t_fmt = (time.perf_counter() - t1) * len(cells) / 200000
    got = ctx.render_text_batch(rots[:1], 2)[0]
    print(json.dumps({"workload": f"Pikachu 360-frame -j export at {W}x{H}", "n_gpus": world, "frames": len(rots),
                      "seconds": dt, "frames_per_s": len(rots) / dt, "text_GB": nbytes / 1e9, "text_GB_per_s": nbytes / 1e9 / dt,
                      "first_frame_matches_host_formatter": got == rs.flush_bytes(ctx.render(rots[0])[0], True, True, True),
                      "cpu_oracle_render_s_per_frame": t_render, "python_format_s_per_frame": t_fmt}))
pin.free(); ctx.close()
if dist:
    dist.barrier(); dist.destroy_process_group()
