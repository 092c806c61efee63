"""7680x4320 single frame, row bands across the GPUs of one box (BASELINE config 5b): bands written straight into
the root's frame over NVLink ("peer", default) or exchanged with one NCCL all-gather ("allgather").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/band_8k.py [freq W H [mode]]

Prints one JSON line on rank 0: frames/s of the banded frame (device-timed, max over ranks) and a parity check of the assembled frame against the single-GPU whole-frame render of rank 0."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes, multigpu

scene = sys.argv[1] if len(sys.argv) > 1 else "708"     # icosphere frequency, or the name of a bundled soup
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (7680, 4320)
mode = sys.argv[4] if len(sys.argv) > 4 else None
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if scene.isdigit():
    freq = int(scene)
    xyz, rgb, s0 = meshes.icosphere(freq)
    label = f"icosphere f={freq}"
else:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenes as S
    xyz, rgb, s0 = S.soup(scene)
    label = scene
ctx = rs.Context.blank(True, device=local)
ctx.set_scene(xyz, rgb, s0)
DEPTH = 8
br = multigpu.BandRenderer(ctx, W, H, rank, world, mode=mode, depth=DEPTH)
pitches = rs.turntable_pitches(0.0, 64)
rots = [rs.rotation_from_euler(0.0, p, 0.0) for p in pitches]
for k in range(4):
    g = br.render(rots[k])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
K = 24
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(torch.cuda.current_stream())
for k in range(K):
    g = br.render(rots[k % 64])      # each frame waits for the previous frame's gather / completion signal
e1.record(torch.cuda.current_stream())
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
batch_ms = None
if br.mode == "peer":   # the same frames, DEPTH per call: geometry k+1 overlaps resolve + NVLink stores of frame k
    rb = np.stack(rots[:DEPTH])
    for _ in range(2):
        br.render_batch(rb)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record(torch.cuda.current_stream())
    for _ in range(K // DEPTH):
        ptrs = br.render_batch(rb)
    e1.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    bms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(bms, op=dist.ReduceOp.MAX)
    batch_ms = float(bms[0]) / (K // DEPTH * DEPTH)
    batch_frame5 = br.to_frame(ptrs[5]) if rank == 0 else None
res = br.render(rots[5])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ok = None
if rank == 0:
    frame = br.to_frame(res)
    whole = rs.Context.blank(True, device=local)
    whole.set_scene(xyz, rgb, s0)
    whole.resize(W, H)
    ref, _ = whole.render(rots[5])
    ok = bool(np.array_equal(ref, frame))
    if batch_ms is not None:
        ok = ok and bool(np.array_equal(ref, batch_frame5))
    whole.close()
    print(json.dumps({"workload": f"{label} ({len(xyz)} triangles) at {W}x{H}, one frame in {world} row bands",
                      "n_gpus": world, "frames_per_s": K / (float(ms[0]) * 1e-3), "ms_per_frame": float(ms[0]) / K,
                      "mode": br.mode, "batched_ms_per_frame": batch_ms, "batched_frames_per_s": (1e3 / batch_ms if batch_ms else None), "band_bytes_per_gpu": 4 * W * H // world, "banded_equals_whole_frame": ok}))
br.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
ctx.close()
