"""7680x4320 single frame, row bands across the GPUs of one box + NCCL all-gather (BASELINE config 5b).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/band_8k.py [freq W H]

Prints one JSON line on rank 0: frames/s of the banded frame (device-timed, max over ranks), the share of the
all-gather, and a parity check of the assembled frame against the single-GPU whole-frame render of rank 0."""
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
br = multigpu.BandRenderer(ctx, W, H, rank, world)
pitches = rs.turntable_pitches(0.0, 64)
rots = [rs.rotation_from_euler(0.0, p, 0.0) for p in pitches]
for k in range(4):
    g = br.render(rots[k])
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
K = 24
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(br.stream)
torch.cuda.current_stream().wait_stream(br.stream)
for k in range(K):
    br.stream.wait_stream(torch.cuda.current_stream())   # next frame's band after the previous gather
    g = br.render(rots[k % 64])
e1.record(torch.cuda.current_stream())
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
frame = br.to_frame(br.render(rots[5]))
ok = None
if rank == 0:
    whole = rs.Context.blank(True, device=local)
    whole.set_scene(xyz, rgb, s0)
    whole.resize(W, H)
    ref, _ = whole.render(rots[5])
    ok = bool(np.array_equal(ref, frame))
    whole.close()
    print(json.dumps({"workload": f"{label} ({len(xyz)} triangles) at {W}x{H}, one frame in {world} row bands",
                      "n_gpus": world, "frames_per_s": K / (float(ms[0]) * 1e-3), "ms_per_frame": float(ms[0]) / K,
                      "gather_bytes_per_gpu": 4 * W * H // world, "banded_equals_whole_frame": ok}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
ctx.close()
