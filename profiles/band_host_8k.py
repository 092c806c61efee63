"""7680x4320 single frame END TO END (device work + the frame in host memory), row bands across the GPUs of one box
with no root GPU: every rank copies its band over its own PCIe link into one page-locked shared-memory frame
(multigpu.HostBandRenderer), band edges re-balanced once from the chunks each band processed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/band_host_8k.py [freq W H]
Rank 0 prints one JSON line: wall-clock ms per frame (max over ranks, barrier on both sides), before and after
re-balancing, and whether the assembled frame equals rank 0's own whole-frame render."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes, multigpu

freq = int(sys.argv[1]) if len(sys.argv) > 1 else 708
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (7680, 4320)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)
xyz, rgb, s0 = meshes.icosphere(freq)
ctx = rs.Context.blank(True, device=local)
ctx.set_scene(xyz, rgb, s0)
hb = multigpu.HostBandRenderer(ctx, W, H, rank, world, name=f"sloth_band_{os.environ.get('MASTER_PORT', '0')}", barrier=barrier)
rots = [rs.rotation_from_euler(0.0, p, 0.0) for p in rs.turntable_pitches(0.0, 64)]

def timed(n):
    for k in range(3):
        hb.render(rots[k])
    barrier()
    t0 = time.perf_counter()
    for k in range(n):
        hb.render(rots[k % 64])
    hb.wait(hb.k)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / n * 1e3

K = 24
ms_equal = timed(K)
edges_equal = list(hb.edges)
ms_bal, edges_bal = None, None
if world > 1:
    ctx.stats_enable(count_fragments=True)
    hb.render(rots[0])
    mine = torch.tensor([float(ctx.stats()["chunks_processed"])], dtype=torch.float64, device=f"cuda:{local}")
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    ctx.stats_enable()
    edges_bal = hb.rebalance([float(c[0]) for c in allc])
    ms_bal = timed(K)
# parity of the assembled frame (rank 0 renders the whole frame on its own GPU)
k_last = hb.render(rots[7])
frame = hb.wait(k_last).copy()
barrier()
ok = None
if rank == 0:
    ref = rs.Context.blank(True, device=local)
    ref.set_scene(xyz, rgb, s0)
    ref.resize(W, H)
    whole, _ = ref.render(rots[7])
    ref.close()
    ok = bool(np.array_equal(whole, frame))
    print(json.dumps({"workload": f"icosphere f={freq} ({len(xyz)} triangles) at {W}x{H}, one frame in {world} row bands, "
                                  "each band copied to the shared host frame by its own GPU",
                      "n_gpus": world, "e2e_ms_per_frame_equal_rows": ms_equal, "edges_equal_rows": edges_equal,
                      "e2e_ms_per_frame_balanced": ms_bal, "edges_balanced": edges_bal,
                      "frame_bytes": 4 * (W * H + H), "assembled_equals_whole_frame_render": ok}), flush=True)
barrier()
hb.close()
ctx.close()
if world > 1:
    dist.destroy_process_group()
