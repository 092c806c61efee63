"""Bare device->host ceiling of the box: n_active of the N ranks copy a frame-sized buffer (33 MB, the 4K cell buffer)
from their GPU into page-locked host memory at the same time, nothing else running.  This is the roofline of every
end-to-end number in bench.py (`e2e.d2h_ceiling_frames_per_s`).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/d2h_sweep.py
Rank 0 prints one JSON line per n_active in (1, 2, 4, 8) <= N."""
import json, os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NBYTES = 4 * (3840 * 2160 + 2160)
src = torch.empty(NBYTES, dtype=torch.uint8, device=f"cuda:{local}")
dst = [torch.empty(NBYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
REPS = 48
for n_active in (1, 2, 4, 8):
    if n_active > world:
        break
    for d in dst:
        d.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = 0.0
    if rank < n_active:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(REPS):
            dst[i & 1].copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        secs = float(t[0]) * 1e-3
        print(json.dumps({"n_active": n_active, "n_ranks": world, "bytes_per_copy": NBYTES, "copies_per_rank": REPS,
                          "aggregate_GB_per_s": n_active * REPS * NBYTES / secs / 1e9,
                          "per_gpu_GB_per_s": REPS * NBYTES / secs / 1e9,
                          "frames_per_s_ceiling_4k": n_active * REPS / secs}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
