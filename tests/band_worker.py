"""Worker of tests/test_gpu_configs.py::test_row_bands_across_two_ranks_match_the_oracle: one process per GPU
(torch.distributed.run), a banded frame assembled in `peer` and `allgather` mode, even and odd widths, compared on
rank 0 with the ORACLE (not with the product's own single-GPU frame)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy as np
import torch
import torch.distributed as dist

import oracle
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes, multigpu
import scenes as S

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
results = []
cases = [("icosphere", meshes.icosphere(60), 640, 360), ("icosphere", meshes.icosphere(40), 333, 201),
         ("pikachu", S.soup("pikachu"), 320, 160), ("pikachu-wrap", S.soup("pikachu"), 161, 80)]
for name, (xyz, rgb, s0), W, H in cases:
    pitch = oracle.turntable(0.0, 360)[144] if "wrap" in name else np.float32(np.pi) + 0.3
    rot = oracle.rotation(0.1, pitch, 0.05)
    for mode in ("peer", "allgather"):
        edges_ok = all((e * W) % 4 == 0 for e in multigpu.band_edges(H, world))
        if mode == "peer" and not edges_ok:
            continue
        ctx = rs.Context.blank(True, device=local)
        ctx.set_scene(xyz, rgb, s0)
        br = multigpu.BandRenderer(ctx, W, H, rank, world, mode=mode)
        for _ in range(2):      # twice: the second frame reuses the key planes / frame slots
            out = br.render(rot)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            cells = br.to_frame(out)
            ocells, _, _ = oracle.render(xyz, rgb, s0, W, H, rot, image=True, mode=0)
            results.append({"case": name, "W": W, "H": H, "mode": mode, "equal": bool(np.array_equal(cells, ocells)),
                            "differing": int((cells != ocells).sum())})
        dist.barrier()
        br.close()
        ctx.close()
if rank == 0:
    print("BAND_RESULTS " + json.dumps(results), flush=True)
dist.barrier()
dist.destroy_process_group()
