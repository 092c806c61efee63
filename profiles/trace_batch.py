"""`SLOTH_DEBUG=512 python profiles/trace_batch.py` -- timeline of the first frames of a device batch on the bench workload
(the library prints begin / end of every stage per frame to stderr)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes

xyz, rgb, s0 = meshes.icosphere(708)
ctx = rs.Context.blank(True)
ctx.set_scene(xyz, rgb, s0)
ctx.resize(3840, 2160)
pitches = rs.turntable_pitches(0.0, 64)
rots = np.stack([rs.rotation_from_euler(0.0, p, 0.0) for p in pitches[:16]])
d = torch.empty(ctx.cells_per_frame() + 2, dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.render_device_batch(rots, d.data_ptr(), 0)
    ctx.sync()
ctx.close()
