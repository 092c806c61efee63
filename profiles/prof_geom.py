"""Short driver for ncu: a few frames of the benchmark workload (icosphere f=708 at 3840x2160)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes

freq = int(sys.argv[1]) if len(sys.argv) > 1 else 708
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160)
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 4
xyz, rgb, s0 = meshes.icosphere(freq)
ctx = rs.Context.blank(True)
ctx.set_scene(xyz, rgb, s0)
ctx.resize(W, H)
pitches = rs.turntable_pitches(0.0, 64)
out = np.empty(ctx.cells_per_frame(), np.uint32)
for k in range(frames):
    cells, _ = ctx.render(rs.rotation_from_euler(0.0, pitches[k], 0.0))
print("done", ctx.stats())
ctx.close()
