"""`python profiles/cone_stats.py [freq W H]` -- how many chunks k_super_pass takes off k_tri's list on the bench workload
(chunks_processed of sloth_stats with the fragment counter on), per turntable angle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes

freq = int(sys.argv[1]) if len(sys.argv) > 1 else 708
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160)
xyz, rgb, s0 = meshes.icosphere(freq)
ctx = rs.Context.blank(True)
ctx.set_scene(xyz, rgb, s0)
ctx.resize(W, H)
n_chunks = (len(xyz) + 31) // 32
pitches = rs.turntable_pitches(0.0, 64)
ctx.stats_enable(count_fragments=True, kernel_timing=True)
for k in (0, 7, 21, 40):
    ctx.render(rs.rotation_from_euler(0.0, pitches[k], 0.0))
    st = ctx.stats()
    print(f"angle {k}: chunks through k_tri {st['chunks_processed']} of {n_chunks} ({100.0 * st['chunks_processed'] / n_chunks:.1f} %), "
          f"fragments {st['fragments']}, xform+super {st['xform_ms'] * 1e3:.1f} us, k_tri {st['geom_ms'] * 1e3:.1f} us")
ctx.close()
