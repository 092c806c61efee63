"""Per-kernel device times for the BASELINE configs other than the headline one.
python profiles/perf_scenes.py  -> one line per scene (uses tests/golden soups; no /root/reference)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes
import scenes as S

CASES = [("pikachu", 80, 40), ("pikachu", 1920, 1080), ("skull", 1920, 1080), ("hand", 1920, 1080),
         ("suzy_suzy", 3840, 2160), ("cube", 3840, 2160), ("vaporeon", 1920, 1080), ("icosphere64", 3840, 2160)]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] in sys.argv[1:]]
pitches = rs.turntable_pitches(0.0, 32)
for name, W, H in CASES:
    xyz, rgb, s0 = meshes.icosphere(64) if name == "icosphere64" else S.soup(name)
    ctx = rs.Context.blank(True)
    ctx.set_scene(xyz, rgb, s0)
    ctx.resize(W, H)
    ctx.stats_enable(count_fragments=True, kernel_timing=True)
    acc = {"geom_ms": [], "walk_ms": [], "resolve_ms": [], "last_frame_ms": []}
    frags = 0
    for k in range(40):
        ctx.render(rs.rotation_from_euler(0.0, pitches[k % 32], 0.0))
        st = ctx.stats()
        if k >= 8:
            for key in acc:
                acc[key].append(st[key])
            frags += st["fragments"]
    rots = np.stack([rs.rotation_from_euler(0.0, p, 0.0) for p in pitches])
    pin = rs.PinnedBuffer(32 * ctx.cells_per_frame())
    ctx.stats_enable()
    ctx.render_batch(rots[:4], pin.array)
    t = time.perf_counter(); ctx.render_batch(rots, pin.array); dt = time.perf_counter() - t
    dev = ctx.stats()["last_frame_ms"]
    print(json.dumps({"scene": name, "W": W, "H": H, "n_tri": len(xyz), "frags_per_frame": frags // 32,
                      **{k: round(float(np.mean(v)) * 1e3, 1) for k, v in acc.items()},
                      "batch_dev_us_per_frame": round(dev * 1e3, 1), "batch_e2e_fps": round(32 / dt, 1),
                      "walk_tris": st["walk_tris"], "walk_items": st["walk_items"]}))
    pin.free(); ctx.close()
