"""Load-time comparison for SURVEY 8f next-2: the host loader (host/mesh_io.cpp, what the CLI used before) against
sloth_scene_load (file -> pinned memory -> GPU parse).  Usage: python profiles/loader_bench.py [f]  (20 f^2 triangles)."""
import os
import struct
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes


def main():
    f = int(sys.argv[1]) if len(sys.argv) > 1 else 316
    xyz, rgb, _ = meshes.icosphere(f)
    n = xyz.shape[0]
    tmp = tempfile.mkdtemp()
    t0 = time.time()
    verts, inv = np.unique(xyz.reshape(-1, 3), axis=0, return_inverse=True)
    obj = os.path.join(tmp, "ico.obj")
    with open(obj, "w") as fh:
        fh.write("\n".join("v " + " ".join(r) for r in np.char.mod("%.9g", verts)) + "\n")
        fh.write("\n".join("f " + " ".join(r) for r in (inv.reshape(-1, 3) + 1).astype(str)) + "\n")
    stl = os.path.join(tmp, "ico.stl")
    rec = np.zeros(n, dtype=[("n", "<f4", 3), ("v", "<f4", 9), ("a", "<u2")])
    rec["v"] = xyz
    with open(stl, "wb") as fh:
        fh.write(b"\0" * 80 + struct.pack("<I", n) + rec.tobytes())
    print(f"# {n} triangles, {verts.shape[0]} vertices; wrote files in {time.time() - t0:.1f} s", flush=True)
    ctx = rs.Context.blank(True)
    ctx.load_models(stl)   # warm-up: CUDA context, module load
    rows = []
    for path in (obj, stl):
        size = os.path.getsize(path)
        t0 = time.time(); ms = rs.match_meshes(path); t_host = time.time() - t0
        hx = np.concatenate([m.xyz for m in ms]).reshape(-1, 9)
        hr = np.concatenate([m.rgb for m in ms]).reshape(-1, 3)
        t0 = time.time(); ctx.set_scene(hx, hr, rs.scene_scale0(ms)); t_up = time.time() - t0
        best = 1e9
        for _ in range(3):
            t0 = time.time(); ctx.load_models(path); best = min(best, time.time() - t0)
        st = ctx.stats()
        phases = "read %.1f / parse %.1f / commit %.1f ms" % (st["load_read_ms"], st["load_parse_ms"], st["load_commit_ms"])
        dx, dr, _ = ctx.scene()
        same = np.array_equal(dx.view(np.uint32), hx.view(np.uint32)) and np.array_equal(dr, hr)
        rows.append((os.path.basename(path), size / 1e6, t_host + t_up, best, same, phases))
    print("| file | MB | host parse + sloth_scene_set (s) | sloth_scene_load (s) | speed-up | identical | last device load |")
    print("|---|---|---|---|---|---|---|")
    for name, mb, th, td, same, phases in rows:
        print(f"| {name} | {mb:.1f} | {th:.3f} | {td:.3f} | {th / td:.1f}x | {same} | {phases} |")


if __name__ == "__main__":
    main()
