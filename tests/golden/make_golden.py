"""Regenerates tests/golden/*.  Run in the build container (needs /root/reference/models):

    python tests/golden/make_golden.py

1. models.npz / hand.npz: the bundled models of the reference as triangle soups, produced by
   the product's host loader (rust-sloth_b200/host/mesh_io.cpp, the tobj/stl_io rules of
   SURVEY.md Appendix C) -- the GPU box has no /root/reference, so the parity tests read these.
2. oracle_frames.json: SHA-256 of the oracle's cell buffer and z-buffer plus its counters for a
   list of (scene, W, H, roll, pitch, yaw) cases, and the full text of the Pikachu 80x40 frame
   (BASELINE config 1).  These pin the ORACLE (a change in oracle/sloth_oracle.c that alters any
   frame shows up here); they are not outputs of the Rust reference, which cannot be built here.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import rust_sloth_b200 as rs  # noqa: E402

MODELS = "/root/reference/models/"
HERE = os.path.dirname(os.path.abspath(__file__))

SCENES = {
    "cube": "cube.obj", "ferris": "ferris.obj", "suzy": "suzy.obj", "pikachu": "Pikachu.obj",
    "skull": "skull.obj", "vaporeon": "Vaporeon.obj", "cube_stl": "cube.stl", "part_stl": "part.stl",
    "suzy_suzy": "suzy.obj suzy.obj", "hand": "hand.obj",
}

PI = float(np.float32(np.pi))
CASES = [  # scene, W, H, roll, pitch, yaw
    ("pikachu", 80, 40, 0.0, PI, 0.0),
    ("pikachu", 100, 100, 0.0, PI, 0.0),
    ("pikachu", 160, 80, 0.0, None, 0.0),      # None -> frame 144 of the 360-frame turntable (row wrap)
    ("pikachu", 1920, 1080, 0.0, PI, 0.0),
    ("skull", 100, 100, 0.0, PI, 0.0),
    ("skull", 1920, 1080, 0.0, PI, 0.0),
    ("suzy", 80, 40, 0.3, PI, 0.2),
    ("suzy_suzy", 101, 57, 0.0, PI, 0.0),
    ("suzy_suzy", 640, 360, 0.0, PI, 0.0),
    ("cube", 80, 40, 0.5, 4.0, 0.25),
    ("ferris", 81, 41, 0.0, PI, 0.0),
    ("vaporeon", 200, 100, 0.0, PI, 0.0),
    ("cube_stl", 80, 40, 0.4, 3.5, 0.1),
    ("part_stl", 120, 60, 0.0, PI, 0.0),
    ("hand", 1920, 1080, 0.0, PI, 0.0),
]


def load(arg):
    meshes = rs.match_meshes(" ".join(MODELS + a for a in arg.split(" ")))
    xyz = np.concatenate([m.xyz for m in meshes])
    rgb = np.concatenate([m.rgb for m in meshes])
    sizes = np.array([len(m) for m in meshes], np.int64)
    return xyz, rgb, np.float32(rs.scene_scale0(meshes)), sizes


def main():
    soups, hand = {}, {}
    data = {}
    for name, arg in SCENES.items():
        xyz, rgb, s0, sizes = load(arg)
        data[name] = (xyz, rgb, s0)
        tgt = hand if name == "hand" else soups
        if name != "suzy_suzy":  # rebuilt from suzy at load time
            tgt[name + "_xyz"], tgt[name + "_rgb"] = xyz, rgb
            tgt[name + "_scale0"], tgt[name + "_sizes"] = s0, sizes
    np.savez_compressed(os.path.join(HERE, "models.npz"), **soups)
    np.savez_compressed(os.path.join(HERE, "hand.npz"), **hand)

    out = {"cases": []}
    pitches = oracle.turntable(0.0, 360)
    for scene, W, H, roll, pitch, yaw in CASES:
        if pitch is None:
            pitch = float(pitches[144])
        xyz, rgb, s0 = data[scene]
        rot = oracle.rotation(roll, pitch, yaw)
        cells, z, cnt = oracle.render(xyz, rgb, s0, W, H, rot, image=True, mode=0)
        rec = {"scene": scene, "W": W, "H": H, "roll": roll, "pitch": pitch, "yaw": yaw,
               "cells_sha256": hashlib.sha256(cells.tobytes()).hexdigest(),
               "z_sha256": hashlib.sha256(z.tobytes()).hexdigest(), "counters": cnt}
        if (scene, W, H) == ("pikachu", 80, 40):
            rec["text"] = oracle.cells_to_text(cells)
        out["cases"].append(rec)
        print(scene, W, H, cnt)
    out["turntable_360_first8"] = [float(p) for p in pitches[:8]]
    out["turntable_360_count"] = int(len(pitches))
    with open(os.path.join(HERE, "oracle_frames.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
