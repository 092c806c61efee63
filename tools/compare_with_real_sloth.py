#!/usr/bin/env python3
"""Pins the oracle (and the host loaders / flush formatter it is fed by) against a REAL `sloth` binary.

Not runnable in the build image (no rustc / cargo, crates not vendored): it is shipped for whoever has a Rust
toolchain.  No GPU is needed -- only the CPU oracle (oracle/), the host loaders (libsloth_host.so) and the stock,
unmodified reference binary.

    cargo build --release --manifest-path /path/to/rust-sloth/Cargo.toml
    python tools/compare_with_real_sloth.py /path/to/rust-sloth/target/release/sloth /path/to/rust-sloth/models

For every case of tests/golden/oracle_frames.json (scene, W, H, roll, pitch, yaw) the binary is run three ways and its
stdout compared byte for byte with what this repo derives from the oracle's frame for the rotation the CLI really uses
(match_turntable adds PI to -y in f32, inputs.rs:131-149):
  1. `-b image -w W -h H`          glyphs only: pins rasterizer.rs:48-93 / geometry.rs:37-56 / context.rs:93-141 and the
                                   loaders' geometry (tobj 3.2.2, stl_io 0.4.2)
  2. `image -w W -h H`             crossterm 0.18 truecolor stream: pins the colours (materials, vertex colours) and the
                                   byte order of flush (context.rs:63-80)
  3. `image -w W -h H -j 3`        the JS-frame export (main.rs:55-58,85-106)
A summary line per case; exit code 1 if anything differs.  The cell hashes are also checked against the committed
golden file, so a green run means: golden hashes == oracle == real binary.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import rust_sloth_b200 as rs  # noqa: E402  (host loaders and formatters only; no CUDA call is made)
from rust_sloth_b200 import turntable as tt  # noqa: E402

SCENES = {
    "cube": "cube.obj", "ferris": "ferris.obj", "suzy": "suzy.obj", "pikachu": "Pikachu.obj", "skull": "skull.obj",
    "vaporeon": "Vaporeon.obj", "cube_stl": "cube.stl", "part_stl": "part.stl", "suzy_suzy": "suzy.obj suzy.obj",
    "hand": "hand.obj",
}
PI32 = np.float32(np.pi)


def main():
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    exe, models = sys.argv[1], sys.argv[2]
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_frames.json")))
    bad = 0
    for case in gold["cases"]:
        files = [os.path.join(models, f) for f in SCENES[case["scene"]].split(" ")]
        meshes = rs.match_meshes(" ".join(files))
        xyz = np.concatenate([m.xyz for m in meshes])
        rgb = np.concatenate([m.rgb for m in meshes])
        s0 = np.float32(rs.scene_scale0(meshes))
        W, H = case["W"], case["H"]
        # the -y value whose f32 sum with PI is the golden pitch (exact for every committed case: checked below)
        y_arg = np.float32(np.float32(case["pitch"]) - PI32)
        pitch = np.float32(y_arg + PI32)
        rot = oracle.rotation(case["roll"], pitch, case["yaw"])
        cells, _, _ = oracle.render(xyz, rgb, s0, W, H, rot, image=True, mode=0)
        same_as_golden = float(pitch) == case["pitch"] and hashlib.sha256(cells.tobytes()).hexdigest() == case["cells_sha256"]
        values = [float(np.float32(case["roll"])), float(y_arg), float(np.float32(case["yaw"]))]
        if any(v < 0 for v in values):   # clap 2 takes a value that starts with '-' for a flag (inputs.rs sets no allow_hyphen_values)
            print(f"{case['scene']} {W}x{H}: skipped (negative rotation value)")
            continue
        rot_args = ["-x", repr(values[0]), "-y", repr(values[1]), "-z", repr(values[2])]
        results = []
        run = lambda extra: subprocess.run([exe] + files + extra, capture_output=True, check=True).stdout  # noqa: E731
        results.append(("plain", run(["-b", "image", "-w", str(W), "-h", str(H)] + rot_args) == rs.flush_bytes(cells, False, False, True)))
        results.append(("colour", run(["image", "-w", str(W), "-h", str(H)] + rot_args) == rs.flush_bytes(cells, True, False, True)))
        if W * H <= 200 * 100:
            frames = [oracle.render(xyz, rgb, s0, W, H, oracle.rotation(case["roll"], p, case["yaw"]), mode=0)[0]
                      for p in oracle.turntable(float(y_arg), 3)]
            results.append(("webify", run(["image", "-w", str(W), "-h", str(H), "-j", "3"] + rot_args) == tt.webify_stream(frames)))
        ok = all(r for _, r in results)
        bad += 0 if ok else 1
        print(f"{case['scene']} {W}x{H}: " + ", ".join(f"{n} {'ok' if r else 'DIFFERS'}" for n, r in results) +
              ("" if same_as_golden else "  (rotation differs from the golden case by one f32 rounding: compared at the CLI's rotation)"))
    print("oracle pinned against the real binary" if bad == 0 else f"{bad} case(s) differ")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
