"""BASELINE.json's own configurations at their full size, CUDA path through the C ABI vs the oracle's FAITHFUL scan
(mode 0: the reference's whole candidate domain, rasterizer.rs:69-70), bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes, turntable
import scenes as S
import textfmt

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frame(ctx, rot, want_z=True):
    cells, z = ctx.render(rot, want_z=want_z)
    return cells, z


def _assert_frame(cells, z, ocells, oz, what):
    bad = np.flatnonzero(cells != ocells)
    assert bad.size == 0, f"{what}: {bad.size} cells differ, first at {bad[:8]}"
    if z is not None:
        badz = np.flatnonzero(z.view(np.uint32) != np.where(oz == 0, np.float32(0), oz).view(np.uint32))
        assert badz.size == 0, f"{what}: {badz.size} z values differ"


def test_config5_ten_million_triangles_4k_vs_faithful_scan():
    """Config 5 / the bench workload: icosphere f=708 (10,025,280 triangles) at 3840x2160 against the oracle's
    mode 0 (about 20 s of CPU), on the path AUTO picks (indexed: k_xform + k_tri) and on the soup path."""
    xyz, rgb, s0 = meshes.icosphere(708)
    assert len(xyz) == 10_025_280
    rot = oracle.rotation(0.0, oracle.turntable(0.0, 64)[5], 0.0)
    ocells, oz, ocnt = oracle.render(xyz, rgb, s0, 3840, 2160, rot, mode=0)
    for path in (rs.PATH_AUTO, rs.PATH_SOUP):
        ctx = rs.Context.blank(True, path=path)
        try:
            ctx.set_scene(xyz, rgb, s0)
            ctx.resize(3840, 2160)
            ctx.stats_enable(count_fragments=True)
            cells, z = _frame(ctx, rot)
            st = ctx.stats()
            assert st["geom_path"] == (rs.PATH_INDEXED if path == rs.PATH_AUTO else rs.PATH_SOUP)
            _assert_frame(cells, z, ocells, oz, f"icosphere f=708 4K path={path}")
            assert st["fragments"] == ocnt["covered"]
            # the batch entry point the bench times leaves the same frame
            batch = ctx.render_batch(np.stack([rot, rot, rot]))
            assert all(np.array_equal(b, ocells) for b in batch)
        finally:
            ctx.close()


def test_config4_all_360_pikachu_turntable_frames_1080p():
    """Config 4: `sloth models/Pikachu.obj image -w 1920 -h 1080 -j 360` -- every one of the 360 frames (71 of them
    wrap fragments past the end of a row) through sloth_render_batch against the oracle's mode 0; every 30th frame
    also as `-j` text through sloth_render_text_batch against the test-side formatter, all text lengths, and the
    stream framing of main.rs:55-57,85-87,98-104."""
    xyz, rgb, s0 = S.soup("pikachu")
    W, H, N = 1920, 1080, 360
    pitches = oracle.turntable(0.0, N)
    assert len(pitches) == N and np.array_equal(pitches, rs.turntable_pitches(0.0, N))
    rots = np.stack([oracle.rotation(0.0, p, 0.0) for p in pitches])
    ctx = rs.Context.blank(True)
    try:
        ctx.set_scene(xyz, rgb, s0)
        ctx.resize(W, H)
        wraps = 0
        text_frames = {}
        for k0 in range(0, N, 40):
            got = ctx.render_batch(rots[k0:k0 + 40])
            for j, cells in enumerate(got):
                k = k0 + j
                ocells, _, ocnt = oracle.render(xyz, rgb, s0, W, H, rots[k], mode=0)
                assert np.array_equal(cells, ocells), f"frame {k}: {(cells != ocells).sum()} cells differ"
                # a wrapped fragment shows in columns 0/1 of a row (x >= 1 never reaches them directly, rasterizer.rs:80)
                body = ocells[:W * H].reshape(H, W)
                wraps += int(((body[:, 0] & 0xFF) != ord(" ")).any())
                if k % 30 == 0:
                    text_frames[k] = ocells
        assert wraps > 0, "no turntable frame wrapped: the test lost its point"
        # text: all lengths, every 30th frame byte for byte
        ks = sorted(text_frames)
        texts = ctx.render_text_batch(rots[ks], 2)
        for k, t in zip(ks, texts):
            assert len(t) == textfmt.webify_length(text_frames[k])
            assert t == textfmt.webify_cells(text_frames[k]), f"frame {k}: -j text differs"
        stream = turntable.webify_stream([text_frames[k] for k in ks[:2]])
        assert stream == b"let frames = [\n`\n" + texts[0] + b"`,\n`\n" + texts[1] + b"`];\n"
    finally:
        ctx.close()
    # the soup path on a strided subset (incl. the wrap frame 144)
    ctx = rs.Context.blank(True, path=rs.PATH_SOUP)
    try:
        ctx.set_scene(xyz, rgb, s0)
        ctx.resize(W, H)
        ks = list(range(0, N, 9))
        got = ctx.render_batch(rots[ks])
        for k, cells in zip(ks, got):
            ocells, _, _ = oracle.render(xyz, rgb, s0, W, H, rots[k], mode=0)
            assert np.array_equal(cells, ocells), f"soup path, frame {k}"
    finally:
        ctx.close()


def test_config2_two_model_scene_1080p():
    """Config 2 (`skull.obj` + the missing `discobole.obj`, substituted by `hand.obj`, 72,958 triangles): a two-model
    mesh queue at 1920x1080 in draw order skull, hand."""
    sx, sr, s0a = S.soup("skull")
    hx, hr, s0b = S.soup("hand")
    xyz, rgb, s0 = np.concatenate([sx, hx]), np.concatenate([sr, hr]), np.float32(max(s0a, s0b))
    for pitch in (S.PI, S.PI + 0.9):
        rot = oracle.rotation(0.0, pitch, 0.0)
        ocells, oz, ocnt = oracle.render(xyz, rgb, s0, 1920, 1080, rot, mode=0)
        for path in (rs.PATH_AUTO, rs.PATH_SOUP):
            ctx = rs.Context.blank(True, path=path)
            try:
                ctx.set_scene(xyz, rgb, s0)
                ctx.resize(1920, 1080)
                ctx.stats_enable(count_fragments=True)
                cells, z = _frame(ctx, rot)
                _assert_frame(cells, z, ocells, oz, f"skull+hand 1080p path={path}")
                assert ctx.stats()["fragments"] == ocnt["covered"]
            finally:
                ctx.close()


def test_parity_grid_bundled_models():
    """SURVEY 8(d)'s parity grid: every bundled model at 80x40, 100x100 and 1920x1080, angle pi and 16 turntable
    angles (mode 0 at the two small sizes; at 1080p pi in mode 0 and the 16 angles in the row-terminating mode 1,
    which test_oracle.py ties to mode 0)."""
    names = ["cube", "ferris", "suzy", "pikachu", "skull", "vaporeon", "cube_stl", "part_stl", "hand"]
    avail = [n for n in names if _has_scene(n)]
    assert len(avail) >= 6, avail
    angles = [S.PI] + [float(p) for p in oracle.turntable(0.0, 16)]
    for name in avail:
        xyz, rgb, s0 = S.soup(name)
        ctx = rs.Context.blank(True)
        try:
            ctx.set_scene(xyz, rgb, s0)
            for W, H in ((80, 40), (100, 100), (1920, 1080)):
                ctx.resize(W, H)
                for i, a in enumerate(angles):
                    rot = oracle.rotation(0.0, a, 0.0)
                    mode = 0 if (W < 1000 or i == 0) else 1
                    if name == "hand" and W > 1000 and i > 4:
                        continue
                    ocells, oz, _ = oracle.render(xyz, rgb, s0, W, H, rot, mode=mode)
                    cells, z = _frame(ctx, rot)
                    _assert_frame(cells, z, ocells, oz, f"{name} {W}x{H} angle {a}")
        finally:
            ctx.close()


def _has_scene(name):
    try:
        S.soup(name)
        return True
    except KeyError:
        return False


def test_row_bands_across_two_ranks_match_the_oracle():
    """Multi-rank band mode on real GPUs: 2 processes (one per GPU), peer stores and all-gather, even and odd
    widths, a wrap frame -- the assembled frame against the oracle.  Skipped on single-GPU boxes."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "band_worker.py")],
                         capture_output=True, text=True, env=env, timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith("BAND_RESULTS ")]
    assert out.returncode == 0 and line, out.stdout[-2000:] + out.stderr[-2000:]
    results = json.loads(line[0][len("BAND_RESULTS "):])
    assert len(results) >= 6 and {r["mode"] for r in results} == {"peer", "allgather"}
    assert all(r["equal"] for r in results), results
