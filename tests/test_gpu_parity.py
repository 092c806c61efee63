"""GPU parity: the CUDA path through the C ABI vs the CPU oracle, bit for bit
(cells: glyph + colour of every terminal cell; z-buffer: every depth winner)."""
import hashlib
import os

import numpy as np
import pytest

import oracle
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes
import scenes as S

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["soup", "indexed", "indexed-tiles", "indexed-pairs", "indexed-nocone"])
def geom_path(request, monkeypatch):
    """Runs the test once per geometry path: SLOTH_PATH pins what every context created inside the test uses
    (1 = k_geom3 over the soup, 2 = k_xform + k_tri over the deduplicated vertices), whatever AUTO would pick;
    "-tiles" additionally sends the queued large triangles through the binned tile path on frames of any size
    (by default small frames leave them to the walk kernel)."""
    monkeypatch.setenv("SLOTH_PATH", "1" if request.param == "soup" else "2")
    monkeypatch.setenv("SLOTH_TILES", "2" if request.param.endswith("tiles") else "0")
    monkeypatch.setenv("SLOTH_TRI2", "1" if request.param.endswith("pairs") else "0")   # k_tri2: two chunks per warp turn
    monkeypatch.setenv("SLOTH_CONE", "0" if request.param.endswith("nocone") else "1")  # 0: no super-chunk is skipped
    return request.param


def gpu_frame(xyz, rgb, s0, W, H, rot, image=True, want_z=True, band=None):
    ctx = rs.Context.blank(image)
    try:
        ctx.set_scene(xyz, rgb, s0)
        ctx.resize(W, H)
        if band:
            ctx.set_band(*band)
        ctx.stats_enable(count_fragments=True)
        cells, z = ctx.render(rot, want_z=want_z and not band)
        return cells, z, ctx.stats()
    finally:
        ctx.close()


def assert_same(cells, z, ocells, oz, what):
    bad = np.flatnonzero(cells != ocells)
    assert bad.size == 0, f"{what}: {bad.size} cells differ, first at {bad[:8]}: gpu={cells[bad[:8]]} oracle={ocells[bad[:8]]}"
    if z is not None:
        badz = np.flatnonzero(z.view(np.uint32) != np.where(oz == 0, np.float32(0), oz).view(np.uint32))
        assert badz.size == 0, f"{what}: {badz.size} z values differ, first at {badz[:8]}"


@pytest.mark.parametrize("case", S.golden()["cases"], ids=lambda c: f"{c['scene']}-{c['W']}x{c['H']}")
def test_bundled_models_match_oracle_and_golden(case, geom_path):
    xyz, rgb, s0 = S.soup(case["scene"])
    rot = oracle.rotation(case["roll"], case["pitch"], case["yaw"])
    ocells, oz, ocnt = oracle.render(xyz, rgb, s0, case["W"], case["H"], rot, image=True, mode=0)
    assert hashlib.sha256(ocells.tobytes()).hexdigest() == case["cells_sha256"]
    cells, z, st = gpu_frame(xyz, rgb, s0, case["W"], case["H"], rot)
    assert_same(cells, z, ocells, oz, str(case["scene"]))
    assert hashlib.sha256(cells.tobytes()).hexdigest() == case["cells_sha256"]
    assert st["fragments"] == ocnt["covered"]


@pytest.mark.parametrize("image", [True, False])
@pytest.mark.parametrize("kind", ["uniform", "small", "sliver", "collinear", "dup", "axis"])
def test_fuzz_soups(kind, image, geom_path):
    sizes = [(7, 7), (8, 9), (33, 20), (64, 64), (101, 57), (160, 80), (199, 200), (2, 2), (3, 1), (1, 5)]
    for seed in range(40):
        n = 1 + (seed * 7) % 64
        xyz, rgb, s0 = meshes.random_soup(seed, n, kind=kind)
        W, H = sizes[seed % len(sizes)]
        rot = oracle.rotation(0.1 * seed, np.float32(np.pi) + 0.37 * seed, 0.05 * seed)
        ocells, oz, _ = oracle.render(xyz, rgb, s0, W, H, rot, image=image, mode=0)
        cells, z, _ = gpu_frame(xyz, rgb, s0, W, H, rot, image=image)
        assert_same(cells, z, ocells, oz, f"{kind} seed={seed} n={n} {W}x{H} image={image}")


def test_non_finite_and_huge_coordinates(geom_path):
    xyz, rgb, s0 = meshes.random_soup(3, 24)
    xyz = xyz.copy()
    xyz[1, 2] = np.nan
    xyz[2, 0] = np.inf
    xyz[3, 4] = -np.inf
    xyz[4] *= np.float32(1e30)
    xyz[5, 8] = np.nan
    xyz[6, 5] = np.float32(3e38)
    rot = oracle.rotation(0.2, 3.0, 0.1)
    for (W, H) in [(40, 20), (41, 21)]:
        ocells, oz, _ = oracle.render(xyz, rgb, s0, W, H, rot, mode=0)
        cells, z, st = gpu_frame(xyz, rgb, s0, W, H, rot)
        assert st["irregular_tris"] > 0
        assert_same(cells, z, ocells, oz, f"nonfinite {W}x{H}")


def test_degenerate_scene_scale(geom_path):
    xyz, rgb, _ = meshes.random_soup(5, 16)
    rot = oracle.rotation(0.0, 3.0, 0.0)
    for s0 in [0.0, np.inf, np.nan, 1e-30]:
        ocells, oz, _ = oracle.render(xyz, rgb, s0, 30, 20, rot, mode=0)
        cells, z, _ = gpu_frame(xyz, rgb, s0, 30, 20, rot)
        assert_same(cells, z, ocells, oz, f"scale0={s0}")


def test_empty_scene_and_tiny_frames(geom_path):
    rot = oracle.rotation(0.0, 3.0, 0.0)
    e_xyz, e_rgb = np.zeros((0, 9), np.float32), np.zeros((0, 3), np.uint8)
    for (W, H) in [(1, 1), (2, 3), (80, 40)]:
        ocells, oz, _ = oracle.render(e_xyz, e_rgb, 1.0, W, H, rot)
        cells, z, _ = gpu_frame(e_xyz, e_rgb, 1.0, W, H, rot)
        assert_same(cells, z, ocells, oz, f"empty {W}x{H}")


def test_icosphere_small_frequencies(geom_path):
    for f, (W, H) in [(4, (80, 40)), (24, (160, 80)), (64, (320, 200))]:
        xyz, rgb, s0 = meshes.icosphere(f)
        for k in range(3):
            rot = oracle.rotation(0.0, np.float32(np.pi) + 0.5 * k, 0.0)
            ocells, oz, ocnt = oracle.render(xyz, rgb, s0, W, H, rot, mode=0)
            cells, z, st = gpu_frame(xyz, rgb, s0, W, H, rot)
            assert_same(cells, z, ocells, oz, f"icosphere f={f}")
            assert st["fragments"] == ocnt["covered"]


def test_custom_shader_table(geom_path):
    xyz, rgb, s0 = S.soup("suzy")
    rot = oracle.rotation(0.0, S.PI, 0.0)
    thr = np.array([0.1, 0.15, 0.35, 0.5, 0.55, 0.6, 0.85, 0.95, 2.0], np.float32)
    gl = b"abcdefghi?"
    ocells, oz, _ = oracle.render(xyz, rgb, s0, 120, 60, rot, thr=thr, glyph=gl)
    ctx = rs.Context.blank(True)
    ctx.set_scene(xyz, rgb, s0)
    ctx.resize(120, 60)
    ctx.set_shader(thr, gl)
    cells, z = ctx.render(rot, want_z=True)
    ctx.close()
    assert_same(cells, z, ocells, oz, "custom shader")


def test_row_bands_reassemble_to_the_whole_frame(geom_path):
    xyz, rgb, s0 = S.soup("pikachu")
    pitches = oracle.turntable(0.0, 360)
    for (W, H, nb) in [(160, 80, 4), (161, 83, 3), (80, 40, 8)]:
        rot = oracle.rotation(0.0, pitches[144], 0.0)  # a frame with row wrap
        ocells, _, _ = oracle.render(xyz, rgb, s0, W, H, rot, mode=0)
        parts = []
        edges = [H * i // nb for i in range(nb + 1)]
        for i in range(nb):
            cells, _, _ = gpu_frame(xyz, rgb, s0, W, H, rot, band=(edges[i], edges[i + 1]))
            assert cells.size == (edges[i + 1] - edges[i]) * W
            parts.append(cells)
        whole = np.concatenate(parts + [np.full(H, ord(" "), np.uint32)])
        bad = np.flatnonzero(whole != ocells)
        assert bad.size == 0, f"bands {W}x{H}/{nb}: {bad.size} cells differ, first {bad[:8]}"


def test_row_bands_with_chunk_culling_on_a_large_scene(geom_path):
    """Band contexts skip whole chunks of 32 triangles by bounding sphere: many thin bands over a dense mesh
    (most chunks are culled in every band), rotations that tilt the mesh, even and odd widths, and the cull
    switched off (SLOTH_DEBUG=4 is read at context creation) must all give the whole frame."""
    xyz, rgb, s0 = meshes.icosphere(96)                       # 184,320 triangles, 5,760 chunks
    skull = S.soup("skull")
    for (scene, W, H, nb, ang) in [((xyz, rgb, s0), 640, 400, 16, (0.3, S.PI + 0.2, 0.1)),
                                   ((xyz, rgb, s0), 333, 250, 7, (1.1, 0.4, 2.0)),
                                   (skull, 300, 200, 9, (0.0, S.PI + 0.7, 0.0))]:
        rot = oracle.rotation(*ang)
        whole, _, _ = gpu_frame(*scene, W, H, rot)
        edges = [H * i // nb for i in range(nb + 1)]
        res = [gpu_frame(*scene, W, H, rot, band=(edges[i], edges[i + 1])) for i in range(nb)]
        banded = np.concatenate([r[0] for r in res] + [np.full(H, ord(" "), np.uint32)])
        bad = np.flatnonzero(banded != whole)
        assert bad.size == 0, f"{W}x{H}/{nb}: {bad.size} cells differ, first {bad[:8]}"
        n_chunks = (len(scene[0]) + 31) // 32
        done = [r[2]["chunks_processed"] for r in res]
        if scene[0] is xyz:   # spatially coherent chunks: the cull must really be in effect
            assert max(done) < n_chunks and sum(done) < nb * n_chunks // 2, (done, n_chunks)
    ocells, _, _ = oracle.render(xyz, rgb, s0, 333, 250, oracle.rotation(1.1, 0.4, 2.0), mode=1)
    assert np.array_equal(gpu_frame(xyz, rgb, s0, 333, 250, oracle.rotation(1.1, 0.4, 2.0))[0], ocells)


def test_batch_equals_single_frames_and_is_deterministic(geom_path):
    xyz, rgb, s0 = S.soup("pikachu")
    pitches = oracle.turntable(0.0, 12)
    rots = np.stack([oracle.rotation(0.0, p, 0.0) for p in pitches])
    ctx = rs.Context.blank(True)
    ctx.set_scene(xyz, rgb, s0)
    ctx.resize(200, 100)
    a = ctx.render_batch(rots).copy()
    b = ctx.render_batch(rots).copy()
    assert np.array_equal(a, b)
    for k in range(len(pitches)):
        ocells, _, _ = oracle.render(xyz, rgb, s0, 200, 100, rots[k], mode=1)
        assert np.array_equal(a[k], ocells), f"frame {k}"
    ctx.close()


def test_reference_style_host_api():
    """Reads like the reference's main loop (main.rs:76-89)."""
    xyz, rgb, s0 = S.soup("pikachu")
    sizes = S.mesh_sizes("pikachu")
    offs = np.concatenate([[0], np.cumsum(sizes)])
    mesh_queue = [rs.SimpleMesh(xyz[offs[i]:offs[i + 1]], rgb[offs[i]:offs[i + 1]]) for i in range(len(sizes))]
    context = rs.Context.blank(True)
    context.width, context.height = 80, 40            # match_dimensions
    rot = rs.rotation_from_euler(0.0, S.PI, 0.0)
    context.update((0, 0), mesh_queue)
    context.clear()
    for mesh in mesh_queue:
        rs.draw_mesh(context, mesh, rot, rs.default_shader)
    text = context.flush(False, False).decode("latin-1")
    gold = [c for c in S.golden()["cases"] if c["scene"] == "pikachu" and c["W"] == 80][0]
    assert text == gold["text"] + "\n"
    context.close()


def _write_obj(path, xyz):
    with open(path, "w") as f:
        for v in xyz.reshape(-1, 3):
            f.write("v %s %s %s\n" % tuple(repr(float(np.float32(c))) for c in v))
        for t in range(xyz.shape[0]):
            f.write("f %d %d %d\n" % (3 * t + 1, 3 * t + 2, 3 * t + 3))


def test_cli_drop_in_image_and_webify(tmp_path):
    """The C++ `sloth` binary: `image -w -h` text and the `-j` JS-frame export, byte for byte."""
    import subprocess
    from rust_sloth_b200 import turntable as tt
    exe = os.path.join(os.path.dirname(rs.LIB_PATH), "bin", "sloth")
    xyz, rgb, s0 = S.soup("pikachu")
    obj = str(tmp_path / "pikachu.obj")
    _write_obj(obj, xyz)
    white = np.ones_like(rgb)                       # no mtllib -> colour (1,1,1), geometry.rs:91
    out = subprocess.run([exe, obj, "-b", "image", "-w", "80", "-h", "40"], capture_output=True, check=True).stdout
    gold = [c for c in S.golden()["cases"] if c["scene"] == "pikachu" and c["W"] == 80][0]
    assert out.decode("latin-1") == gold["text"] + "\n"
    out = subprocess.run([exe, obj, "image", "-w", "64", "-h", "30", "-j", "5", "-x", "0.25"], capture_output=True, check=True).stdout
    frames = [oracle.render(xyz, white, s0, 64, 30, oracle.rotation(0.25, p, 0.0), mode=0)[0] for p in oracle.turntable(0.0, 5)]
    assert out == tt.webify_stream(frames)
    out = subprocess.run([exe, obj, "image", "-w", "31"], capture_output=True, check=True).stdout   # -h defaults to -w, colour on
    cells = oracle.render(xyz, white, s0, 31, 31, oracle.rotation(0.0, S.PI, 0.0), mode=0)[0]
    assert out == rs.flush_bytes(cells, True, False, True)


def test_band_renderer_single_process():
    """multigpu.BandRenderer with world=1 (device buffers, external stream) equals the host path."""
    import torch
    from rust_sloth_b200 import multigpu
    xyz, rgb, s0 = S.soup("skull")
    rot = oracle.rotation(0.0, S.PI, 0.0)
    ctx = rs.Context.blank(True)
    ctx.set_scene(xyz, rgb, s0)
    ocells, _, _ = oracle.render(xyz, rgb, s0, 320, 200, rot, mode=0)
    for mode in ("allgather", "peer"):
        # "peer": the frame lives in an exported cudaMalloc buffer and the band is resolved into it at its row
        # offset (with more ranks the same pointer arithmetic runs on an IPC mapping, profiles/band_8k.py)
        br = multigpu.BandRenderer(ctx, 320, 200, 0, 1, mode=mode, depth=3)
        assert br.mode == mode
        for _ in range(3):                                   # both frame regions get used
            frame = br.to_frame(br.render(rot))
            assert np.array_equal(frame, ocells), mode
        if mode == "peer":
            rot2 = oracle.rotation(0.1, S.PI + 0.5, 0.0)
            o2, _, _ = oracle.render(xyz, rgb, s0, 320, 200, rot2, mode=0)
            for _ in range(2):
                ptrs = br.render_batch(np.stack([rot, rot2, rot]))
                assert [np.array_equal(br.to_frame(p), o) for p, o in zip(ptrs, (ocells, o2, ocells))] == [True] * 3
        br.close()
    ctx.close()


def test_ipc_helpers_round_trip():
    """sloth_device_alloc / write / read / ipc_export on one process (opening needs a second process)."""
    ctx = rs.Context.blank(True)
    p = rs.device_alloc(ctx.device, 4 * 1000)
    cells = np.arange(1000, dtype=np.uint32)
    ctx.write_device(p, cells)
    assert np.array_equal(ctx.read_device(p + 4 * 10, 990), cells[10:])
    assert len(rs.ipc_export(ctx.device, p)) == rs.IPC_HANDLE_BYTES
    rs.device_free(ctx.device, p)
    ctx.close()


def test_newline_stamp_vs_wrapped_fragment_order(geom_path):
    """A fragment that wraps onto column 0/1 of a stamped row (2x == W or W+1): the cell shows the
    fragment only if its triangle is later than every triangle stamping that row."""
    rot = np.eye(4, dtype=np.float32).reshape(16)
    hits = 0
    for W, H in [(40, 20), (41, 21), (64, 48), (63, 47)]:
        for seed in range(30):
            rng = np.random.default_rng(1000 + seed)
            n = 40
            # triangles straddling x_screen = W/2 (object x around +1 with scene_max = 1) on many rows
            c = np.stack([rng.uniform(0.7, 1.3, n), rng.uniform(-0.9, 0.9, n), rng.uniform(-1, 1, n)], axis=1)
            xyz = (c[:, None, :] + rng.uniform(-0.25, 0.25, (n, 3, 3))).reshape(n, 9).astype(np.float32)
            rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
            ocells, oz, _ = oracle.render(xyz, rgb, 1.0, W, H, rot, mode=0)
            cells, z, st = gpu_frame(xyz, rgb, np.float32(1.0), W, H, rot)
            assert_same(cells, z, ocells, oz, f"stamp order {W}x{H} seed={seed}")
            hits += st["stamp_fixups"]
    assert hits > 0   # the exact same-chunk check was exercised


def test_one_column_frames_keep_their_newline_stamps(geom_path):
    """W == 1: no x candidate exists, but the newline stamp of row y is cell y*W+1 = (row y+1, column 0)
    (found by the hypothesis fuzz: nine degenerate triangles and one that spans a row, 1x4)."""
    tris = np.zeros((10, 9), np.float32)
    tris[9, 7] = -1.0
    rgb = np.full((10, 3), 7, np.uint8)
    rot = oracle.rotation(0.0, 0.0, 0.0)
    ocells, oz, _ = oracle.render(tris, rgb, 1.0, 1, 4, rot, image=True, mode=0)
    assert ocells[3] == ord("\n")
    cells, z, _ = gpu_frame(tris, rgb, np.float32(1.0), 1, 4, rot)
    assert_same(cells, z, ocells, oz, "1x4")
    for seed in range(12):
        xyz, rgb, s0 = meshes.random_soup(seed, 30)
        for H in (2, 5, 9, 40):
            rot = oracle.rotation(0.1 * seed, S.PI + 0.3 * seed, 0.0)
            for image in (True, False):
                ocells, oz, _ = oracle.render(xyz, rgb, s0, 1, H, rot, image=image, mode=0)
                cells, z, _ = gpu_frame(xyz, rgb, s0, 1, H, rot, image=image)
                assert_same(cells, z, ocells, oz, f"1x{H} seed {seed}")
            # the same frame in row bands: a stamp lands in the band below the row that was stamped
            ocells, _, _ = oracle.render(xyz, rgb, s0, 1, H, rot, image=True, mode=0)
            nb = min(3, H)
            edges = [H * i // nb for i in range(nb + 1)]
            parts = [gpu_frame(xyz, rgb, s0, 1, H, rot, band=(edges[i], edges[i + 1]))[0] for i in range(nb)]
            banded = np.concatenate(parts + [np.full(H, ord(" "), np.uint32)])
            assert np.array_equal(banded, ocells), f"1x{H} seed {seed} in {nb} bands"


def test_tall_frame_uses_global_row_stamps(geom_path):
    xyz, rgb, s0 = S.soup("suzy")
    rot = oracle.rotation(0.0, S.PI, 0.0)
    W, H = 16, 9000   # H > 8192: the row-stamp array no longer fits in shared memory
    ocells, oz, _ = oracle.render(xyz, rgb, s0, W, H, rot, mode=1)
    cells, z, _ = gpu_frame(xyz, rgb, s0, W, H, rot)
    assert_same(cells, z, ocells, oz, "tall frame")


def test_device_batch_overlap_equals_single_frames(geom_path):
    """sloth_render_device_batch (geometry k+1 overlapping resolve k on two streams) == frame by frame."""
    import torch
    xyz, rgb, s0 = meshes.icosphere(40)
    pitches = oracle.turntable(0.0, 9)
    rots = np.stack([oracle.rotation(0.1, p, 0.0) for p in pitches])
    for (W, H) in [(320, 200), (161, 83)]:
        ctx = rs.Context.blank(True)
        ctx.set_scene(xyz, rgb, s0)
        ctx.resize(W, H)
        cpf = ctx.cells_per_frame()
        stride = (cpf + 3) & ~3
        buf = torch.zeros(len(rots) * stride, dtype=torch.int32, device="cuda")
        ctx.render_device_batch(rots, buf.data_ptr(), stride)
        ctx.sync()
        got = buf.cpu().numpy().view(np.uint32).reshape(len(rots), stride)[:, :cpf]
        for k in range(len(rots)):
            single, _ = ctx.render(rots[k])
            assert np.array_equal(got[k], single), f"{W}x{H} frame {k}"
            ocells, _, _ = oracle.render(xyz, rgb, s0, W, H, rots[k], mode=1)
            assert np.array_equal(single, ocells)
        ctx.close()


def test_device_flush_is_byte_exact():
    """GPU-side Context::flush (plain / ANSI / <span>) == the host formatter, byte for byte."""
    xyz, rgb, s0 = S.soup("vaporeon")                      # per-triangle colours: all digit counts occur
    rng = np.random.default_rng(7)
    rgb = rng.choice(np.array([0, 5, 9, 10, 42, 99, 100, 163, 255], np.uint8), size=rgb.shape)
    pitches = oracle.turntable(0.0, 5)
    rots = np.stack([oracle.rotation(0.0, p, 0.2) for p in pitches])
    for (W, H) in [(200, 100), (131, 77), (33, 5)]:
        ctx = rs.Context.blank(True)
        ctx.set_scene(xyz, rgb, s0)
        ctx.resize(W, H)
        frames = ctx.render_batch(rots).copy()
        for mode, (color, webify) in {0: (False, False), 1: (True, False), 2: (True, True)}.items():
            texts = ctx.render_text_batch(rots, mode)
            for k in range(len(rots)):
                want = rs.flush_bytes(frames[k], color, webify, True)
                if mode == 0:
                    want = want[:-1]                       # println's newline is the host's
                assert texts[k] == want, f"{W}x{H} mode {mode} frame {k}"
        ctx.close()


def test_error_codes_and_call_order():
    ctx = rs.Context.blank(True)
    rot = oracle.rotation(0.0, S.PI, 0.0)
    with pytest.raises(rs.SlothError) as e:
        ctx.render(rot)
    assert e.value.code == rs.SLOTH_E_STATE                       # no scene yet
    xyz, rgb, s0 = S.soup("cube")
    ctx.set_scene(xyz, rgb, s0)
    ctx.width, ctx.height = 0, 0
    with pytest.raises(rs.SlothError) as e:
        ctx.render(rot)
    assert e.value.code == rs.SLOTH_E_STATE                       # no size yet
    for (W, H, code) in [(0, 10, rs.SLOTH_E_ARG), (70000, 10, rs.SLOTH_E_TOO_LARGE), (65535, 65535, rs.SLOTH_E_TOO_LARGE)]:
        with pytest.raises(rs.SlothError) as e:
            ctx.resize(W, H)
        assert e.value.code == code
    ctx.resize(40, 20)
    with pytest.raises(rs.SlothError) as e:
        ctx.set_band(10, 30)
    assert e.value.code == rs.SLOTH_E_ARG
    cells, _ = ctx.render(rot)
    ocells, _, _ = oracle.render(xyz, rgb, s0, 40, 20, rot)
    assert np.array_equal(cells, ocells)                          # the context still works after the errors
    ctx.close()
    with pytest.raises(rs.SlothError) as e:
        rs.Context.blank(True, device=999)
    assert e.value.code == rs.SLOTH_E_ARG


def test_full_size_baseline_configs_bit_exact():
    """BASELINE configs at their full size: 'suzy suzy' at 3840x2160 (every fragment of the second copy is an exact
    depth tie) and skull at 1920x1080 against the faithful scan.  (The 10M-triangle frame, the 360-frame turntable
    and the two-model scene are in test_gpu_configs.py.)"""
    for scene, W, H in [("suzy_suzy", 3840, 2160), ("skull", 1920, 1080)]:
        xyz, rgb, s0 = S.soup(scene)
        rot = oracle.rotation(0.0, S.PI, 0.0)
        ocells, oz, ocnt = oracle.render(xyz, rgb, s0, W, H, rot, mode=0)
        cells, z, st = gpu_frame(xyz, rgb, s0, W, H, rot)
        assert_same(cells, z, ocells, oz, f"{scene} {W}x{H}")
        assert st["fragments"] == ocnt["covered"]
    # depth-tie property of "suzy suzy": the second copy never wins a cell
    xyz1, rgb1, s01 = S.soup("suzy")
    one, _, _ = gpu_frame(xyz1, rgb1, s01, 3840, 2160, rot)
    two, _, _ = gpu_frame(np.concatenate([xyz1, xyz1]), np.concatenate([rgb1, 255 - rgb1]), s01, 3840, 2160, rot)
    assert np.array_equal(one, two)


def test_tma_feed_variant_is_bit_exact(monkeypatch):
    """SLOTH_TMA=1: k_geom3 fed by cp.async.bulk + mbarrier (off by default: measured 3 % slower)."""
    monkeypatch.setenv("SLOTH_TMA", "1")
    for f, (W, H) in [(24, (160, 80)), (90, (640, 360))]:
        xyz, rgb, s0 = meshes.icosphere(f)
        for k in range(2):
            rot = oracle.rotation(0.3 * k, np.float32(np.pi) + 0.5 * k, 0.0)
            ocells, oz, ocnt = oracle.render(xyz, rgb, s0, W, H, rot, mode=1)
            cells, z, st = gpu_frame(xyz, rgb, s0, W, H, rot)
            assert_same(cells, z, ocells, oz, f"tma icosphere f={f}")
            assert st["fragments"] == ocnt["covered"]
    xyz, rgb, s0 = meshes.random_soup(11, 50)
    rot = oracle.rotation(0.2, 3.0, 0.1)
    ocells, oz, _ = oracle.render(xyz, rgb, s0, 101, 57, rot, mode=0)
    cells, z, _ = gpu_frame(xyz, rgb, s0, 101, 57, rot)
    assert_same(cells, z, ocells, oz, "tma fuzz")


from hypothesis import given, settings, strategies as st, HealthCheck

_coord = st.one_of(
    st.floats(min_value=-1.5, max_value=1.5, width=32),
    st.sampled_from([0.0, -0.0, 1.0, -1.0, 0.5, 0.25, 1e-30, -1e-30, 1e-45, 3e38, -3e38, float("inf"), float("-inf"), float("nan")]))


@settings(max_examples=int(os.environ.get("HYP_EXAMPLES", "120")), deadline=None, suppress_health_check=list(HealthCheck))
@given(tris=st.lists(st.lists(_coord, min_size=9, max_size=9), min_size=1, max_size=40),
       W=st.integers(min_value=1, max_value=70), H=st.integers(min_value=1, max_value=40),
       angles=st.tuples(st.floats(-7, 7, width=32), st.floats(-7, 7, width=32), st.floats(-7, 7, width=32)),
       image=st.booleans(), s0=st.sampled_from([1.0, 0.7, 2.5, 0.0]))
def test_hypothesis_fuzz_gpu_vs_oracle(tris, W, H, angles, image, s0, geom_path):
    """Arbitrary small soups incl. NaN / inf / denormal / huge coordinates, 1-cell frames, odd widths."""
    xyz = np.array(tris, np.float32)
    rgb = (np.arange(xyz.shape[0] * 3, dtype=np.int64) * 37 % 256).astype(np.uint8).reshape(-1, 3)
    rot = oracle.rotation(*angles)
    ocells, oz, ocnt = oracle.render(xyz, rgb, s0, W, H, rot, image=image, mode=0)
    cells, z, st_ = gpu_frame(xyz, rgb, np.float32(s0), W, H, rot, image=image)
    assert_same(cells, z, ocells, oz, "hypothesis")
    assert st_["fragments"] == ocnt["covered"]


def test_indexed_scene_dedupes_by_bit_pattern_and_auto_picks_the_path():
    """sloth_scene_set deduplicates corners (index.cuh): an icosphere of frequency f has 10 f^2 + 2 distinct vertices;
    AUTO takes the per-vertex path for shared meshes and keeps the soup path for soups without sharing;
    sloth_scene_set_indexed (positions + indices, geometry.rs:99-107) ends in the same frames."""
    f = 24
    xyz, rgb, s0 = meshes.icosphere(f)
    W, H = 200, 100
    rot = oracle.rotation(0.2, np.float32(np.pi) + 0.4, 0.1)
    ocells, oz, _ = oracle.render(xyz, rgb, s0, W, H, rot, image=True, mode=0)
    ctx = rs.Context.blank(True)
    try:
        ctx.set_scene(xyz, rgb, s0)
        ctx.resize(W, H)
        st = ctx.stats()
        assert st["geom_path"] == rs.PATH_INDEXED and st["n_vert"] == 10 * f * f + 2
        cells, z = ctx.render(rot, want_z=True)
        assert_same(cells, z, ocells, oz, "icosphere, AUTO -> indexed")
        # the caller's own indexing: unique rows of the soup + inverse map
        corners = np.ascontiguousarray(xyz.reshape(-1, 3))
        rows = corners.view(np.uint32).view([("x", np.uint32), ("y", np.uint32), ("z", np.uint32)]).reshape(-1)
        _, first, inv = np.unique(rows, return_index=True, return_inverse=True)
        inv = inv.reshape(-1)
        assert np.array_equal(corners[first][inv].view(np.uint32), corners.view(np.uint32))
        ctx.set_scene_indexed(corners[first], inv.reshape(-1, 3), rgb, s0)
        assert ctx.stats()["n_vert"] == 10 * f * f + 2
        cells2, z2 = ctx.render(rot, want_z=True)
        assert_same(cells2, z2, ocells, oz, "icosphere, set_scene_indexed")
        dxyz, drgb, _ = ctx.scene()
        assert np.array_equal(dxyz.view(np.uint32), xyz.view(np.uint32)) and np.array_equal(drgb, rgb)
        with pytest.raises(rs.SlothError) as e:
            bad = inv.reshape(-1, 3).copy()
            bad[7, 1] = len(first)
            ctx.set_scene_indexed(corners[first], bad, rgb, s0)
        assert e.value.code == rs.SLOTH_E_ARG
        # -0.0 and +0.0 are different vertices (bit patterns), NaN payloads too: nothing is merged by value
        tri = np.array([[0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0],
                        [-0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0]], np.float32)
        ctx.set_path(rs.PATH_INDEXED)
        ctx.set_scene(tri, rgb[:2], 1.0)
        assert ctx.stats()["n_vert"] == 4
        # a soup without sharing stays on the soup path under AUTO
        ctx.set_path(rs.PATH_AUTO)
        sxyz, srgb, ss0 = meshes.random_soup(3, 500)
        ctx.set_scene(sxyz, srgb, ss0)
        assert ctx.stats()["geom_path"] == rs.PATH_SOUP and ctx.stats()["n_vert"] == 0
    finally:
        ctx.close()


def test_binned_tile_path_takes_the_large_triangles(monkeypatch):
    """Config 3 ('suzy suzy' at 3840x2160, every fragment of the second copy an exact depth tie): the triangles the
    geometry kernel queues are binned to 32 x 8 tiles and rasterised per tile (tile_kernels.cuh); the frame, the
    z-buffer and the fragment count equal the oracle's, and equal the walk kernel's with the tile path off."""
    xyz, rgb, s0 = S.soup("suzy_suzy")
    rot = oracle.rotation(0.0, S.PI, 0.0)
    ocells, oz, ocnt = oracle.render(xyz, rgb, s0, 3840, 2160, rot, mode=0)
    monkeypatch.setenv("SLOTH_TILES", "1")     # the tile path is opt-in (measured slower than the walk kernel)
    cells, z, st = gpu_frame(xyz, rgb, s0, 3840, 2160, rot)
    assert st["walk_tris"] > 500 and st["tile_tris"] > 0.9 * st["walk_tris"] and st["tiles_used"] > 1000, st
    assert_same(cells, z, ocells, oz, "suzy suzy 4K, tile path")
    assert st["fragments"] == ocnt["covered"]
    monkeypatch.setenv("SLOTH_TILES", "0")
    cells0, z0, st0 = gpu_frame(xyz, rgb, s0, 3840, 2160, rot)
    assert st0["tile_tris"] == 0 and np.array_equal(cells0, cells) and np.array_equal(z0, z)
    monkeypatch.setenv("SLOTH_TILES", "1")
    # wrap frame of the Pikachu turntable at a size where the tile path is on by default, odd width as well
    xyz, rgb, s0 = S.soup("pikachu")
    rot = oracle.rotation(0.0, oracle.turntable(0.0, 360)[144], 0.0)
    for W, H in ((1280, 640), (1281, 641)):
        ocells, oz, ocnt = oracle.render(xyz, rgb, s0, W, H, rot, mode=0)
        cells, z, st = gpu_frame(xyz, rgb, s0, W, H, rot)
        assert st["tile_tris"] > 0
        assert_same(cells, z, ocells, oz, f"pikachu wrap frame {W}x{H}, tile path")
        assert st["fragments"] == ocnt["covered"]


def test_span_wire_returns_the_same_cells():
    """SLOTH_WIRE_SPANS (run lists over PCIe, cells rebuilt by host threads inside the library) hands back the same
    buffers as the plain wire: single frames, batches (staging slots wrap), image and interactive mode, odd widths,
    a resize in between, and a noise frame that has no runs (sent as plain cells)."""
    rng = np.random.default_rng(5)
    centre = (rng.random((60000, 1, 3), np.float32) * 1.9 - 0.95).astype(np.float32)   # confetti: runs of 2 cells
    noise = (centre + (rng.random((60000, 3, 3), np.float32) * 0.05 - 0.025).astype(np.float32)).reshape(-1, 9)
    noise_rgb = rng.integers(1, 255, size=(60000, 3)).astype(np.uint8)
    cases = [("pikachu", *S.soup("pikachu"), True, 200, 100), ("skull", *S.soup("skull"), False, 321, 120),
             ("icosphere", *meshes.icosphere(40), True, 640, 360), ("noise", noise, noise_rgb, np.float32(1.0), True, 300, 200)]
    for name, xyz, rgb, s0, image, W, H in cases:
        ctx = rs.Context.blank(image)
        try:
            ctx.set_scene(xyz, rgb, s0)
            ctx.resize(W, H)
            rots = np.stack([oracle.rotation(0.1 * k, S.PI + 0.37 * k, 0.05 * k) for k in range(17)])
            plain = ctx.render_batch(rots).copy()
            ctx.set_wire(rs.WIRE_SPANS)
            spans = ctx.render_batch(rots)
            assert np.array_equal(spans, plain), name
            one, _ = ctx.render(rots[3])
            assert np.array_equal(one, plain[3]), name
            st = ctx.wire_stats()
            assert st["frames"] == 18 and st["threads"] >= 1
            plain_bytes = 18 * plain.shape[1] * 4
            if name == "noise":
                assert st["plain_frames"] > 0
            else:
                assert st["plain_frames"] == 0 and st["d2h_bytes"] < plain_bytes // 4, (name, st, plain_bytes)
            ocells, _, _ = oracle.render(xyz, rgb, s0, W, H, rots[3], image=image, mode=0)
            assert np.array_equal(one, ocells), name
            ctx.resize(W // 2 + 1, H // 2)                      # other run-list capacity, odd / even flips
            small = ctx.render_batch(rots[:5])
            ctx.set_wire(rs.WIRE_CELLS)
            assert np.array_equal(small, ctx.render_batch(rots[:5])), name
        finally:
            ctx.close()


def test_super_chunk_certificate_under_general_matrices():
    """k_super_cert skips whole super-chunks (128 triangles) it can prove back-facing for the frame's matrix.  The
    caller's `transform` is any 4x4 (draw_mesh takes it as is, rasterizer.rs:39-46): rotations, anisotropic scale,
    shear, mirror images (which turn the back faces into front faces), translations that push the model half off the
    frame, a collapsed axis.  Every frame must equal the oracle's, newline stamps included, and the plain rotations
    must actually skip something."""
    rng = np.random.default_rng(17)
    xyz, rgb, s0 = meshes.icosphere(40)                  # 32 000 triangles = 250 super-chunks, ~1 cell each at 640x360
    n_chunks = (len(xyz) + 31) // 32
    ctx = rs.Context.blank(True)
    try:
        ctx.set_scene(xyz, rgb, s0)
        ctx.resize(640, 360)
        ctx.stats_enable(count_fragments=True)
        skipped_some = 0
        for k in range(14):
            rot = oracle.rotation(*(rng.random(3) * 6.3)).reshape(4, 4).astype(np.float64)   # column-major
            if k >= 3:
                lin = np.diag(rng.uniform(0.3, 1.4, 3))
                lin[0, 1] = rng.uniform(-0.6, 0.6)                                   # shear
                if k % 3 == 0:
                    lin[2, 2] = -lin[2, 2]                                           # mirror image
                if k == 9:
                    lin[1, :] = 0.0                                                  # collapsed axis: zero-area triangles
                rot[:3, :3] = rot[:3, :3] @ lin
                rot[3, :3] = rng.uniform(-0.7, 0.7, 3)                               # translation (last row, column-major)
            rot = rot.astype(np.float32).reshape(16)
            ocells, oz, ocnt = oracle.render(xyz, rgb, s0, 640, 360, rot, image=True, mode=0)
            cells, z = ctx.render(rot, want_z=True)
            st = ctx.stats()
            assert_same(cells, z, ocells, oz, f"matrix {k}")
            assert st["fragments"] == ocnt["covered"], k
            if st["chunks_processed"] < n_chunks:
                skipped_some += 1
        assert skipped_some >= 3, skipped_some
    finally:
        ctx.close()


def test_left_edge_triangles_and_list_tails():
    """Triangles whose scan domain ends within three columns of the left edge (the footprint of k_tri hands them to the
    general path), and scenes whose triangle count leaves chunks behind the last full super-chunk (the static head of
    k_tri's flat work list) or has no full super-chunk at all: every frame must equal the oracle's."""
    rng = np.random.default_rng(5)
    ctx = rs.Context.blank(True)
    try:
        for freq, W, H in [(40, 640, 360), (7, 200, 120), (3, 64, 48), (13, 321, 200)]:
            xyz, rgb, s0 = meshes.icosphere(freq)          # 20 f^2 triangles: 32 000 / 980 / 180 / 3 380
            ctx.set_scene(xyz, rgb, s0)
            ctx.resize(W, H)
            ctx.stats_enable(count_fragments=True)
            for k, tx in enumerate([-1.0, -0.99, -0.97, -0.9, 0.0]):
                rot = oracle.rotation(*(rng.random(3) * 6.3)).reshape(4, 4).astype(np.float64)   # column-major
                rot[3, 0] = tx                              # the sphere's centre lands on x' = (1 + tx) W / 4: at the edge
                rot = rot.astype(np.float32).reshape(16)
                ocells, oz, ocnt = oracle.render(xyz, rgb, s0, W, H, rot, image=True, mode=0)
                cells, z = ctx.render(rot, want_z=True)
                assert_same(cells, z, ocells, oz, f"f={freq} tx={tx}")
                assert ctx.stats()["fragments"] == ocnt["covered"], (freq, tx)
    finally:
        ctx.close()
