"""CPU tests of the oracle itself: the reference's own known-answer vectors
(src/geometry.rs:196-234), the committed golden frames, the independent numpy
restatement, and the quirks of SURVEY.md Appendix A (ties, wrap, newline, NaN)."""
import hashlib

import numpy as np
import pytest

import oracle
from oracle import restate_np
import scenes as S

F = np.float32


def v4(*a):
    return np.array(a, F)


# ---- the three unit tests of the reference -----------------------------------------------
DEFAULT_TRI = np.concatenate([v4(1, -1, -1, 1), v4(-1, -1, 1, 1), v4(1, 1, -1, 1)])  # geometry.rs:25-34


def _aabb(tri):
    mn, mx = np.empty(4, F), np.empty(4, F)
    L = oracle.lib()
    L.oracle_triangle_aabb(oracle._fp(tri), oracle._fp(mn), oracle._fp(mx))
    return mn, mx


def test_reference_test_aabb():  # geometry.rs:196-206
    mn, mx = _aabb(DEFAULT_TRI.copy())
    assert np.array_equal(mn, v4(-1, -1, -1, 1)) and np.array_equal(mx, v4(1, 1, 1, 1))


def test_reference_test_transform():  # geometry.rs:208-220
    T = np.eye(4, dtype=F)
    T[:3, 3] = 1.0  # Matrix4::new_translation((1,1,1))
    tri = DEFAULT_TRI.copy()
    oracle.lib().oracle_triangle_mul(oracle._fp(np.ascontiguousarray(T.T).reshape(16)), oracle._fp(tri))
    mn, mx = _aabb(tri)
    assert np.array_equal(mn, v4(0, 0, 0, 1)) and np.array_equal(mx, v4(2, 2, 2, 1))


def test_reference_test_normal():  # geometry.rs:222-234
    tri = np.concatenate([v4(-1, 1, 0, 1), v4(0, 1, 1, 1), v4(1, 1, 0, 1)])
    n, ref = np.empty(4, F), np.empty(4, F)
    oracle.lib().oracle_triangle_normal(oracle._fp(tri), oracle._fp(n))
    oracle.lib().oracle_vec4_normalize(oracle._fp(v4(0, 1, 0, 0)), oracle._fp(ref))
    assert np.array_equal(n, ref) and np.array_equal(n, v4(0, 1, 0, 0))


# ---- golden frames -----------------------------------------------------------------------
@pytest.mark.parametrize("case", [c for c in S.golden()["cases"] if c["W"] * c["H"] <= 1920 * 1080 and c["scene"] != "hand"],
                         ids=lambda c: f"{c['scene']}-{c['W']}x{c['H']}")
def test_oracle_reproduces_committed_golden_frames(case):
    xyz, rgb, s0 = S.soup(case["scene"])
    rot = oracle.rotation(case["roll"], case["pitch"], case["yaw"])
    cells, z, cnt = oracle.render(xyz, rgb, s0, case["W"], case["H"], rot, image=True, mode=0)
    assert hashlib.sha256(cells.tobytes()).hexdigest() == case["cells_sha256"]
    assert hashlib.sha256(z.tobytes()).hexdigest() == case["z_sha256"]
    assert cnt == case["counters"]
    # the row-terminating variant (used for sizes where mode 0 takes minutes) gives the same frame
    c1, z1, cnt1 = oracle.render(xyz, rgb, s0, case["W"], case["H"], rot, image=True, mode=1)
    assert np.array_equal(c1, cells) and np.array_equal(z1, z)
    assert cnt1["covered"] == cnt["covered"] and cnt1["zwrites"] == cnt["zwrites"]


def test_survey_probe_counts():
    """counts produced during the survey by a separate throwaway
    float32 restatement (candidates, covered, z-writes)."""
    expect = {("pikachu", 80, 40): (54451, 477, 432), ("skull", 1920, 1080): (49841499, 417298, 373020),
              ("pikachu", 1920, 1080): (31459027, 273828, 249391), ("pikachu", 160, 80): (261717, 1844, 1811)}
    for c in S.golden()["cases"]:
        key = (c["scene"], c["W"], c["H"])
        if key in expect:
            k = c["counters"]
            assert (k["candidates"], k["covered"], k["zwrites"]) == expect[key]


def test_pikachu_80x40_looks_like_the_committed_text():
    case = [c for c in S.golden()["cases"] if c["scene"] == "pikachu" and c["W"] == 80][0]
    rows = case["text"].split("\n")
    assert len(case["text"]) == 80 * 40 + 40
    assert sum(ch not in " \n" for ch in case["text"]) == 2 * 416  # 416 visible ids, two cells each


# ---- independent numpy restatement ----------------------------------------------------------
@pytest.mark.parametrize("scene,W,H,angles", [
    ("cube", 40, 20, (0.5, 4.0, 0.25)), ("suzy", 50, 31, (0.0, S.PI, 0.0)), ("cube_stl", 37, 20, (0.4, 3.5, 0.1)),
    ("ferris", 60, 30, (0.1, 2.0, 0.3)), ("part_stl", 48, 24, (0.0, S.PI, 0.0))])
def test_numpy_restatement_agrees_bit_for_bit(scene, W, H, angles):
    xyz, rgb, s0 = S.soup(scene)
    rot = oracle.rotation(*angles)
    cells, z, _ = oracle.render(xyz, rgb, s0, W, H, rot, image=True, mode=0)
    c2, z2 = restate_np.render(xyz, rgb, s0, W, H, rot, image=True)
    assert np.array_equal(cells, c2)
    assert np.array_equal(z.view(np.uint32), z2.view(np.uint32))


def test_numpy_restatement_on_fuzz_and_wrap():
    from rust_sloth_b200 import meshes
    for seed, kind in enumerate(["uniform", "small", "sliver", "collinear", "dup", "axis"]):
        xyz, rgb, s0 = meshes.random_soup(100 + seed, 24, kind=kind)
        rot = oracle.rotation(0.3 * seed, 3.0 + seed, 0.1)
        for (W, H, image) in [(33, 20, True), (40, 21, False)]:
            cells, z, _ = oracle.render(xyz, rgb, s0, W, H, rot, image=image, mode=0)
            c2, z2 = restate_np.render(xyz, rgb, s0, W, H, rot, image=image)
            assert np.array_equal(cells, c2) and np.array_equal(z.view(np.uint32), z2.view(np.uint32))


def test_matrix_helpers_agree():
    rot = oracle.rotation(0.3, 3.3, -0.7)
    R = rot.reshape(4, 4).T
    T = restate_np.utransform(80, 40, 7.5)
    assert np.array_equal(oracle.utransform(80, 40, 7.5).reshape(4, 4).T, T)
    assert np.array_equal(oracle.mat4_mul(oracle.utransform(80, 40, 7.5), rot).reshape(4, 4).T, restate_np.matmul44(T, R))


# ---- quirks -----------------------------------------------------------
def _one(xyz, rgb, W, H, image=True, s0=1.0, rot=None, mode=0):
    rot = np.eye(4, dtype=F).reshape(16) if rot is None else rot
    return oracle.render(np.asarray(xyz, F), np.asarray(rgb, np.uint8), s0, W, H, rot, image=image, mode=mode)


def test_exact_depth_ties_first_triangle_wins():
    tri = [-0.5, -0.5, 0.2, -0.5, 0.6, 0.2, 0.7, -0.5, 0.2]   # front-facing after the y flip of utransform
    cells, z, cnt = _one([tri, tri], [[10, 20, 30], [200, 100, 50]], 40, 20)
    assert cnt["covered"] > 0 and cnt["covered"] == 2 * cnt["zwrites"]
    drawn = cells[(cells & 0xFF) != ord(" ")]
    drawn = drawn[(drawn & 0xFF) != ord("\n")]
    assert np.all((drawn >> 8) == (10 | 20 << 8 | 30 << 16))


def test_double_cell_write_and_newline_column():
    tri = [-0.9, -0.9, 0.0, -0.9, 0.9, 0.0, 0.9, -0.9, 0.0]
    cells, z, cnt = _one([tri], [[1, 2, 3]], 40, 20)
    grid = cells[:40 * 20].reshape(20, 40)
    ids = np.flatnonzero(z != oracle.F32_MAX)
    assert np.all(ids % 2 == 0)                                  # id = y*W + 2x
    assert np.array_equal(cells[ids], cells[ids + 1])            # both cells of the pair
    rows = np.unique(ids // 40)
    assert np.all((grid[rows, 1] & 0xFF) == ord("\n"))           # rasterizer.rs:89-91
    assert np.all(cells[40 * 20:] == ord(" "))                   # context.rs:38-39 tail stays blank
    cells_i, _, _ = _one([tri], [[1, 2, 3]], 40, 20, image=False)
    assert cells_i.size == 40 * 20 and not np.any((cells_i & 0xFF) == ord("\n"))


def test_row_wrap_lands_on_next_row():
    # x_screen >= W/2 happens when a rotated x exceeds scale0: ids run past the end of the row
    tri = [1.2, -0.4, 0.0, 1.2, 0.4, 0.0, 1.9, -0.4, 0.0]
    W, H = 40, 20
    cells, z, cnt = _one([tri], [[9, 9, 9]], W, H, image=False)
    ids = np.flatnonzero(z != oracle.F32_MAX)
    assert cnt["covered"] > 0 and ids.size > 0
    ut = oracle.utransform(W, H, 1.0).reshape(4, 4).T
    ys = ut[1, 1] * np.array([-0.4, 0.4], F) + ut[1, 3]
    assert (ids // W).max() > np.ceil(ys.max()) - 1               # some fragment sits one row below its source row


def test_nan_and_inf_never_win_but_neg_inf_does():
    base = [-0.5, -0.5, 0.0, -0.5, 0.6, 0.0, 0.7, -0.5, 0.0]
    for bad, wins in [(np.nan, False), (np.inf, False)]:
        tri = list(base)
        tri[2] = bad
        cells, z, cnt = _one([tri], [[5, 5, 5]], 40, 20, image=False)
        assert cnt["zwrites"] == 0 and np.all(cells == ord(" "))


def test_shader_thresholds_are_inclusive():
    xyz, rgb, s0 = S.soup("cube")
    rot = oracle.rotation(0.5, 4.0, 0.25)
    cells, _, _ = oracle.render(xyz, rgb, s0, 80, 40, rot)
    glyphs = set(bytes((cells & 0xFF).astype(np.uint8)).decode())
    assert glyphs <= set(".:-=+*#%@ \n")


def test_turntable_sequence():
    g = S.golden()
    p = oracle.turntable(0.0, 360)
    assert len(p) == g["turntable_360_count"] == 360
    assert [float(x) for x in p[:8]] == g["turntable_360_first8"]
    pi = F(np.pi)
    step = F(F(2.0) * pi) * F(F(1.0) / F(360))
    acc = pi
    for k in range(360):
        assert p[k] == acc
        acc = F(acc + step)
    assert len(oracle.turntable(0.0, 1)) == 1
    assert len(oracle.turntable(5.0, 100)) < 100                 # stops when pitch exceeds 9.42477 (main.rs:99)


def test_sampling_partitions_the_triangle_list():
    xyz, rgb, s0 = S.soup("suzy")
    rot = oracle.rotation(0.0, S.PI, 0.0)
    _, _, full = oracle.render(xyz, rgb, s0, 120, 60, rot)
    parts = [oracle.render(xyz, rgb, s0, 120, 60, rot, tri_first=i, tri_step=4)[2] for i in range(4)]
    assert sum(p["candidates"] for p in parts) == full["candidates"]
    assert sum(p["covered"] for p in parts) == full["covered"]


# ---- property fuzz (hypothesis): the three CPU statements of the algorithm agree -----------------
from hypothesis import given, settings, strategies as st, HealthCheck

_coord = st.one_of(
    st.floats(min_value=-1.5, max_value=1.5, width=32),
    st.sampled_from([0.0, -0.0, 1.0, -1.0, 0.5, 0.25, 1e-30, -1e-30, 1e-45, 3e38, -3e38, float("inf"), float("-inf"), float("nan")]))


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(tris=st.lists(st.lists(_coord, min_size=9, max_size=9), min_size=1, max_size=6),
       W=st.integers(min_value=1, max_value=40), H=st.integers(min_value=1, max_value=24),
       angles=st.tuples(st.floats(-7, 7, width=32), st.floats(-7, 7, width=32), st.floats(-7, 7, width=32)),
       image=st.booleans(), s0=st.sampled_from([1.0, 0.7, 2.5, 0.0]))
def test_fuzz_three_statements_agree(tris, W, H, angles, image, s0):
    """mode 0 (faithful scan) == mode 1 (row termination; falls back to mode 0 for non-regular triangles)
    == the numpy restatement, including NaN/inf/denormal coordinates, 1-cell frames and odd widths."""
    xyz = np.array(tris, np.float32)
    rgb = (np.arange(xyz.shape[0] * 3, dtype=np.int64) * 37 % 256).astype(np.uint8).reshape(-1, 3)
    rot = oracle.rotation(*angles)
    c0, z0, k0 = oracle.render(xyz, rgb, s0, W, H, rot, image=image, mode=0)
    c1, z1, k1 = oracle.render(xyz, rgb, s0, W, H, rot, image=image, mode=1)
    assert np.array_equal(c0, c1) and np.array_equal(z0.view(np.uint32), z1.view(np.uint32))
    assert k0["covered"] == k1["covered"] and k0["zwrites"] == k1["zwrites"]
    c2, z2 = restate_np.render(xyz, rgb, s0, W, H, rot, image=image)
    assert np.array_equal(c0, c2) and np.array_equal(z0.view(np.uint32), z2.view(np.uint32))


def test_narrow_frames_agree_between_the_two_restatements():
    """W = 1, 2, 3: no or hardly any x candidates, but newline stamps still land at cell y*W+1 -- for W = 1 that is
    column 0 of the NEXT row (rasterizer.rs:89-91 indexes the flat buffer).  Both restatements must agree there;
    this is the behaviour the GPU path was fixed to follow (tests/test_gpu_parity.py::test_one_column_frames_*)."""
    from rust_sloth_b200 import meshes
    tris = np.zeros((10, 9), np.float32)
    tris[9, 7] = -1.0                                     # the hypothesis counter-example: one triangle spanning row 2
    rgb = np.full((10, 3), 7, np.uint8)
    cells, _, _ = oracle.render(tris, rgb, 1.0, 1, 4, oracle.rotation(0.0, 0.0, 0.0), image=True, mode=0)
    assert list(cells) == [32, 32, 32, 10, 32, 32, 32, 32]
    stamped = 0
    for seed in range(8):
        xyz, rgb, s0 = meshes.random_soup(seed, 30)
        rot = oracle.rotation(0.1 * seed, 3.1 + 0.3 * seed, 0.0)
        for W in (1, 2, 3):
            for H in (2, 5, 9, 40):
                for image in (True, False):
                    cells, z, _ = oracle.render(xyz, rgb, s0, W, H, rot, image=image, mode=0)
                    c2, z2 = restate_np.render(xyz, rgb, s0, W, H, rot, image=image)
                    assert np.array_equal(cells, c2) and np.array_equal(z.view(np.uint32), z2.view(np.uint32)), (seed, W, H)
                    m1, _, _ = oracle.render(xyz, rgb, s0, W, H, rot, image=image, mode=1)
                    assert np.array_equal(cells, m1)
                    stamped += int((cells == 10).sum())
    assert stamped > 50


def test_back_face_proof_never_discards_a_covering_triangle():
    """DESIGN.md section 3, `backface_proven` (csrc/kernels.cuh): a triangle whose computed orientation satisfies
    A_c < -2^-18 * L * D is dropped by the GPU path without looking at any candidate.  Brute force over the
    reference's whole scan domain in strict float32 (numpy restates rasterizer.rs:30-36,58-75): whenever the criterion
    fires, no candidate may pass the three edge tests.  The sample leans on the dangerous cases: slivers and
    nearly collinear triangles of either orientation, tiny and huge coordinates."""
    F = np.float32
    rng = np.random.default_rng(5)
    fired = covered_total = fired_tiny_margin = 0
    for it in range(6000):
        W, H = [(64, 48), (33, 57), (200, 16)][it % 3]
        scale = F([1.0, 1.0, 30.0, 1e3, 1e-2][it % 5])
        a = rng.uniform(-20, 80, 2)
        b = a + rng.uniform(-40, 40, 2)
        t = rng.uniform(-0.5, 1.5)
        off = rng.normal() * 10.0 ** rng.uniform(-7, 1)                 # distance of the third vertex from the line a-b
        if it % 4 == 0:
            off = rng.uniform(-30, 30)                                  # ordinary triangles as well
        d = b - a
        nrm = np.array([-d[1], d[0]]) / (np.hypot(*d) + 1e-12)
        c = a + t * d + off * nrm
        v = (np.stack([a, b, c]) * scale).astype(F)
        if it % 2:
            v = v[[0, 2, 1]]
        (x1, y1), (x2, y2), (x3, y3) = v
        wm1, hm1 = F(W - 1), F(H - 1)
        with np.errstate(all="ignore"):
            mn0, mx0 = min(x1, x2, x3), max(x1, x2, x3)
            mn1, mx1 = min(y1, y2, y3), max(y1, y2, y3)
            # the criterion, operation by operation as in backface_proven
            dx1, dy1, dx2, dy2 = F(x1 - x3), F(y1 - y3), F(x2 - x1), F(y2 - y1)
            area = F(F(dy2 * dx1) - F(dx2 * dy1))
            L = max(F(mx0 - mn0), F(mx1 - mn1))
            D = max(F(max(abs(mn0), abs(mx0)) + wm1), F(max(abs(mn1), abs(mx1)) + hm1))
            T = F(F(L * D) * F(3.814697265625e-06))
            proven = bool(T > F(1e-30) and area < -T)
            # the reference's scan domain and coverage test
            minx = int(np.ceil(max(mn0, F(1.0)))); miny = int(np.ceil(max(mn1, F(1.0))))
            maxx = int(np.ceil(min(F(mx0 * F(2.0)), wm1))); maxy = int(np.ceil(min(mx1, hm1)))
            covered = 0
            if maxx > minx and maxy > miny:
                px = np.arange(minx, maxx, dtype=np.int64).astype(F)[None, :]
                py = np.arange(miny, maxy, dtype=np.int64).astype(F)[:, None]

                def orient(ax, ay, bx, by):
                    return (F(bx - ax) * (py - ay).astype(F)).astype(F) - (F(by - ay) * (px - ax).astype(F)).astype(F)
                w0, w1, w2 = orient(x2, y2, x3, y3), orient(x3, y3, x1, y1), orient(x1, y1, x2, y2)
                covered = int(((w0.astype(F) >= 0) & (w1.astype(F) >= 0) & (w2.astype(F) >= 0)).sum())
        covered_total += covered > 0
        if proven:
            fired += 1
            fired_tiny_margin += bool(area > F(-4.0) * T)
            assert covered == 0, (v.tolist(), float(area), float(T), covered)
    # the sample must exercise both outcomes, including criterion hits close to the threshold
    assert fired > 1000 and covered_total > 100 and fired_tiny_margin > 20, (fired, covered_total, fired_tiny_margin)


def test_band_chunk_cull_bound_is_conservative():
    """Row bands skip a chunk of 32 triangles when its bounding sphere cannot reach the band (k_geom3 `culled`,
    k_chunk_bounds, build_params in csrc/).  Restated in strict float32 / float64 exactly as the kernels and the host
    compute it, and checked against the per-triangle band condition of phase A on random chunks whose bands are placed
    right at the edge of the chunk's row range: a culled chunk must not contain a triangle the band would keep."""
    F = np.float32
    rng = np.random.default_rng(9)
    culled_n = kept_n = near_miss = 0
    for it in range(3000):
        W, H = [(640, 400), (333, 250), (7680, 4320), (80, 40)][it % 4]
        size = 10.0 ** rng.uniform(-3, 4)
        centre = rng.uniform(-1, 1, 3) * size
        spread = size * 10.0 ** rng.uniform(-3, 0)
        tri = (centre + rng.uniform(-1, 1, (32, 3, 3)) * spread).astype(F)
        absmax = F(max(np.abs(tri).max(), size))                      # the scene is at least as large as the chunk
        s0 = F(size)
        rot = oracle.rotation(*rng.uniform(-3.2, 3.2, 3))
        M = oracle.mat4_mul(oracle.utransform(W, H, s0), rot).reshape(4, 4).T   # (row, col)
        m4, m5, m6, m7 = (F(M[1, k]) for k in range(4))

        def yprime(p):   # xform_row: ((m0*x + m1*y) + m2*z) + m3, every operation rounded
            return F(F(F(F(m4 * p[..., 0]) + F(m5 * p[..., 1])) + F(m6 * p[..., 2])) + m7)
        with np.errstate(all="ignore"):
            y = yprime(tri).astype(F)                                   # (32, 3)
            mn1, mx1 = y.min(axis=1), y.max(axis=1)
            miny = np.ceil(np.maximum(mn1, F(1.0))).astype(np.int64)
            maxy = np.ceil(np.minimum(mx1, F(H - 1))).astype(np.int64)
            # k_chunk_bounds
            lo, hi = tri.reshape(-1, 3).min(axis=0), tri.reshape(-1, 3).max(axis=0)
            c = (F(0.5) * lo + F(0.5) * hi).astype(F)
            dlt = (tri.reshape(-1, 3) - c).astype(F)
            r2 = F(((dlt[:, 0] * dlt[:, 0]).astype(F) + (dlt[:, 1] * dlt[:, 1]).astype(F)).astype(F) + (dlt[:, 2] * dlt[:, 2]).astype(F)).max()
            r = F(F(np.sqrt(F(r2))) * F(1.0001) + F(1e-30))
            # build_params (double on the host)
            norm = float(np.sqrt(float(m4) ** 2 + float(m5) ** 2 + float(m6) ** 2)) * 1.0001
            mag = (abs(float(m4)) + abs(float(m5)) + abs(float(m6))) * float(absmax) + abs(float(m7))
            pad = mag * 2.0 ** -19 + 1.0e-3
            if not (np.isfinite(norm) and np.isfinite(pad) and pad < 1.0e6):
                continue
            cull_scale, cull_pad = F(norm), F(pad * 1.0001)
            yc = yprime(c)
            R = F(F(r * cull_scale) + cull_pad)
            # bands right at the edge of the chunk's rows, above and below, with and without the odd-width halo row
            top, bot = int(miny.min()), int(maxy.max())
            for row0, row1 in [(bot + k, bot + k + 7) for k in (-1, 0, 1, 2, 3)] + [(max(0, top - k - 7), max(1, top - k)) for k in (-1, 0, 1, 2, 3)]:
                if row1 <= row0:
                    continue
                for krow0 in {row0, max(0, row0 - 1)}:
                    culled = bool(F(yc - R) >= F(row1) or F(F(yc + R) + F(2.0)) <= F(krow0))
                    keeps = (miny < maxy) & (miny < row1) & (maxy + 1 > krow0)     # has_rows of the BAND variants
                    if culled:
                        culled_n += 1
                        assert not keeps.any(), (it, row0, row1, krow0, float(yc), float(R), miny[keeps][:4], maxy[keeps][:4])
                    else:
                        kept_n += 1
                        near_miss += not keeps.any()
    assert culled_n > 3000 and kept_n > 3000, (culled_n, kept_n)
