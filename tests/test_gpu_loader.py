"""GPU tests of the device-side model loader (inputs.rs:95-129 + geometry.rs:83-189 on the
GPU).  The checker is the host loader (host/mesh_io.cpp, itself checked on the CPU against the committed soups of
the reference's models and against the crates' rules in test_host_loaders.py): same files, bit-identical soup,
colours and scene scale.  Decimal -> f32 conversion is additionally checked against libc strtof on fuzzed tokens."""
import ctypes
import os
import struct

import numpy as np
import pytest

import oracle
import rust_sloth_b200 as rs
from rust_sloth_b200 import meshes
import scenes as S

pytestmark = pytest.mark.gpu


def write(p, data):
    with open(p, "wb" if isinstance(data, bytes) else "w") as f:
        f.write(data)
    return str(p)


def host_scene(arg):
    ms = rs.match_meshes(arg)
    return (np.concatenate([m.xyz for m in ms]).reshape(-1, 9), np.concatenate([m.rgb for m in ms]).reshape(-1, 3),
            rs.scene_scale0(ms))


def device_scene(arg):
    ctx = rs.Context.blank(True)
    try:
        n, smax = ctx.load_models(arg)
        xyz, rgb, smax2 = ctx.scene()
        assert n == xyz.shape[0] and smax == smax2
        return xyz, rgb, smax
    finally:
        ctx.close()


def assert_same_scene(arg):
    hx, hr, hs = host_scene(arg)
    dx, dr, ds = device_scene(arg)
    assert dx.shape == hx.shape, (dx.shape, hx.shape)
    assert np.array_equal(dx.view(np.uint32), hx.view(np.uint32))
    assert np.array_equal(dr, hr)
    # numeric, not bitwise: for a scene without positive coordinates fmax(0.0, -0.0) is -0.0 in glibc and either
    # zero under IEEE maxNum (Rust's f32::max); the device fold starts from +0.0 and keeps it
    assert np.float32(ds) == np.float32(hs)
    return hx, hr, hs


OBJ_FEATURES = """# every grammar feature the bundled models use, and a few they do not
mtllib m.mtl
o first
v 0 0 0
v 1 0 0
v\t1 1 0
v 0 1 0
v 0.5 1.5 -0.25
v +2.5e-1 -1.25E+1 .5
v 1. 007 -0
vn 0 0 1
vt 0.5 0.5
usemtl red
f 1 2 3 4 5
s off
usemtl grey
f -5//1 -4//1 -3//1
l 1 2
f 1 2
g second
f 1/1/1 2/2/2 3/3/3 4/4/4
usemtl   name with spaces
f 7 6 5 4 3 2 1
usemtl red
f 3/1 2/1 1/1
"""
MTL_FEATURES = "newmtl red\nKd 1.0 0.5 0.003\nnewmtl grey\nKd 0.64 0.64 0.64\nnewmtl name with spaces\nKd 0 2 -1\n"


def test_obj_grammar_matches_host_loader(tmp_path):
    write(tmp_path / "m.mtl", MTL_FEATURES)
    hx, hr, _ = assert_same_scene(write(tmp_path / "a.obj", OBJ_FEATURES))
    assert hx.shape[0] == 3 + 1 + 2 + 5 + 1
    assert tuple(hr[0]) == (255, 127, 0) and tuple(hr[6]) == (0, 255, 0)
    # CRLF line endings and no trailing newline
    assert_same_scene(write(tmp_path / "crlf.obj", OBJ_FEATURES.replace("\n", "\r\n").rstrip()))


def test_vertex_colours_and_missing_mtllib(tmp_path):
    write(tmp_path / "c.mtl", "newmtl k\nKd 0.2 0.2 0.2\n")
    body = "v 0 0 0 1 0 0.5\nv 1 0 0 0 1 0\nv 0 1 0 0 0 1\nv 2 2 2 0.25 0.75 1.5\nf 2 3 1\nf 1 2 3\nf 4 1 2 3\n"
    hx, hr, _ = assert_same_scene(write(tmp_path / "c.obj", "mtllib c.mtl\nusemtl k\n" + body))
    assert tuple(hr[0]) == (0, 255, 0) and tuple(hr[1]) == (255, 0, 127) and tuple(hr[2]) == (63, 191, 255)
    _, hr, _ = assert_same_scene(write(tmp_path / "b.obj", body))            # no materials: colour (1,1,1)
    assert (hr == 1).all()
    # colours on some vertices only: the flat colour array is indexed by position index, as in the reference
    ragged = "v 0 0 0\nv 1 0 0 0 1 0\nv 0 1 0 0 0 1\nv 3 3 3 1 1\nf 1 2 3\n"
    assert_same_scene(write(tmp_path / "r.obj", "mtllib c.mtl\nusemtl k\n" + ragged))


def test_usemtl_sees_only_materials_loaded_so_far(tmp_path):
    write(tmp_path / "one.mtl", "newmtl a\nKd 1 0 0\n")
    write(tmp_path / "two.mtl", "newmtl a\nKd 0 1 0\nnewmtl b\nKd 0 0 1\n")
    text = "v 0 0 0\nv 1 0 0\nv 0 1 0\nmtllib one.mtl\nusemtl a\nf 1 2 3\nmtllib two.mtl\nusemtl a\nf 1 2 3\nusemtl b\nf 3 2 1\n"
    _, hr, _ = assert_same_scene(write(tmp_path / "late.obj", text))
    assert [tuple(c) for c in hr] == [(255, 0, 0), (0, 255, 0), (0, 0, 255)]


def strtof_array(tokens):
    libc = ctypes.CDLL("libc.so.6")
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    return np.array([libc.strtof(t.encode(), None) for t in tokens], np.float32)


def fuzz_tokens(rng, n):
    toks = []
    f32 = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    for i in range(n):
        kind = i % 6
        v = float(f32[i])
        if kind == 0 and np.isfinite(v) and 1e-15 < abs(v) < 1e15:
            toks.append("%.9g" % v)                                  # shortest round-trip style
        elif kind == 1:
            toks.append("%.6f" % rng.uniform(-1000, 1000))           # what Blender writes
        elif kind == 2:
            toks.append("%.17g" % rng.uniform(-10, 10))              # repr of a double: 17 significant digits
        elif kind == 3:                                              # random digit strings, up to 19 digits
            nd = int(rng.integers(1, 20))
            digits = "".join(str(d) for d in rng.integers(0, 10, nd))
            frac = int(rng.integers(0, nd + 1))
            ex = int(rng.integers(-27 + frac, 28 - 19 + frac))       # keeps the decimal exponent inside [-27, 27]
            body = digits[:nd - frac] + ("." + digits[nd - frac:] if frac or rng.integers(2) else "")
            if body.startswith("."):
                body = ("0" if rng.integers(2) else "") + body
            toks.append(("-" if rng.integers(2) else "") + body + ("e%d" % ex if ex or rng.integers(2) else ""))
        elif kind == 4:                                              # integers around 2^24 .. 2^26: exact ties
            toks.append(str(int(rng.integers(2**24 - 64, 2**26))))
        else:                                                        # halfway points with short expansions
            k = int(rng.integers(1, 2**23))
            toks.append("%d.%s" % (k, "5" if rng.integers(2) else "50000000001"))
    return toks


def test_decimal_to_f32_is_correctly_rounded(tmp_path):
    rng = np.random.default_rng(7)
    toks = fuzz_tokens(rng, 30000 * 3)
    lines = []
    for i in range(0, len(toks), 3):
        lines.append("v %s %s %s" % (toks[i], toks[i + 1], toks[i + 2]))
    n_v = len(lines)
    # one triangle per three vertices, by negative indices
    text = "\n".join(" \n".join(lines[i:i + 3]) + "\nf -3 -2 -1" for i in range(0, n_v - n_v % 3, 3)) + "\n"
    dx, _, _ = device_scene(write(tmp_path / "fuzz.obj", text))
    want = strtof_array(toks[: (n_v - n_v % 3) * 3])
    got = dx.reshape(-1)
    bad = np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))
    assert bad.size == 0, [(toks[i], got[i], want[i]) for i in bad[:8]]


def test_tokens_the_device_does_not_decide_are_reported(tmp_path):
    tri = "\nv 1 0 0\nv 0 1 0\nf 1 2 3\n"
    # 1 + 2^-24 written out in full: more than 19 digits and exactly on a rounding boundary
    for tok in ["1.000000059604644775390625", "nan", "inf", "1e40", "1e-40"]:
        ctx = rs.Context.blank(True)
        with pytest.raises(rs.SlothError) as e:
            ctx.load_models(write(tmp_path / "u.obj", "v %s 0 0" % tok + tri))
        assert e.value.code == rs.SLOTH_E_UNSUPPORTED, tok
        assert "line 1" in str(e.value)
        ctx.close()
    # long expansions away from a boundary are decided (and agree with strtof)
    toks = ["0.1000000000000000055511151231257827", "3.14159265358979323846264338327950288", "123456789012345678901234567890e-25"]
    dx, _, _ = device_scene(write(tmp_path / "long.obj", "v %s %s %s" % tuple(toks) + tri))
    assert np.array_equal(dx[0, :3].view(np.uint32), strtof_array(toks).view(np.uint32))


def soup_to_obj(xyz, rgb, mtl_name):
    """Every triangle as three v lines + one face; colours through one material per distinct colour."""
    cols, inv = np.unique(rgb, axis=0, return_inverse=True)
    mtl = "".join("newmtl c%d\nKd %.9g %.9g %.9g\n" % (i, *(np.float32(c) / np.float32(255.0) + np.float32(0.001)))
                  for i, c in enumerate(cols))
    out = ["mtllib %s" % mtl_name]
    last = -1
    for t in range(xyz.shape[0]):
        if inv[t] != last:
            last = int(inv[t])
            out.append("usemtl c%d" % last)
        v = xyz[t]
        out.append("v %.9g %.9g %.9g\nv %.9g %.9g %.9g\nv %.9g %.9g %.9g\nf -3 -2 -1" % tuple(float(x) for x in v))
    return "\n".join(out) + "\n", mtl


@pytest.mark.parametrize("scene", ["cube", "suzy", "pikachu", "skull", "vaporeon"])
def test_committed_soups_survive_a_round_trip_through_obj_text(scene, tmp_path):
    xyz, rgb, s0 = S.soup(scene)
    text, mtl = soup_to_obj(xyz, rgb, "s.mtl")
    write(tmp_path / "s.mtl", mtl)
    dx, dr, ds = device_scene(write(tmp_path / "s.obj", text))
    assert np.array_equal(dx.view(np.uint32), xyz.view(np.uint32))
    assert np.array_equal(dr, rgb)
    assert np.float32(ds) == np.float32(max(0.0, float(xyz.max())))


def stl_ascii(xyz):
    out = ["solid s"]
    for v in xyz:
        out.append(" facet normal 0 0 0\n  outer loop")
        for k in range(3):
            out.append("   vertex %.9g %.9g %.9g" % tuple(float(x) for x in v[3 * k:3 * k + 3]))
        out.append("  endloop\n endfacet")
    out.append("endsolid s")
    return "\n".join(out) + "\n"


def stl_binary(xyz):
    body = b"".join(struct.pack("<12fH", 0, 0, 0, *[float(x) for x in v], 0) for v in xyz)
    return b"binary stl".ljust(80, b"\0") + struct.pack("<I", xyz.shape[0]) + body


def test_stl_ascii_and_binary(tmp_path):
    xyz, _, _ = S.soup("part_stl")
    for name, data in [("a.stl", stl_ascii(xyz)), ("b.STL", stl_binary(xyz)), ("neg.stl", stl_binary(-np.abs(xyz)))]:
        hx, hr, hs = assert_same_scene(write(tmp_path / name, data))
        assert (hr == [255, 255, 0]).all()
    assert np.array_equal(hx, -np.abs(xyz)) and hs == 0.0                  # all-negative scene: scale stays 0
    hx, _, _ = assert_same_scene(str(tmp_path / "a.stl"))
    assert np.array_equal(hx.view(np.uint32), xyz.view(np.uint32))
    assert_same_scene(write(tmp_path / "empty.stl", stl_binary(xyz[:0])))


def test_scene_load_of_several_files_renders_like_the_host_path(tmp_path):
    xyz, rgb, _ = S.soup("suzy")
    text, mtl = soup_to_obj(xyz, rgb, "s.mtl")
    write(tmp_path / "s.mtl", mtl)
    write(tmp_path / "s.obj", text)
    write(tmp_path / "p.stl", stl_binary(S.soup("cube_stl")[0]))
    arg = "%s %s %s" % (tmp_path / "s.obj", tmp_path / "p.stl", tmp_path / "s.obj")
    hx, hr, hs = assert_same_scene(arg)
    assert hx.shape[0] == 2 * xyz.shape[0] + 12
    rot = oracle.rotation(0.0, S.PI + 0.4, 0.0)
    ocells, _, _ = oracle.render(hx, hr, hs, 160, 80, rot, image=True, mode=0)
    ctx = rs.Context.blank(True)
    try:
        ctx.load_models(arg)
        ctx.resize(160, 80)
        cells, _ = ctx.render(rot)
        assert np.array_equal(cells, ocells)
        ctx.load_models(str(tmp_path / "p.stl"))                                   # replacing the scene works
        assert ctx.scene()[0].shape[0] == 12
    finally:
        ctx.close()


def test_non_finite_and_huge_coordinates_switch_off_the_regular_shortcut(tmp_path):
    # binary STL can carry any bit pattern: the device-side scan must flag the scene like sloth_scene_set does
    xyz, rgb, s0 = meshes.random_soup(3, 40, kind="uniform")
    xyz = xyz.copy()
    xyz[5, 2] = np.float32("nan"); xyz[9, 0] = np.float32("-inf"); xyz[11, 4] = np.float32(-3e30)
    path = write(tmp_path / "wild.stl", stl_binary(xyz))
    hx, hr, hs = host_scene(path)
    rot = oracle.rotation(0.2, S.PI + 0.3, 0.1)
    ocells, _, _ = oracle.render(hx, hr, hs, 64, 48, rot, image=True, mode=0)
    ctx = rs.Context.blank(True)
    try:
        _, smax = ctx.load_models(path)
        assert np.float32(smax).view(np.uint32) == np.float32(hs).view(np.uint32)
        ctx.resize(64, 48)
        cells, _ = ctx.render(rot)
        assert np.array_equal(cells, ocells)
    finally:
        ctx.close()


@pytest.mark.parametrize("text,code,frag", [
    ("v 0 0\nf 1 1 1\n", rs.SLOTH_E_PARSE, "position parse error"),
    ("v 0 0 zero\n", rs.SLOTH_E_PARSE, "position parse error"),
    ("v 0 0 0\nf 1 x 1\n", rs.SLOTH_E_PARSE, "face parse error"),
    ("v 0 0 0\nf\n", rs.SLOTH_E_PARSE, "face parse error"),
    ("v 0 0 0\nf 1 2 3\n", rs.SLOTH_E_PARSE, "face references a missing vertex"),
    ("v 0 0 0\nf 1 1 0\n", rs.SLOTH_E_PARSE, "face references a missing vertex"),
    ("usemtl\n", rs.SLOTH_E_PARSE, "material parse error"),
    ("mtllib\n", rs.SLOTH_E_PARSE, "material parse error"),
    ("mtllib nowhere.mtl\nv 0 0 0\nf 1 1 1\n", rs.SLOTH_E_IO, "Expected to have materials."),
    ("mtllib ok.mtl\nv 0 0 0\nf 1 1 1\n", rs.SLOTH_E_PARSE, "model has no material"),
    # an unknown material only matters once a face uses it (material_id.unwrap() sits in the per-triangle loop)
    ("mtllib ok.mtl\nusemtl k\nv 0 0 0\nf 1 1 1\nusemtl other\nf 1 1 1\n", rs.SLOTH_E_PARSE, "model has no material"),
])
def test_malformed_obj_is_rejected_with_the_host_loaders_message(text, code, frag, tmp_path):
    write(tmp_path / "ok.mtl", "newmtl k\nKd 1 1 1\n")
    path = write(tmp_path / "bad.obj", text)
    with pytest.raises(rs.SlothError) as host:
        rs.match_meshes(path)
    assert frag in str(host.value)
    ctx = rs.Context.blank(True)
    try:
        with pytest.raises(rs.SlothError) as dev:
            ctx.load_models(path)
        assert dev.value.code == code
        assert frag in str(dev.value) and "tobj couldnt load/parse OBJ" in str(dev.value)
    finally:
        ctx.close()


def test_file_level_errors(tmp_path):
    ctx = rs.Context.blank(True)
    try:
        for arg, code, frag in [("nothing", rs.SLOTH_E_ARG, "couldn't determine filename extension"),
                                ("x.ply", rs.SLOTH_E_ARG, "unknown filename extension"),
                                (str(tmp_path / "missing.obj"), rs.SLOTH_E_IO, "tobj couldnt load/parse OBJ"),
                                (str(tmp_path / "missing.stl"), rs.SLOTH_E_IO, "STL load failed"),
                                (" a.obj", rs.SLOTH_E_ARG, "filename: []")]:
            with pytest.raises(rs.SlothError) as e:
                ctx.load_models(arg)
            assert e.value.code == code and frag in str(e.value), (arg, str(e.value))
        with pytest.raises(rs.SlothError) as e:
            ctx.load_models(write(tmp_path / "t.stl", "solid x\n vertex 0 0 0\n vertex 1 0 0\nendsolid x\n"))
        assert "truncated facet" in str(e.value)
        with pytest.raises(rs.SlothError) as e:
            ctx.load_models(write(tmp_path / "s.stl", b"\1" * 90))
        assert "stl_io couldnt parse STL" in str(e.value)
    finally:
        ctx.close()


def test_large_obj_matches_host_loader(tmp_path):
    """~330 k triangles of indexed OBJ (shared vertices, positive indices): every scan crosses many blocks."""
    xyz, rgb, s0 = meshes.icosphere(128)
    verts, inv = np.unique(xyz.reshape(-1, 3), axis=0, return_inverse=True)
    vt = np.char.mod("%.9g", verts)
    text = "\n".join("v " + " ".join(r) for r in vt) + "\n"
    faces = (inv.reshape(-1, 3) + 1).astype(str)
    text += "\n".join("f " + " ".join(r) for r in faces) + "\n"
    path = write(tmp_path / "ico.obj", text)
    hx, hr, hs = assert_same_scene(path)
    assert np.array_equal(hx.view(np.uint32), xyz.view(np.uint32)) and hs == np.float32(xyz.max())


def test_stl_flavour_probe_and_material_rules_on_the_device(tmp_path):
    """The device loader decides ASCII / binary like stl_io's probe (first line valid UTF-8 and starting with "solid ")
    and applies the reference's per-triangle material unwrap: the same crafted files as the host loader's tests."""
    import test_host_loaders as H
    ctx = rs.Context.blank(True)
    try:
        for k, (data, want) in enumerate(H.stl_probe_cases()):
            path = tmp_path / f"p{k}.stl"
            path.write_bytes(data)
            if want is None:
                with pytest.raises(rs.SlothError) as e:
                    ctx.load_models(str(path))
                assert "stl_io couldnt parse STL" in str(e.value), (k, str(e.value))
            else:
                n, _ = ctx.load_models(str(path))
                assert n == want, k
        (tmp_path / "m.mtl").write_text("newmtl red\nKd 1 0 0\n")
        base = "mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl red\nf 1 2 3\n"
        for tail in ("usemtl nosuch\n", "o empty\n"):
            (tmp_path / "a.obj").write_text(base + tail)
            n, _ = ctx.load_models(str(tmp_path / "a.obj"))
            assert n == 1 and tuple(ctx.scene()[1][0]) == (255, 0, 0)
        (tmp_path / "b.obj").write_text(base + "usemtl nosuch\nf 1 2 3\n")
        with pytest.raises(rs.SlothError):
            ctx.load_models(str(tmp_path / "b.obj"))
        (tmp_path / "c.obj").write_text("mtllib missing.mtl\n" + base)
        n, _ = ctx.load_models(str(tmp_path / "c.obj"))
        assert n == 1
    finally:
        ctx.close()
