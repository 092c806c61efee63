"""CPU tests of the host-side loaders (the tobj / stl_io rules of SURVEY.md Appendix C)."""
import os
import struct

import numpy as np
import pytest

import rust_sloth_b200 as rs
import scenes as S

MODELS = "/root/reference/models/"


def write(p, text):
    with open(p, "w") as f:
        f.write(text)


def test_obj_fan_triangulation_materials_and_order(tmp_path):
    write(tmp_path / "m.mtl", "newmtl red\nKd 1.0 0.5 0.003\nnewmtl grey\nKd 0.64 0.64 0.64\n")
    write(tmp_path / "a.obj", """mtllib m.mtl
o first
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0.5 1.5 0
usemtl red
f 1 2 3 4 5
usemtl grey
f -5//1 -4//1 -3//1
o second
f 1/1/1 2/2/2 3/3/3 4/4/4
""")
    meshes = rs.match_meshes(str(tmp_path / "a.obj"))
    assert [len(m) for m in meshes] == [3, 1, 2]                    # pentagon fan, material split, quad
    m0 = meshes[0]
    assert np.array_equal(m0.xyz[0], [0, 0, 0, 1, 0, 0, 1, 1, 0])
    assert np.array_equal(m0.xyz[1], [0, 0, 0, 1, 1, 0, 0, 1, 0])
    assert np.array_equal(m0.xyz[2], [0, 0, 0, 0, 1, 0, 0.5, 1.5, 0])
    assert tuple(m0.rgb[0]) == (255, 127, 0)                        # (Kd*255) as u8: truncation
    assert tuple(meshes[1].rgb[0]) == (163, 163, 163)
    assert np.array_equal(meshes[1].xyz[0], m0.xyz[0])              # negative indices
    assert tuple(meshes[2].rgb[0]) == (163, 163, 163)               # material carries over to the next object
    assert rs.scene_scale0(meshes) == np.float32(1.5)
    assert np.array_equal(meshes[0].bounding_box.min, [0, 0, 0])    # OBJ bbox fold starts at 0


def test_obj_without_mtllib_is_colour_1_1_1_and_vertex_colours_need_materials(tmp_path):
    write(tmp_path / "b.obj", "v 0 0 0 1 0 0\nv 1 0 0 0 1 0\nv 0 1 0 0 0 1\nf 1 2 3\n")
    (m,) = rs.match_meshes(str(tmp_path / "b.obj"))
    assert tuple(m.rgb[0]) == (1, 1, 1)
    write(tmp_path / "c.mtl", "newmtl k\nKd 0.2 0.2 0.2\n")
    write(tmp_path / "c.obj", "mtllib c.mtl\nusemtl k\nv 0 0 0 1 0 0.5\nv 1 0 0 0 1 0\nv 0 1 0 0 0 1\nf 2 3 1\nf 1 2 3\n")
    (m,) = rs.match_meshes(str(tmp_path / "c.obj"))
    assert tuple(m.rgb[0]) == (0, 255, 0) and tuple(m.rgb[1]) == (255, 0, 127)   # first corner's colour


def test_stl_ascii_and_binary(tmp_path):
    write(tmp_path / "t.stl", """solid x
 facet normal 0 0 1
  outer loop
   vertex 0 0 0
   vertex 2 0 0
   vertex 0 3 -1
  endloop
 endfacet
endsolid x
""")
    (m,) = rs.match_meshes(str(tmp_path / "t.stl"))
    assert len(m) == 1 and tuple(m.rgb[0]) == (255, 255, 0)
    assert np.array_equal(m.bounding_box.max, [2, 3, 0]) and np.array_equal(m.bounding_box.min, [0, 0, -1])
    with open(tmp_path / "b.STL", "wb") as f:
        f.write(b"\0" * 80 + struct.pack("<I", 1) + struct.pack("<12fH", 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 3, -1, 0))
    (b,) = rs.match_meshes(str(tmp_path / "b.STL"))                 # extension match is case-insensitive
    assert np.array_equal(b.xyz, m.xyz)


def test_error_messages_follow_the_reference(tmp_path):
    for arg, frag in [("nothing", "couldn't determine filename extension"), ("x.ply", "unknown filename extension"),
                      (str(tmp_path / "missing.obj"), "tobj couldnt load/parse OBJ"), (" a.obj", "filename: []")]:
        with pytest.raises(rs.SlothError) as e:
            rs.match_meshes(arg)
        assert frag in str(e.value)


@pytest.mark.skipif(not os.path.isdir(MODELS), reason="reference checkout not present (GPU box)")
def test_bundled_models_match_the_committed_soups():
    table = {"cube.obj": 12, "ferris.obj": 1004, "suzy.obj": 968, "Pikachu.obj": 2742, "skull.obj": 3185,
             "Vaporeon.obj": 5540, "cube.stl": 12, "part.stl": 276}
    names = {"cube.obj": "cube", "ferris.obj": "ferris", "suzy.obj": "suzy", "Pikachu.obj": "pikachu",
             "skull.obj": "skull", "Vaporeon.obj": "vaporeon", "cube.stl": "cube_stl", "part.stl": "part_stl"}
    for f, n in table.items():
        meshes = rs.match_meshes(MODELS + f)
        assert sum(len(m) for m in meshes) == n
        xyz, rgb, s0 = S.soup(names[f])
        assert np.array_equal(np.concatenate([m.xyz for m in meshes]), xyz)
        assert np.array_equal(np.concatenate([m.rgb for m in meshes]), rgb)
        assert rs.scene_scale0(meshes) == s0
    two = rs.match_meshes(MODELS + "suzy.obj " + MODELS + "suzy.obj")
    assert sum(len(m) for m in two) == 1936


# ---- the device loader's decimal -> f32 conversion, run on the host (csrc/dec_float.cuh) ------------------------
def _strtof(tokens):
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    return np.array([libc.strtof(t.encode(), None) for t in tokens], np.float32)


def _fuzz_tokens(rng, n):
    toks = []
    f32 = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    for i in range(n):
        kind, v = i % 8, float(f32[i])
        if kind == 0 and np.isfinite(v) and 1e-15 < abs(v) < 1e15:
            toks.append("%.9g" % v)                                  # shortest round trip of a random f32
        elif kind == 1:
            toks.append("%.6f" % rng.uniform(-1000, 1000))           # what Blender writes
        elif kind == 2:
            toks.append("%.17g" % rng.uniform(-10, 10))              # repr of a double
        elif kind == 3:                                              # random digit strings, up to 19 digits
            nd = int(rng.integers(1, 20))
            digits = "".join(str(d) for d in rng.integers(0, 10, nd))
            frac = int(rng.integers(0, nd + 1))
            ex = int(rng.integers(-27 + frac, 9 + frac))             # decimal exponent stays inside [-27, 27]
            body = digits[:nd - frac] + ("." + digits[nd - frac:] if frac or rng.integers(2) else "")
            if body.startswith("."):
                body = ("0" if rng.integers(2) else "") + body
            toks.append(("-" if rng.integers(2) else "+" if rng.integers(4) == 0 else "") + body +
                        (("e%d" if rng.integers(2) else "E%+d") % ex if ex or rng.integers(2) else ""))
        elif kind == 4:                                              # integers around 2^24 .. 2^26: exact ties
            toks.append(str(int(rng.integers(2**24 - 64, 2**26))))
        elif kind == 5:                                              # halfway points with short expansions
            k = int(rng.integers(1, 2**23))
            toks.append("%d.%s" % (k, "5" if rng.integers(2) else "50000000001" if rng.integers(2) else "49999999999"))
        elif kind == 6 and np.isfinite(v) and 1e-20 < abs(v) < 1e20:  # exact midpoint of two neighbouring f32, in full
            import fractions
            a = np.float32(abs(v)); b = np.nextafter(a, np.float32(np.inf))
            mid = (fractions.Fraction(float(a)) + fractions.Fraction(float(b))) / 2
            num, den = mid.numerator, mid.denominator                # den is a power of two
            k = den.bit_length() - 1
            toks.append("%d.%s" % (num >> k, str((num & (den - 1)) * 5**k).rjust(k, "0")) if k else str(num))
        else:
            toks.append("%.8e" % rng.uniform(-1e6, 1e6))
    return toks


def test_device_decimal_to_f32_routine_matches_strtof_on_the_host():
    rng = np.random.default_rng(11)
    toks = _fuzz_tokens(rng, 200000)
    got, status = rs.parse_f32_tokens(toks)
    want = _strtof(toks)
    decided = status == 0
    assert not (status == 1).any(), [t for t, s in zip(toks, status) if s == 1][:5]
    bad = np.flatnonzero(decided & (got.view(np.uint32) != want.view(np.uint32)))
    assert bad.size == 0, [(toks[i], got[i], want[i]) for i in bad[:8]]
    # only the long exact-midpoint expansions may be left undecided, and most tokens are decided
    undecided = [t for t, s in zip(toks, status) if s == 2]
    assert all(len(t.replace(".", "").replace("-", "").lstrip("0")) > 19 for t in undecided), undecided[:5]
    assert decided.mean() > 0.9


def test_device_decimal_routine_grammar():
    toks = ["1", "-0", "+.5", "5.", "1e5", "1E-5", "007", "0.000", "-0.0e9", "1e", ".", "-", "+", "1.2.3", "1f", "0x10",
            "abc", "nan", "inf", "-Infinity", "1e40", "1e-40", "340282356779733661637539395458142568448",
            "340282346638528859811704183484516925440", "0.00000000000000000000000000000000000001"]
    got, status = rs.parse_f32_tokens(toks)
    want = _strtof(toks)
    # the 39-digit number at index 22 is FLT_MAX + half an ulp exactly (the overflow boundary): left undecided
    expect = [0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 0, 2]
    assert list(status) == expect, list(zip(toks, status))
    for t, g, w, s in zip(toks, got, want, status):
        if s == 0:
            assert np.float32(g).view(np.uint32) == np.float32(w).view(np.uint32), (t, g, w)
    assert np.signbit(got[1]) and np.signbit(got[8]) and got[23] == np.float32(3.4028235e38)


def _binary_stl(header: bytes, tris) -> bytes:
    body = b"".join(struct.pack("<12fH", 0, 0, 1, *t, 0) for t in tris)
    return header.ljust(80, b"\0")[:80] + struct.pack("<I", len(tris)) + body


STL_TRI = (0.0, 0.0, 0.0, 2.0, 0.0, 0.0, 0.0, 3.0, -1.0)


def stl_probe_cases():
    """(file bytes, expected triangles or None = the reference fails with 'stl_io couldnt parse STL').
    stl_io 0.4.2 probes the FIRST LINE: ASCII iff it is valid UTF-8 and starts with "solid " (restated, unpinned)."""
    return [
        # binary, header starts with "solid" but the first line (up to the first 0x0A) is not valid UTF-8 -> binary reader
        (_binary_stl(b"solid \xff\xfe made by some CAD tool", [STL_TRI]), 1),
        # binary, header "solidworks ..." (no space after "solid") -> binary reader
        (_binary_stl(b"solidworks export", [STL_TRI, STL_TRI]), 2),
        # binary body behind a header that passes the probe (all bytes up to the first newline are ASCII): the
        # reference's ASCII reader fails on it; it must not load as an empty mesh
        (_binary_stl(b"solid ascii-looking header\n", [STL_TRI]), None),
        # leading whitespace before "solid": not ASCII for stl_io -> binary reader -> too short -> error
        (b"  solid x\n facet normal 0 0 1\n  outer loop\n   vertex 0 0 0\n   vertex 2 0 0\n   vertex 0 3 -1\n  endloop\n endfacet\nendsolid x\n", None),
        # plain ASCII with zero facets: an empty mesh is what the reference gets, too
        (b"solid empty\nendsolid empty\n", 0),
    ]


def test_stl_flavour_probe_follows_stl_io(tmp_path):
    for k, (data, want) in enumerate(stl_probe_cases()):
        path = tmp_path / f"p{k}.stl"
        path.write_bytes(data)
        if want is None:
            with pytest.raises(rs.SlothError) as e:
                rs.match_meshes(str(path))
            assert "stl_io couldnt parse STL" in str(e.value), (k, str(e.value))
        else:
            (m,) = rs.match_meshes(str(path))
            assert len(m) == want, k
            if want:
                assert np.array_equal(m.xyz[0], np.array(STL_TRI, np.float32))


def test_material_unwrap_happens_per_triangle(tmp_path):
    """geometry.rs:109-110: material_id.unwrap() sits inside the triangle loop -- a model without faces (the trailing
    one tobj always pushes, `o name` with nothing after it, a final `usemtl <unknown>`) loads; a FACE without a known
    material does not.  tobj returns Ok(materials) when the final list is non-empty, whatever an earlier mtllib did."""
    write(tmp_path / "m.mtl", "newmtl red\nKd 1 0 0\n")
    base = "mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl red\nf 1 2 3\n"
    for tail in ("usemtl nosuch\n", "o empty\n", "usemtl nosuch\no also_empty\n"):
        write(tmp_path / "a.obj", base + tail)
        meshes = rs.match_meshes(str(tmp_path / "a.obj"))
        assert sum(len(m) for m in meshes) == 1 and tuple(meshes[0].rgb[0]) == (255, 0, 0)
    write(tmp_path / "b.obj", base + "usemtl nosuch\nf 1 2 3\n")
    with pytest.raises(rs.SlothError) as e:
        rs.match_meshes(str(tmp_path / "b.obj"))
    assert "no material" in str(e.value)
    write(tmp_path / "c.obj", "mtllib missing.mtl\n" + base)
    meshes = rs.match_meshes(str(tmp_path / "c.obj"))
    assert sum(len(m) for m in meshes) == 1
