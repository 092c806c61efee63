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
