// mesh_io.hpp -- host-side model loading for the sloth drop-in (C++17, CPU only).
//
// Mirrors what the reference does at load time, before the raster path:
//   src/inputs.rs:95-129   match_meshes: split the argument on ' ', dispatch on
//                          the lower-cased extension (obj -> tobj, stl -> stl_io)
//   src/geometry.rs:83-142 tobj Mesh  -> SimpleMesh (de-indexed soup, colour rules, bbox from 0)
//   src/geometry.rs:151-189 stl_io IndexedMesh -> SimpleMesh (colour 255,255,0, bbox from +-f32::MAX)
// The OBJ/MTL/STL parsing itself lives in un-vendored crates (tobj 3.2.2,
// stl_io 0.4.2, Cargo.lock); their documented behaviour is restated here:
// triangulate (fan) + single_index, a new model at every o/g or material change,
// f32 coordinates, optional per-vertex colours on `v` lines, Kd -> diffuse.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sloth {

// geometry.rs:78-81 SimpleMesh, as a soup: 9 floats + 3 colour bytes per triangle.
struct SimpleMesh {
    std::vector<float> xyz;    // n*9: v1.xyz v2.xyz v3.xyz (w = 1 implied)
    std::vector<uint8_t> rgb;  // n*3: Triangle.color
    float bbox_min[3];
    float bbox_max[3];
    size_t size() const { return xyz.size() / 9; }
};

// One `newmtl` block of a .mtl file (tobj Material: only the name and Kd are used, geometry.rs:109-115).
struct MtlMaterial {
    std::string name;
    float diffuse[3] = {0.f, 0.f, 0.f};
};
// Appends the materials of one .mtl file in file order.  false + err when the file cannot be read or a Kd is bad.
bool load_mtl_file(const std::string& path, std::vector<MtlMaterial>& out, std::string& err);
// Rust `f32 as u8`: truncate toward zero, saturate, NaN -> 0 (geometry.rs:111-124).
uint8_t f32_as_u8(float v);

// match_meshes(): `arg` is the single CLI value, split on ' '.  On failure
// returns false and sets `err` to the reference's message format
// ("filename: [..] couldn't load, ..").
bool match_meshes(const std::string& arg, std::vector<SimpleMesh>& out, std::string& err);

// stl_io 0.4.2's AsciiStlReader::probe (restated; the crate is not vendored): the stream is ASCII iff its first line
// -- the bytes up to and including the first '\n', or all of them -- is valid UTF-8 and starts with "solid " (with
// the space; no leading whitespace is skipped).  Everything else is read as binary.  `n` = bytes available in
// `head` (the whole file or a prefix that contains the first line).
bool stl_probe_ascii(const unsigned char* head, size_t n);

bool load_obj(const std::string& path, std::vector<SimpleMesh>& out, std::string& err);
bool load_stl(const std::string& path, std::vector<SimpleMesh>& out, std::string& err);

// context.rs:106-113: scale0 = fold(max) over meshes of bbox.max.{x,y,z}, from 0.0.
float scene_scale0(const std::vector<SimpleMesh>& meshes);

// Concatenate meshes in draw order (main.rs:80-83) into one soup.
void flatten(const std::vector<SimpleMesh>& meshes, std::vector<float>& xyz, std::vector<uint8_t>& rgb);

}  // namespace sloth
