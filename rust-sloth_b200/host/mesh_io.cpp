// mesh_io.cpp -- see mesh_io.hpp for the reference lines this follows.
#include "mesh_io.hpp"

#include <algorithm>
#include <cctype>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

namespace sloth {
namespace {

std::vector<std::string> split_ws(const std::string& line)
{
    std::vector<std::string> w;
    size_t i = 0, n = line.size();
    while (i < n) {
        while (i < n && std::isspace((unsigned char)line[i])) ++i;
        size_t j = i;
        while (j < n && !std::isspace((unsigned char)line[j])) ++j;
        if (j > i) w.emplace_back(line, i, j - i);
        i = j;
    }
    return w;
}

std::string trim(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

// Rust's f32::from_str is correctly rounded; so is glibc strtof.  The whole
// token must be consumed.
bool parse_f32(const std::string& tok, float& out)
{
    if (tok.empty()) return false;
    char* end = nullptr;
    out = std::strtof(tok.c_str(), &end);
    return end && *end == '\0';
}

// tobj parse_floatn: take up to n tokens, push each parsed float, succeed only
// if exactly n were pushed.
bool parse_floatn(const std::vector<std::string>& w, size_t& pos, std::vector<float>& vals, size_t n)
{
    size_t got = 0;
    for (; got < n && pos < w.size(); ++got, ++pos) {
        float f;
        if (!parse_f32(w[pos], f)) return false;
        vals.push_back(f);
    }
    return got == n;
}

void dirname_of(const std::string& path, std::string& dir)
{
    size_t p = path.find_last_of('/');
    dir = (p == std::string::npos) ? std::string() : path.substr(0, p + 1);
}

using Material = MtlMaterial;

bool load_mtl_entries(const std::string& path, std::vector<Material>& loaded, std::string& err)
{
    std::ifstream in(path);
    if (!in) {
        err = "open failed: " + path;
        return false;
    }
    std::string line;
    bool have = false;
    Material cur;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        auto w = split_ws(line);
        if (w.empty() || w[0][0] == '#') continue;
        if (w[0] == "newmtl") {
            if (have) loaded.push_back(cur);
            cur = Material();
            cur.name = trim(line.substr(line.find("newmtl") + 6));
            have = true;
        } else if (w[0] == "Kd") {
            std::vector<float> v;
            size_t pos = 1;
            if (!parse_floatn(w, pos, v, 3)) {
                err = "Kd parse error in " + path;
                return false;
            }
            for (int i = 0; i < 3; ++i) cur.diffuse[i] = v[i];
        }
    }
    if (have) loaded.push_back(cur);
    return true;
}

bool load_mtl(const std::string& path, std::vector<Material>& mats, std::map<std::string, size_t>& mat_map,
              std::string& err)
{
    const size_t offset = mats.size();
    std::vector<Material> loaded;
    if (!load_mtl_entries(path, loaded, err)) return false;
    for (size_t i = 0; i < loaded.size(); ++i) {
        mat_map[loaded[i].name] = offset + i;
        mats.push_back(loaded[i]);
    }
    return true;
}

// One corner of a face: position index only matters for the soup.
bool parse_corner(const std::string& tok, size_t n_pos, long& vi)
{
    size_t slash = tok.find('/');
    std::string first = tok.substr(0, slash);
    if (first.empty()) return false;
    char* end = nullptr;
    long x = std::strtol(first.c_str(), &end, 10);
    if (!end || *end != '\0') return false;
    vi = (x < 0) ? (long)n_pos + x : x - 1;
    return true;
}

struct ObjState {
    std::vector<float> pos, vcol;
    std::vector<std::vector<long>> faces;  // current model's faces (corner position indices)
};

// tobj export_faces + geometry.rs:83-142 in one step: triangulate the current
// faces as fans and emit the soup with the reference's colour rules.
bool export_model(const ObjState& st, const std::vector<Material>& mats, bool have_mat, size_t mat_id,
                  SimpleMesh& mesh, std::string& err)
{
    for (int i = 0; i < 3; ++i) mesh.bbox_min[i] = mesh.bbox_max[i] = 0.0f;  // geometry.rs:85-88
    const size_t n_pos = st.pos.size() / 3;
    const bool has_vcol = !st.vcol.empty();
    uint8_t base[3] = {1, 1, 1};  // geometry.rs:91
    if (!mats.empty() && have_mat)
        for (int i = 0; i < 3; ++i) base[i] = f32_as_u8(mats[mat_id].diffuse[i] * 255.0f);
    auto emit = [&](long a, long b, long c) -> bool {
        // material_id.unwrap() sits inside the per-triangle loop (geometry.rs:109-110): a model without faces --
        // the trailing one tobj always pushes, an `o name` with nothing after it -- never reaches it
        if (!mats.empty() && !have_mat) {
            err = "model has no material although the material list is non-empty "
                  "(the reference panics on material_id.unwrap(), geometry.rs:110)";
            return false;
        }
        const long idx[3] = {a, b, c};
        for (int k = 0; k < 3; ++k) {
            if (idx[k] < 0 || (size_t)idx[k] >= n_pos) {
                err = "face references a missing vertex";
                return false;
            }
            for (int d = 0; d < 3; ++d) {
                float v = st.pos[(size_t)idx[k] * 3 + d];
                mesh.xyz.push_back(v);
                // fold of Triangle::aabb into the mesh bbox, geometry.rs:129-135
                mesh.bbox_min[d] = std::fmin(v, mesh.bbox_min[d]);
                mesh.bbox_max[d] = std::fmax(v, mesh.bbox_max[d]);
            }
        }
        uint8_t col[3] = {base[0], base[1], base[2]};
        if (!mats.empty() && has_vcol) {  // geometry.rs:117-126: first corner's vertex colour
            size_t ci = (size_t)a * 3;
            if (ci + 2 >= st.vcol.size()) {
                err = "vertex colour index out of range";
                return false;
            }
            for (int d = 0; d < 3; ++d) col[d] = f32_as_u8(st.vcol[ci + d] * 255.0f);
        }
        mesh.rgb.insert(mesh.rgb.end(), col, col + 3);
        return true;
    };
    for (const auto& f : st.faces) {
        if (f.size() < 3) continue;  // GPU_LOAD_OPTIONS ignores points and lines
        for (size_t i = 2; i < f.size(); ++i)
            if (!emit(f[0], f[i - 1], f[i])) return false;
    }
    return true;
}

}  // namespace

uint8_t f32_as_u8(float v)
{
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

bool load_mtl_file(const std::string& path, std::vector<MtlMaterial>& out, std::string& err)
{
    std::vector<MtlMaterial> loaded;
    if (!load_mtl_entries(path, loaded, err)) return false;
    out.insert(out.end(), loaded.begin(), loaded.end());
    return true;
}

bool load_obj(const std::string& path, std::vector<SimpleMesh>& out, std::string& err)
{
    std::ifstream in(path);
    if (!in) {
        err = "open failed";
        return false;
    }
    std::string dir;
    dirname_of(path, dir);

    ObjState st;
    std::vector<Material> mats;
    std::map<std::string, size_t> mat_map;
    bool have_mat = false;
    size_t mat_id = 0;
    bool mtl_failed = false;
    std::string mtl_err;

    struct Pending {
        std::vector<std::vector<long>> faces;
        bool have_mat;
        size_t mat_id;
    };
    std::vector<Pending> models;  // exported after the whole file is read: vertex colours
                                  // and positions are file-global, materials may load late
    auto flush_model = [&]() {
        models.push_back({st.faces, have_mat, mat_id});
        st.faces.clear();
    };

    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        auto w = split_ws(line);
        if (w.empty() || w[0] == "#") continue;
        const std::string& key = w[0];
        if (key == "v") {
            size_t pos = 1;
            if (!parse_floatn(w, pos, st.pos, 3)) {
                err = "position parse error";
                return false;
            }
            // optional per-vertex colour; tobj ignores the result and keeps whatever parsed
            parse_floatn(w, pos, st.vcol, 3);
        } else if (key == "f" || key == "l") {
            std::vector<long> face;
            for (size_t i = 1; i < w.size(); ++i) {
                long vi;
                if (!parse_corner(w[i], st.pos.size() / 3, vi)) {
                    err = "face parse error";
                    return false;
                }
                face.push_back(vi);
            }
            if (face.empty()) {
                err = "face parse error";
                return false;
            }
            st.faces.push_back(std::move(face));
        } else if (key == "o" || key == "g") {
            if (!st.faces.empty()) flush_model();
        } else if (key == "mtllib") {
            if (w.size() < 2) {
                err = "material parse error";
                return false;
            }
            std::string e;
            if (!load_mtl(dir + w[1], mats, mat_map, e)) {
                mtl_failed = true;
                mtl_err = e;
            }
        } else if (key == "usemtl") {
            std::string name = trim(line.substr(line.find("usemtl") + 6));
            if (name.empty()) {
                err = "material parse error";
                return false;
            }
            auto it = mat_map.find(name);
            bool new_have = it != mat_map.end();
            size_t new_id = new_have ? it->second : 0;
            if ((new_have != have_mat || (new_have && new_id != mat_id)) && !st.faces.empty()) flush_model();
            have_mat = new_have;
            mat_id = new_id;
        }
    }
    flush_model();  // tobj always pushes the trailing model
    if (mtl_failed && mats.empty()) {
        // inputs.rs:112: present.1.expect("Expected to have materials.") -- tobj 3.2.2 returns Ok(materials) whenever the
        // final material list is non-empty, whatever an earlier mtllib statement did
        err = "Expected to have materials. (" + mtl_err + ")";
        return false;
    }
    for (auto& m : models) {
        ObjState view;
        view.pos.swap(st.pos);
        view.vcol.swap(st.vcol);
        view.faces.swap(m.faces);
        SimpleMesh mesh;
        bool ok = export_model(view, mats, m.have_mat, m.mat_id, mesh, err);
        st.pos.swap(view.pos);
        st.vcol.swap(view.vcol);
        if (!ok) return false;
        out.push_back(std::move(mesh));
    }
    return true;
}

bool stl_probe_ascii(const unsigned char* head, size_t n)
{
    size_t end = 0;
    while (end < n && head[end] != '\n') ++end;
    if (end < n) ++end;   // read_line includes the newline
    // UTF-8 validation of the first line (BufRead::read_line fails on invalid UTF-8 and stl_io then reads binary)
    for (size_t i = 0; i < end;) {
        const unsigned char c = head[i];
        size_t len = 0;
        uint32_t cp = 0;
        if (c < 0x80) { ++i; continue; }
        else if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1Fu; }
        else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0Fu; }
        else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07u; }
        else return false;
        if (i + len > end) return false;
        for (size_t k = 1; k < len; ++k) {
            if ((head[i + k] & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (head[i + k] & 0x3Fu);
        }
        if ((len == 2 && cp < 0x80) || (len == 3 && cp < 0x800) || (len == 4 && cp < 0x10000) || cp > 0x10FFFF ||
            (cp >= 0xD800 && cp <= 0xDFFF))
            return false;
        i += len;
    }
    return end >= 6 && std::memcmp(head, "solid ", 6) == 0;
}

bool load_stl(const std::string& path, std::vector<SimpleMesh>& out, std::string& err)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) {
        err = "open failed";
        return false;
    }
    std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    SimpleMesh mesh;
    for (int i = 0; i < 3; ++i) {  // geometry.rs:153-156
        mesh.bbox_min[i] = FLT_MAX;
        mesh.bbox_max[i] = -FLT_MAX;
    }
    auto push_vertex = [&](const float v[3]) {
        for (int d = 0; d < 3; ++d) {
            mesh.xyz.push_back(v[d]);
            mesh.bbox_min[d] = std::fmin(v[d], mesh.bbox_min[d]);
            mesh.bbox_max[d] = std::fmax(v[d], mesh.bbox_max[d]);
        }
    };
    // stl_io::create_stl_reader: AsciiStlReader::probe decides (first line valid UTF-8 and starting with "solid ")
    const bool ascii = stl_probe_ascii(reinterpret_cast<const unsigned char*>(data.data()), data.size());
    size_t n_tri = 0;
    if (ascii) {
        std::istringstream ss(data);
        std::string line;
        int nv = 0;
        while (std::getline(ss, line)) {
            auto w = split_ws(line);
            if (w.size() == 4 && w[0] == "vertex") {
                float v[3];
                for (int d = 0; d < 3; ++d)
                    if (!parse_f32(w[1 + d], v[d])) {
                        err = "stl_io couldnt parse STL: bad vertex";
                        return false;
                    }
                push_vertex(v);
                if (++nv == 3) {
                    nv = 0;
                    ++n_tri;
                }
            }
        }
        if (nv != 0) {
            err = "stl_io couldnt parse STL: truncated facet";
            return false;
        }
        // A binary file whose 80-byte header happens to pass the probe: stl_io's ASCII reader then meets bytes
        // that are no facet and fails; it must not load as an empty mesh (blank frame, infinite scale).
        if (n_tri == 0 && data.size() >= 84) {
            uint32_t n = 0;
            std::memcpy(&n, data.data() + 80, 4);
            if (n != 0 && data.size() == 84 + (size_t)n * 50) {
                err = "stl_io couldnt parse STL: ASCII header (\"solid \") on a binary body";
                return false;
            }
        }
    } else {
        if (data.size() < 84) {
            err = "stl_io couldnt parse STL: short binary header";
            return false;
        }
        uint32_t n;
        std::memcpy(&n, data.data() + 80, 4);
        if (data.size() < 84 + (size_t)n * 50) {
            err = "stl_io couldnt parse STL: truncated binary body";
            return false;
        }
        for (uint32_t t = 0; t < n; ++t) {
            const char* rec = data.data() + 84 + (size_t)t * 50 + 12;  // skip the normal
            for (int k = 0; k < 3; ++k) {
                float v[3];
                std::memcpy(v, rec + k * 12, 12);
                push_vertex(v);
            }
        }
        n_tri = n;
    }
    mesh.rgb.resize(n_tri * 3);
    for (size_t t = 0; t < n_tri; ++t) {  // geometry.rs:161-162
        mesh.rgb[t * 3 + 0] = 0xFF;
        mesh.rgb[t * 3 + 1] = 0xFF;
        mesh.rgb[t * 3 + 2] = 0x00;
    }
    out.push_back(std::move(mesh));
    return true;
}

bool match_meshes(const std::string& arg, std::vector<SimpleMesh>& out, std::string& err)
{
    // inputs.rs:97: value.split(' ') -- consecutive spaces yield empty slices,
    // which then fail the extension test exactly like the reference.
    size_t start = 0;
    for (;;) {
        size_t sp = arg.find(' ', start);
        std::string slice = arg.substr(start, sp == std::string::npos ? std::string::npos : sp - start);
        auto fail = [&](const std::string& s, const std::string& e) {
            err = "filename: [" + slice + "] couldn't load, " + s + ". " + e;
            return false;
        };
        // Path::extension(): text after the last '.' of the file name, None if there
        // is no '.', or the name starts with its only '.'.
        size_t slash = slice.find_last_of('/');
        std::string fname = slash == std::string::npos ? slice : slice.substr(slash + 1);
        size_t dot = fname.find_last_of('.');
        if (dot == std::string::npos || dot == 0) return fail("couldn't determine filename extension", "");
        std::string ext = fname.substr(dot + 1);
        std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char c) { return std::tolower(c); });
        std::string e;
        if (ext == "obj") {
            if (!load_obj(slice, out, e)) return fail("tobj couldnt load/parse OBJ", e);
        } else if (ext == "stl") {
            if (!load_stl(slice, out, e))
                return fail(e == "open failed" ? "STL load failed" : "stl_io couldnt parse STL", e);
        } else {
            return fail("unknown filename extension", "");
        }
        if (sp == std::string::npos) break;
        start = sp + 1;
    }
    return true;
}

float scene_scale0(const std::vector<SimpleMesh>& meshes)
{
    float scale = 0.0f;
    for (const auto& m : meshes)
        scale = std::fmax(std::fmax(std::fmax(scale, m.bbox_max[0]), m.bbox_max[1]), m.bbox_max[2]);
    return scale;
}

void flatten(const std::vector<SimpleMesh>& meshes, std::vector<float>& xyz, std::vector<uint8_t>& rgb)
{
    xyz.clear();
    rgb.clear();
    for (const auto& m : meshes) {
        xyz.insert(xyz.end(), m.xyz.begin(), m.xyz.end());
        rgb.insert(rgb.end(), m.rgb.begin(), m.rgb.end());
    }
}

}  // namespace sloth
