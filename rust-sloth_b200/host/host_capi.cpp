// host_capi.cpp -- C entry points over the host-side loaders (CPU only), so the
// Python tests and the fixture script use the same loader as the `sloth` CLI.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mesh_io.hpp"
#include "../csrc/dec_float.cuh"   // the device loader's decimal -> f32 routine, compiled for the host

extern "C" {

struct sloth_host_scene {
    float* xyz;           // n_tri*9
    uint8_t* rgb;         // n_tri*3
    size_t n_tri;
    size_t n_meshes;
    size_t* mesh_sizes;   // triangles per mesh, draw order
    float* mesh_bbox;     // n_meshes*6: min.xyz, max.xyz
    float scale0;         // context.rs:106-113
    char error[512];
};

__attribute__((visibility("default"))) sloth_host_scene* sloth_host_load(const char* arg)
{
    auto* s = (sloth_host_scene*)std::calloc(1, sizeof(sloth_host_scene));
    std::vector<sloth::SimpleMesh> meshes;
    std::string err;
    if (!sloth::match_meshes(arg, meshes, err)) {
        std::strncpy(s->error, err.c_str(), sizeof s->error - 1);
        return s;
    }
    std::vector<float> xyz;
    std::vector<uint8_t> rgb;
    sloth::flatten(meshes, xyz, rgb);
    s->n_tri = xyz.size() / 9;
    s->n_meshes = meshes.size();
    s->xyz = (float*)std::malloc(xyz.size() * sizeof(float) + 4);
    s->rgb = (uint8_t*)std::malloc(rgb.size() + 4);
    std::memcpy(s->xyz, xyz.data(), xyz.size() * sizeof(float));
    std::memcpy(s->rgb, rgb.data(), rgb.size());
    s->mesh_sizes = (size_t*)std::malloc(sizeof(size_t) * (meshes.size() + 1));
    s->mesh_bbox = (float*)std::malloc(sizeof(float) * 6 * (meshes.size() + 1));
    for (size_t i = 0; i < meshes.size(); ++i) {
        s->mesh_sizes[i] = meshes[i].size();
        for (int d = 0; d < 3; ++d) {
            s->mesh_bbox[i * 6 + d] = meshes[i].bbox_min[d];
            s->mesh_bbox[i * 6 + 3 + d] = meshes[i].bbox_max[d];
        }
    }
    s->scale0 = sloth::scene_scale0(meshes);
    return s;
}

__attribute__((visibility("default"))) void sloth_host_free(sloth_host_scene* s)
{
    if (!s) return;
    std::free(s->xyz);
    std::free(s->rgb);
    std::free(s->mesh_sizes);
    std::free(s->mesh_bbox);
    std::free(s);
}

// The decimal -> f32 conversion of the device loader (csrc/dec_float.cuh), run on the host so that the CPU test
// suite can compare it with strtof on millions of tokens.  Tokens are separated by '\n'; status[i] is 0 (ok),
// 1 (not a number) or 2 (outside what the routine decides); out[i] is written for status 0.
__attribute__((visibility("default"))) size_t sloth_host_parse_f32(const char* text, size_t len, float* out, uint8_t* status,
                                                                   size_t cap)
{
    size_t n = 0, b = 0;
    for (size_t i = 0; i <= len && n < cap; ++i) {
        if (i == len || text[i] == '\n') {
            if (i > b) {
                float f = 0.0f;
                status[n] = (uint8_t)sloth::ld::parse_float(reinterpret_cast<const unsigned char*>(text + b),
                                                            reinterpret_cast<const unsigned char*>(text + i), f);
                out[n] = f;
                ++n;
            }
            b = i + 1;
        }
    }
    return n;
}
}
