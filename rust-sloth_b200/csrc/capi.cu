// capi.cu -- the C ABI of include/sloth_b200.h over the kernels in kernels.cuh.
//
// Host-side scalar arithmetic that the reference does once per frame stays on
// the host, in the reference's operation order and without FMA contraction
// (this file is compiled with -Xcompiler -ffp-contract=off):
//   Context::update matrix        src/context.rs:104-133
//   Rotation3::from_euler_angles  src/main.rs:76-77 (nalgebra 0.22.1)
//   utransform * transform        src/rasterizer.rs:57 (nalgebra gemm/gemv/axcpy order)
// There is no CPU raster fallback anywhere in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "../../include/sloth_b200.h"
#include "kernels.cuh"
#include "index.cuh"
#include "tri_kernel.cuh"
#include "tri2_kernel.cuh"
#include "tile_kernels.cuh"
#include "spans.cuh"
#ifndef STAMP_BLOCKS_PER_SM
#define STAMP_BLOCKS_PER_SM 4u
#endif
#include "../host/wire.hpp"
#include "flush.cuh"
#include "loader.cuh"

using namespace sloth;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(SLOTH_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// column-major helpers (nalgebra storage): (r,c) -> m[c*4+r]
inline float& at(float* m, int r, int c) { return m[c * 4 + r]; }
inline float at(const float* m, int r, int c) { return m[c * 4 + r]; }

// nalgebra 0.22.1 Matrix4*Matrix4: per output element, products accumulated left to right over k.
void mat4_mul(const float* A, const float* B, float* C)
{
    float out[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            float acc = at(A, i, 0) * at(B, 0, j);
            for (int k = 1; k < 4; ++k) acc = acc + at(A, i, k) * at(B, k, j);
            out[j * 4 + i] = acc;
        }
    std::memcpy(C, out, sizeof out);
}

const float k_default_thr[9] = {0.20f, 0.30f, 0.40f, 0.50f, 0.60f, 0.70f, 0.80f, 0.90f, 1.0f};
const char k_default_glyph[10] = {'.', ':', '-', '=', '+', '*', '#', '%', '@', ' '};

enum { EV_START = 0, EV_XFORM, EV_GEOM, EV_WALK, EV_RESOLVE_BEGIN, EV_END, EV_N };

}  // namespace

struct sloth_ctx {
    int device = 0;
    bool image = false;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    int sm_count = 148;

    // scene
    float4* sc_a = nullptr;
    float4* sc_b = nullptr;
    float* sc_z3 = nullptr;
    uint32_t* sc_rgb = nullptr;
    float* sc_chunks = nullptr;      // TMA feed: 1280-byte chunks of 32 triangles
    float4* sc_bounds = nullptr;     // bounding sphere per chunk of 32 triangles (band cull)
    float scene_absmax = 0.0f;       // largest |coordinate| of the scene
    uint32_t geom_blocks_per_sm = G3_BLOCKS_PER_SM;   // SLOTH_GRID overrides (profiling)
    int carveout_override = -1;
    int carveout_set = -1;            // shared-memory carveout (%) currently applied to the frame kernels
    uint32_t tail_blocks_per_sm = 8;  // SLOTH_TAIL overrides (profiling)
    uint32_t batch_max = 1;           // consecutive chunks per warp turn (SLOTH_BATCH overrides, 1..16)
    bool tma_feed = false;           // SLOTH_TMA=1 feeds k_geom3 through cp.async.bulk + mbarrier (measured 3 % slower)
    uint32_t n_tri = 0;
    float scene_max = 0.0f;
    bool have_scene = false;

    // indexed scene (index.cuh): unique vertices + per-triangle records, transformed once per frame by k_xform
    int path_pref = SLOTH_PATH_AUTO;  // sloth_ctx_set_path / SLOTH_PATH
    bool indexed = false;             // the resident scene renders through k_xform + k_tri
    uint32_t n_vert = 0;
    float* sc_pos = nullptr;          // x[n_vert+1] y[n_vert+1] z[n_vert+1] (one allocation, SoA)
    size_t pos_stride = 0;            // floats between the x, y and z arrays
    uint4* sc_rec = nullptr;          // [n_tri padded to 32]
    float2* vxy[2] = {nullptr, nullptr};   // [n_vert + 1] per frame-state set, rewritten every frame (one allocation)
    float* vz[2] = {nullptr, nullptr};
    cudaStream_t xform_stream = nullptr;   // batches: k_xform of frame k+1 runs beside k_tri of frame k
    cudaEvent_t ev_xform[2] = {nullptr, nullptr};
    cudaEvent_t ev_stamped[2] = {nullptr, nullptr};   // k_super_stamp of the set's last frame has read its vxy
    cudaEvent_t ev_batch_start = nullptr;
    // super-chunks (index.cuh): per-scene cone + unique-vertex list of every SC_TRIS (128) triangles, per-frame list of the ones
    // k_super_cert could not certify as back-facing
    ix::SuperChunk* sc_super = nullptr;
    uint32_t* sc_super_ids = nullptr;
    uint32_t* live_sc[2] = {nullptr, nullptr};
    uint32_t* skip_sc[2] = {nullptr, nullptr};
    ConeCounts* cone_cnt[2] = {nullptr, nullptr};
    uint32_t n_super = 0;             // full super-chunks of the resident scene
    bool cone = true;                 // SLOTH_CONE=0: every chunk goes through k_tri
    uint32_t tri_blocks_per_sm = T_BLOCKS_PER_SM;   // SLOTH_TGRID overrides (profiling)
    bool tri_pairs = false;           // SLOTH_TRI2=1: k_tri2 (two chunks per warp turn) for whole-frame contexts -- 5 % fewer
                                      // instructions and 2 % faster alone, but its 80 registers x 768 threads leave no room for
                                      // the neighbouring frames' k_xform / k_resolve blocks: 167 instead of 140 us per frame in batches
    uint32_t pf_chunks = 0;           // SLOTH_PF: L2 prefetch distance of k_tri's record stream (measured: hurts, off)
    size_t l2_persist_max = 0, l2_window_max = 0;   // device limits of the persisting-L2 set-aside / access window
    size_t l2_window_bytes = 0;       // bytes of (vxy, vz) currently covered by the persisting window
    bool l2_persist = false;          // SLOTH_L2PERSIST=1: measured neutral for k_tri and 4 us slower for k_resolve, so off

    // frame state
    uint32_t W = 0, H = 0;
    uint32_t row0 = 0, row1 = 0;  // band; row1 == 0 -> whole frame
    bool sized = false;
    unsigned long long* keys[2] = {nullptr, nullptr};   // two frame-state sets (geometry k+1 overlaps resolve k)
    size_t n_key_slots = 0;       // including the halo row
    uint32_t halo_slots = 0;
    uint32_t* d_cells[2] = {nullptr, nullptr};
    size_t cells_per_frame = 0;
    float* d_z = nullptr;
    cudaEvent_t ev_rendered[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};

    // queues + per-frame aux (rowmax and FrameAux are one allocation, cleared by one memset)
    uint32_t* walk_tri[2] = {nullptr, nullptr};      // k_geom3 -> k_tail queues, one per frame-state set
    unsigned long long* walk_base[2] = {nullptr, nullptr};
    uint32_t* irr_tri[2] = {nullptr, nullptr};
    // binned tile path (tile_kernels.cuh): per frame-state set, one allocation for the tile arrays
    unsigned long long* tile_info[2] = {nullptr, nullptr};   // [n_tri], with the walk queues
    TileTri* tile_setup[2] = {nullptr, nullptr};             // [n_tri]
    uint8_t* tile_region[2] = {nullptr, nullptr};            // TileAux | count[n_tiles+1] | cursor | (16-byte aligned) used | pairs
    size_t tile_clear_bytes = 0;                             // aux + counters: cleared per frame
    uint32_t tiles_x = 0, tiles_y = 0, pair_cap = 0;
    // Off by default: parity-tested, but measured slower than the walk kernel on every bundled scene (config 3,
    // "suzy suzy" at 4K: bin 6.5 + scan 15.6 + fill 12.0 + k_tile 44.7 + k_tail 6.7 = 86 us against 57 us for
    // k_tail alone; profiles/README.md).  SLOTH_TILES=1 turns it on for frames of at least 512 x 256 cells,
    // SLOTH_TILES=2 for frames of any size (tests).
    bool tile_path = false;
    bool tile_always = false;
    int last_set = 0;
    cudaStream_t resolve_stream = nullptr;
    cudaEvent_t ev_geom_done[2] = {nullptr, nullptr}, ev_resolved[2] = {nullptr, nullptr};
    uint32_t debug = 0;              // SLOTH_DEBUG bits, profiling experiments only
    bool scene_clean = false;        // every |coordinate| <= 2^20: no per-triangle regularity test needed
    uint8_t* aux_region[2] = {nullptr, nullptr};
    size_t aux_bytes = 0, rowmax_bytes = 0;

    float thr[9];
    char glyph[10];

    // device-side flush (text serialisation)
    char* d_text[2] = {nullptr, nullptr};
    size_t d_text_cap = 0;
    size_t flush_blocks_cap = 0;     // entries of flush_block_sum / flush_block_off
    uint32_t* flush_block_sum = nullptr;
    unsigned long long* flush_block_off = nullptr;
    unsigned long long* d_text_total = nullptr;      // [2]
    unsigned long long* h_text_total = nullptr;      // [2], pinned
    cudaEvent_t ev_text[2] = {nullptr, nullptr};

    // span wire format (spans.cuh, host/wire.hpp): sloth_ctx_set_wire(SLOTH_WIRE_SPANS)
    int wire = 0;
    static constexpr int WIRE_SLOTS = 6;             // page-locked staging slots (frames whose runs are in flight / being expanded)
    uint2* d_runs[2] = {nullptr, nullptr};           // runs of the frame in cell buffer k&1
    size_t runs_cap = 0;                             // runs per buffer and per staging slot (cells / 4: half the plain bytes)
    uint2* h_runs[WIRE_SLOTS] = {};
    cudaEvent_t ev_staged[WIRE_SLOTS] = {};
    sloth::WirePool* wire_pool = nullptr;
    struct WireSync* wire_sync = nullptr;            // free staging slots (shared with the pool's jobs)
    uint64_t wire_d2h_bytes = 0, wire_frames = 0, wire_plain_frames = 0;   // since sloth_ctx_set_wire

    uint32_t stat_flags = 0;
    uint64_t frames = 0, launches = 0;
    cudaEvent_t ev[EV_N] = {};
    bool ev_valid = false, ev_kernels_valid = false;
    float batch_ms_per_frame = 0.0f;
    size_t batch_frames = 0;
    bool batch_pending = false;      // device batch enqueued, its time not yet read
    bool last_was_batch = false;

    float load_ms[3] = {0.f, 0.f, 0.f};      // last sloth_scene_load: read, parse, commit
    struct LoaderState* loader = nullptr;   // staged soup segments of sloth_loader_* (loader_api.inl)
};

namespace {

int free_frame_state(sloth_ctx* c)
{
    cudaFree(c->keys[0]);
    cudaFree(c->keys[1]);
    cudaFree(c->d_cells[0]);
    cudaFree(c->d_cells[1]);
    cudaFree(c->d_z);
    cudaFree(c->aux_region[0]);
    cudaFree(c->aux_region[1]);
    cudaFree(c->tile_region[0]);
    cudaFree(c->tile_region[1]);
    c->tile_region[0] = c->tile_region[1] = nullptr;
    c->keys[0] = c->keys[1] = nullptr;
    c->d_cells[0] = c->d_cells[1] = nullptr;
    c->d_z = nullptr;
    c->aux_region[0] = c->aux_region[1] = nullptr;
    c->sized = false;
    return 0;
}

int alloc_frame_state(sloth_ctx* c)
{
    free_frame_state(c);
    const uint32_t W = c->W, H = c->H;
    const bool band = c->row1 != 0;
    const uint32_t r0 = band ? c->row0 : 0, r1 = band ? c->row1 : H;
    const uint32_t rows = r1 - r0;
    const bool even = (W & 1u) == 0;
    const uint32_t KW = even ? W / 2 : W;
    c->halo_slots = (!even && r0 > 0) ? KW : 0;
    c->n_key_slots = (size_t)rows * KW + c->halo_slots;
    c->cells_per_frame = (size_t)rows * W + ((c->image && !band) ? H : 0);
    for (int i = 0; i < 2; ++i) {
        CU(cudaMalloc(&c->keys[i], std::max<size_t>(c->n_key_slots, 1) * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->keys[i], 0xFF, std::max<size_t>(c->n_key_slots, 1) * sizeof(unsigned long long), c->stream));
    }
    for (int i = 0; i < 2; ++i) CU(cudaMalloc(&c->d_cells[i], (c->cells_per_frame + 2) * sizeof(uint32_t)));
    c->rowmax_bytes = ((((size_t)H + 31) & ~(size_t)31) + 64) * sizeof(uint32_t);
    c->aux_bytes = c->rowmax_bytes + sizeof(FrameAux);
    for (int i = 0; i < 2; ++i) CU(cudaMalloc(&c->aux_region[i], c->aux_bytes));
    {   // tile path: only worth its four launches on frames with room for large triangles
        c->tiles_x = (W + TILE_W - 1) / TILE_W;
        c->tiles_y = (H + TILE_H - 1) / TILE_H;
        const size_t n_tiles = (size_t)c->tiles_x * c->tiles_y;
        c->pair_cap = (uint32_t)std::min<size_t>(4u << 20, std::max<size_t>(n_tiles * 32, 1u << 16));
        c->tile_clear_bytes = sizeof(TileAux) + (n_tiles + 1) * sizeof(uint32_t);
        const size_t head = (sizeof(TileAux) + (2 * n_tiles + 1) * sizeof(uint32_t) + 63) & ~(size_t)63;
        const size_t bytes = head + n_tiles * sizeof(uint4) + (size_t)c->pair_cap * sizeof(TileTri);
        for (int i = 0; i < 2; ++i) CU(cudaMalloc(&c->tile_region[i], bytes));
    }
    c->sized = true;
    return SLOTH_OK;
}

void build_params(const sloth_ctx* c, const float rot[16], FrameParams& p)
{
    float ut[16];
    sloth_utransform(c->W, c->H, c->scene_max, ut);
    float M[16];
    mat4_mul(ut, rot, M);  // rasterizer.rs:57
    for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 4; ++k) p.m[r * 4 + k] = at(M, r, k);
    std::memcpy(p.thr, c->thr, sizeof p.thr);
    std::memset(p.glyph, 0, sizeof p.glyph);
    std::memcpy(p.glyph, c->glyph, 10);
    p.W = c->W;
    p.H = c->H;
    p.wm1 = (float)(c->W - 1);
    p.hm1 = (float)(c->H - 1);
    const bool even = (c->W & 1u) == 0;
    p.KW = even ? c->W / 2 : c->W;
    p.XS = even ? 1u : 2u;
    const bool band = c->row1 != 0;
    p.row0 = band ? c->row0 : 0;
    p.row1 = band ? c->row1 : c->H;
    p.krow0 = p.row0 - (c->halo_slots ? 1u : 0u);
    p.srow0 = p.row0;
    p.srow1 = p.row1;
    if (c->W == 1u) {   // one column: the stamp of row y is cell y+1, i.e. (row y+1, column 0)
        p.srow0 = p.row0 ? p.row0 - 1u : 0u;
        p.srow1 = p.row1 - 1u;
    }
    p.n_tri = c->n_tri;
    p.image = c->image ? 1u : 0u;
    p.count_frags = (c->stat_flags & 1u) ? 1u : 0u;
    p.debug = c->debug;
    // band cull (k_geom3): |row 1 of M| rounded up, and a bound on the rounding error of any computed y' -- four
    // roundings of terms that are at most (|m4|+|m5|+|m6|) * absmax + |m7| in magnitude, counted twice (vertex and
    // sphere centre) with a factor 4 to spare.  Only for scenes whose coordinates are all finite and <= 2^20.
    // k_tri's frame-wide distance bound for the back-face proof: |x'| <= (|m0|+|m1|+|m2|) * absmax + |m3| for every
    // vertex (same for y'), so D_frame = max(Xabs + (W-1), Yabs + (H-1)) >= every triangle's D; 0.1 % covers the
    // roundings of the computed coordinates and of the bound.  Unknown extent -> +inf (nothing is ever proven).
    p.bf_k = std::numeric_limits<float>::infinity();
    if (c->scene_clean) {
        const double am = (double)c->scene_absmax;
        const double xa = (std::fabs((double)p.m[0]) + std::fabs((double)p.m[1]) + std::fabs((double)p.m[2])) * am + std::fabs((double)p.m[3]);
        const double ya = (std::fabs((double)p.m[4]) + std::fabs((double)p.m[5]) + std::fabs((double)p.m[6])) * am + std::fabs((double)p.m[7]);
        const double D = std::max(xa + (double)p.wm1, ya + (double)p.hm1) * 1.001;
        const float k = (float)(D * 3.814697265625e-06 * 1.0001);   // 2^-18 D, rounded up
        if (std::isfinite(k)) p.bf_k = std::nextafterf(k, std::numeric_limits<float>::infinity());
    }
    p.pf_chunks = c->pf_chunks;
    // super-chunk certificate (k_super_cert): c = row0 x row1, |c|, the larger row norm and the coordinate error bound
    // (four roundings of terms of magnitude <= |row|_1 * absmax + |translation|, with a factor 2 to spare)
    p.cone_on = 0u;
    p.cone_n_super = 0u;
    p.cone_c[0] = p.cone_c[1] = p.cone_c[2] = p.cone_cnorm = p.cone_s = p.cone_ev = 0.0f;
    if (c->scene_clean && c->cone && c->indexed && c->n_super && !band && std::isfinite(p.bf_k)) {
        const double r0[3] = {p.m[0], p.m[1], p.m[2]}, r1[3] = {p.m[4], p.m[5], p.m[6]};
        const double cx = r0[1] * r1[2] - r0[2] * r1[1], cy = r0[2] * r1[0] - r0[0] * r1[2], cz = r0[0] * r1[1] - r0[1] * r1[0];
        const double cn = std::sqrt(cx * cx + cy * cy + cz * cz) * 1.000001;
        const double s0 = std::sqrt(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]), s1 = std::sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
        const double am = (double)c->scene_absmax;
        const double xa = (std::fabs(r0[0]) + std::fabs(r0[1]) + std::fabs(r0[2])) * am + std::fabs((double)p.m[3]);
        const double ya = (std::fabs(r1[0]) + std::fabs(r1[1]) + std::fabs(r1[2])) * am + std::fabs((double)p.m[7]);
        const double ev = std::max(xa, ya) * std::ldexp(1.0, -24) * 8.0 + 1.0e-30;
        const float f_cn = (float)cn, f_s = (float)(std::max(s0, s1) * 1.000001), f_ev = (float)(ev * 1.000001);
        if (std::isfinite(f_cn) && std::isfinite(f_s) && std::isfinite(f_ev) && f_cn > 0.0f) {
            p.cone_on = 1u;
            p.cone_n_super = c->n_super;
            p.cone_c[0] = (float)cx; p.cone_c[1] = (float)cy; p.cone_c[2] = (float)cz;
            p.cone_cnorm = std::nextafterf(f_cn, std::numeric_limits<float>::infinity());
            p.cone_s = std::nextafterf(f_s, std::numeric_limits<float>::infinity());
            p.cone_ev = std::nextafterf(f_ev, std::numeric_limits<float>::infinity());
        }
    }
    p.cull_on = 0u;
    p.cull_scale = p.cull_pad = 0.0f;
    if (band && c->scene_clean && !(c->debug & 4u)) {
        const double m4 = p.m[4], m5 = p.m[5], m6 = p.m[6], m7 = p.m[7];
        const double norm = std::sqrt(m4 * m4 + m5 * m5 + m6 * m6) * 1.0001;
        const double mag = (std::fabs(m4) + std::fabs(m5) + std::fabs(m6)) * (double)c->scene_absmax + std::fabs(m7);
        const double pad = mag * std::ldexp(1.0, -19) + 1.0e-3;
        if (std::isfinite(norm) && std::isfinite(pad) && pad < 1.0e6) {
            p.cull_on = 1u;
            p.cull_scale = (float)norm;
            p.cull_pad = (float)(pad * 1.0001);
        }
    }
}

Queues make_queues(const sloth_ctx* c, int set)
{
    Queues q;
    q.walk_tri = c->walk_tri[set];
    q.walk_base = c->walk_base[set];
    q.irr_tri = c->irr_tri[set];
    q.rowmax = reinterpret_cast<uint32_t*>(c->aux_region[set]);
    q.aux = reinterpret_cast<FrameAux*>(c->aux_region[set] + c->rowmax_bytes);
    q.live_sc = c->live_sc[set];
    q.skip_sc = c->skip_sc[set];
    q.cone_cnt = c->cone_cnt[set];
    return q;
}

Scene scene_of(const sloth_ctx* c, int set)
{
    Scene sc{c->sc_a, c->sc_b, c->sc_z3, c->sc_rgb, c->sc_bounds, c->sc_rec, c->vxy[set], c->vz[set], c->sc_super_ids};
    return sc;
}

// Triangle::mul once per unique vertex of the indexed scene, into frame-state set `set`, on stream `st`.
// a clean scene under a bounded matrix cannot produce |x'|,|y'| > 2^40: no per-triangle regularity test needed
bool bounded_frame(const sloth_ctx* c, const FrameParams& p)
{
    bool bounded = c->scene_clean;
    for (int i = 0; i < 8; ++i) bounded = bounded && std::fabs(p.m[i]) <= 131072.0f;
    return bounded;
}

// whole-frame, bounded: super-chunks certified back-facing for this frame's matrix are taken off k_tri's list
static constexpr size_t LIVE_SLACK = 65536;   // >= 4 * (warps of the largest k_tri grid: 148 SMs x 4 blocks x 16 warps)
bool cone_frame(const sloth_ctx* c, const FrameParams& p)
{
    // (the last condition: k_tri's look-ahead past the end of its work list stays inside the zeroed room behind it)
    return c->row1 == 0 && !c->tri_pairs && p.cone_on && bounded_frame(c, p) &&
           (size_t)c->sm_count * c->tri_blocks_per_sm * T_WARPS * 4u <= LIVE_SLACK;
}

// Everything of the indexed path that can run ahead of k_tri, into frame-state set `set` on stream `st`: Triangle::mul
// once per unique vertex (k_xform) and, for cone frames, the super-chunk certificate (two lists and their lengths,
// which live apart from the per-frame aux region: that one still belongs to the resolve of two frames ago when, in
// batches, this runs on the transform stream one frame ahead of the triangle kernel).
int enqueue_xform(sloth_ctx* c, const FrameParams& p, int set, cudaStream_t st)
{
    const float* px = c->sc_pos;
    const uint32_t n_threads = (c->n_vert + 1 + XFORM_PER_THREAD - 1) / XFORM_PER_THREAD;
    k_xform<<<(n_threads + CO_THREADS - 1) / CO_THREADS, CO_THREADS, 0, st>>>(p, px, px + c->pos_stride, px + 2 * c->pos_stride, c->n_vert, c->vxy[set],
                                                     c->vz[set]);
    c->launches += 1;
    if (cone_frame(c, p)) {
        CU(cudaMemsetAsync(c->cone_cnt[set], 0, sizeof(ConeCounts), st));
        k_super_cert<<<(c->n_super + 255) / 256, 256, 0, st>>>(p, c->sc_super, c->n_super, make_queues(c, set));
        c->launches += 1;
    }
    return SLOTH_OK;
}

// Kernels that run side by side on one SM (the geometry kernel of frame k+1 with k_tail / k_resolve of frame k)
// must agree on the L1 / shared-memory split, or the SM has to drain before it can switch: one carveout for all
// frame kernels, re-applied only when it changes.
int apply_carveout(sloth_ctx* c, int pct)
{
    if (c->carveout_override >= 0) pct = c->carveout_override;
    if (pct == c->carveout_set) return SLOTH_OK;
    const cudaFuncAttribute a = cudaFuncAttributePreferredSharedMemoryCarveout;
    CU(cudaFuncSetAttribute(k_tail, a, pct));
    CU(cudaFuncSetAttribute(k_resolve_even, a, pct));
    CU(cudaFuncSetAttribute(k_resolve_odd, a, pct));
    CU(cudaFuncSetAttribute(k_clear_keys_odd, a, pct));
    CU(cudaFuncSetAttribute(k_xform, a, pct));
    CU(cudaFuncSetAttribute(k_bin_count, a, pct));
    CU(cudaFuncSetAttribute(k_bin_scan, a, pct));
    CU(cudaFuncSetAttribute(k_bin_fill, a, pct));
    CU(cudaFuncSetAttribute(k_tile, a, pct));
    CU(cudaFuncSetAttribute(k_tri<false, false, true, true>, a, pct));
    CU(cudaFuncSetAttribute(k_tri<false, false, false, true>, a, pct));
    CU(cudaFuncSetAttribute(k_super_cert, a, pct));
    CU(cudaFuncSetAttribute(k_super_stamp, a, pct));
    CU(cudaFuncSetAttribute(k_tri2<false, true>, a, pct));
    CU(cudaFuncSetAttribute(k_tri2<true, true>, a, pct));
    CU(cudaFuncSetAttribute(k_tri2<false, false>, a, pct));
    CU(cudaFuncSetAttribute(k_tri2<true, false>, a, pct));
    CU(cudaFuncSetAttribute(k_tri<false, false, true>, a, pct));
    CU(cudaFuncSetAttribute(k_tri<false, true, true>, a, pct));
    CU(cudaFuncSetAttribute(k_tri<true, false, true>, a, pct));
    CU(cudaFuncSetAttribute(k_tri<true, true, true>, a, pct));
    CU(cudaFuncSetAttribute(k_tri<false, false, false>, a, pct));
    CU(cudaFuncSetAttribute(k_tri<false, true, false>, a, pct));
    CU(cudaFuncSetAttribute(k_tri<true, false, false>, a, pct));
    CU(cudaFuncSetAttribute(k_tri<true, true, false>, a, pct));
    CU(cudaFuncSetAttribute(k_geom3<false, false, false>, a, pct));
    CU(cudaFuncSetAttribute(k_geom3<false, true, false>, a, pct));
    CU(cudaFuncSetAttribute(k_geom3<true, false, false>, a, pct));
    CU(cudaFuncSetAttribute(k_geom3<true, true, false>, a, pct));
    CU(cudaFuncSetAttribute(k_geom3<false, false, true>, a, pct));
    CU(cudaFuncSetAttribute(k_geom3<false, true, true>, a, pct));
    CU(cudaFuncSetAttribute(k_geom3<true, false, true>, a, pct));
    CU(cudaFuncSetAttribute(k_geom3<true, true, true>, a, pct));
    c->carveout_set = pct;
    return SLOTH_OK;
}

// Indexed path: k_xform (Triangle::mul once per unique vertex; skipped when the caller has already run it for this
// set on another stream) then k_tri, on stream `st`.
int enqueue_geometry_indexed(sloth_ctx* c, const FrameParams& p, const Scene& sc, const Queues& q, int set, cudaStream_t st,
                             bool kt, bool xform_done)
{
    const uint32_t n_chunks = (c->n_tri + 31) / 32;
    const uint32_t bps = c->tri_blocks_per_sm;
    const uint32_t grid = std::min<uint32_t>((n_chunks + T_WARPS - 1) / T_WARPS, (uint32_t)c->sm_count * bps);
    const bool bounded = bounded_frame(c, p);
    // the per-block copy of rowmax moves to shared memory while it does not cost a resident block (227 KB per SM,
    // 1 KB reserved per block) -- or, for the tallest frames the 8-warp layout still takes, up to 33 KB
    const size_t warp_smem = sizeof(TWarpSmem) * T_WARPS;
    const bool rowmax_shared = c->rowmax_bytes <= 33024u && (T_WARPS <= 8 || (c->rowmax_bytes + warp_smem + 1024u) * bps <= 227u * 1024u);
    const size_t dyn = (rowmax_shared ? c->rowmax_bytes : 0) + warp_smem;
    const bool band_mode = c->row1 != 0;
    if (!band_mode && c->tri_pairs) {
        // whole-frame contexts: two chunks per warp turn (tri2_kernel.cuh).  The per-block rowmax copy stays in shared
        // memory only while it does not cost a resident block.
        const uint32_t n_pairs = (n_chunks + 1) / 2;
        const uint32_t grid2 = std::min<uint32_t>((n_pairs + T_WARPS - 1) / T_WARPS, (uint32_t)c->sm_count * bps);
        const size_t warp_smem2 = sizeof(TWarpSmem2) * T_WARPS;
        const bool shared2 = (c->rowmax_bytes + warp_smem2 + 1024u) * bps <= 228u * 1024u;
        const size_t dyn2 = (shared2 ? c->rowmax_bytes : 0) + warp_smem2;
        void (*kern2)(FrameParams, Scene, unsigned long long*, Queues) =
            shared2 ? (bounded ? k_tri2<false, true> : k_tri2<true, true>) : (bounded ? k_tri2<false, false> : k_tri2<true, false>);
        int pct = (int)(((dyn2 + 1024) * bps * 100 + 228 * 1024 - 1) / (228 * 1024)) + 3;
        pct = std::min(100, std::max(25, pct));
        const int rc = apply_carveout(c, pct);
        if (rc) return rc;
        if (!xform_done) enqueue_xform(c, p, set, st);
        if (kt) CU(cudaEventRecord(c->ev[EV_XFORM], st));
        kern2<<<grid2, T_WARPS * 32, dyn2, st>>>(p, sc, c->keys[set], q);
        c->launches += 1;
        if (kt) CU(cudaEventRecord(c->ev[EV_GEOM], st));
        return SLOTH_OK;
    }
    if (cone_frame(c, p)) {
        // k_super_cert (enqueue_xform) has split the super-chunks into the certified ones (k_super_stamp adds their row
        // stamps before the resolve) and the rest, whose chunks k_tri<CONE> works through
        const int rc = apply_carveout(c, std::min(100, std::max(25, (int)(((dyn + 1024) * bps * 100 + 228 * 1024 - 1) / (228 * 1024)) + 3)));
        if (rc) return rc;
        if (!xform_done) enqueue_xform(c, p, set, st);
        if (kt) CU(cudaEventRecord(c->ev[EV_XFORM], st));
        if (rowmax_shared) k_tri<false, false, true, true><<<grid, T_WARPS * 32, dyn, st>>>(p, sc, c->keys[set], q);
        else k_tri<false, false, false, true><<<grid, T_WARPS * 32, dyn, st>>>(p, sc, c->keys[set], q);
        c->launches += 1;
        if (kt) CU(cudaEventRecord(c->ev[EV_GEOM], st));
        return SLOTH_OK;
    }
    void (*kern)(FrameParams, Scene, unsigned long long*, Queues);
    if (rowmax_shared)
        kern = bounded ? (band_mode ? k_tri<false, true, true> : k_tri<false, false, true>)
                       : (band_mode ? k_tri<true, true, true> : k_tri<true, false, true>);
    else
        kern = bounded ? (band_mode ? k_tri<false, true, false> : k_tri<false, false, false>)
                       : (band_mode ? k_tri<true, true, false> : k_tri<true, false, false>);
    {
        const size_t per_block = dyn + 1024;
        int pct = (int)((per_block * bps * 100 + 228 * 1024 - 1) / (228 * 1024)) + 3;
        pct = std::min(100, std::max(25, pct));
        const int rc = apply_carveout(c, pct);
        if (rc) return rc;
    }
    if (!xform_done) enqueue_xform(c, p, set, st);
    if (kt) CU(cudaEventRecord(c->ev[EV_XFORM], st));
    kern<<<grid, T_WARPS * 32, dyn, st>>>(p, sc, c->keys[set], q);
    c->launches += 1;
    if (kt) CU(cudaEventRecord(c->ev[EV_GEOM], st));
    return SLOTH_OK;
}

// Geometry pass of a frame (aux clear, k_geom3) on stream `st`, into frame-state set `set`.
int enqueue_geometry(sloth_ctx* c, const FrameParams& p, int set, cudaStream_t st, bool kt, bool xform_done = false)
{
    const Scene sc = scene_of(c, set);
    const Queues q = make_queues(c, set);
    CU(cudaMemsetAsync(c->aux_region[set], 0, c->aux_bytes, st));
    if (c->indexed && c->n_tri) return enqueue_geometry_indexed(c, p, sc, q, set, st, kt, xform_done);
    if (c->n_tri) {
        {
            const uint32_t n_chunks = (c->n_tri + 31) / 32;
            // consecutive chunks per warp turn.  1 = chunk i goes to warp i mod n_warps: the finest interleave, so
            // that every SM sees the same mix of cheap (back-facing) and expensive regions of the soup.  Longer
            // runs (16 was the first choice, for locality of rows and cache lines) leave the SMs up to 25 % apart
            // at the end of the kernel: 160 -> 154 us at 4K and 0.60 -> 0.48 ms at 8K for the 10 M-triangle sphere.
            const uint32_t blocks_per_sm = c->geom_blocks_per_sm;
            const uint32_t warps_avail = (uint32_t)c->sm_count * blocks_per_sm * G3_WARPS;
            const uint32_t batch_chunks = std::max<uint32_t>(1u, std::min<uint32_t>(c->batch_max, n_chunks / (warps_avail * 4u)));
            const uint32_t n_batches = (n_chunks + batch_chunks - 1) / batch_chunks;
            const uint32_t grid = std::min<uint32_t>((n_batches + G3_WARPS - 1) / G3_WARPS, (uint32_t)c->sm_count * blocks_per_sm);
            // a clean scene under a bounded matrix cannot produce |x'|,|y'| > 2^40: skip the per-triangle test
            bool bounded = c->scene_clean;
            for (int i = 0; i < 8; ++i) bounded = bounded && std::fabs(p.m[i]) <= 131072.0f;
            // per-block row-stamp array in shared memory while 3 blocks/SM still fit
            const bool tma = c->tma_feed;
            const size_t ring_bytes = tma ? sizeof(TmaRing) * G3_WARPS : 0;
            const uint32_t rowmax_shared = c->rowmax_bytes <= (tma ? 17664u : 33024u) ? 1u : 0u;
            const size_t dyn = ring_bytes + (rowmax_shared ? c->rowmax_bytes : 0);
            const bool band_mode = c->row1 != 0;
            void (*kern)(FrameParams, Scene, const float*, unsigned long long*, Queues, uint32_t, uint32_t);
            if (tma) kern = bounded ? (band_mode ? k_geom3<false, true, true> : k_geom3<false, false, true>)
                                    : (band_mode ? k_geom3<true, true, true> : k_geom3<true, false, true>);
            else kern = bounded ? (band_mode ? k_geom3<false, true, false> : k_geom3<false, false, false>)
                                : (band_mode ? k_geom3<true, true, false> : k_geom3<true, false, false>);
            {
                // one carveout that holds three geometry blocks, shared by all frame kernels (apply_carveout)
                const size_t per_block = sizeof(G3Queue) * G3_WARPS + dyn + 1024;
                int pct = (int)((per_block * blocks_per_sm * 100 + 228 * 1024 - 1) / (228 * 1024)) + 3;
                pct = std::min(100, std::max(50, pct));
                const int rc = apply_carveout(c, pct);
                if (rc) return rc;
            }
            kern<<<grid, G3_WARPS * 32, dyn, st>>>(p, sc, c->sc_chunks, c->keys[set], q, batch_chunks, rowmax_shared);
        }
        c->launches += 1;
    }
    if (kt) CU(cudaEventRecord(c->ev[EV_GEOM], st));
    return SLOTH_OK;
}

TileState tile_state(const sloth_ctx* c, int set)
{
    const size_t n_tiles = (size_t)c->tiles_x * c->tiles_y;
    uint8_t* base = c->tile_region[set];
    TileState ts;
    ts.aux = reinterpret_cast<TileAux*>(base);
    ts.count = reinterpret_cast<uint32_t*>(base + sizeof(TileAux));
    ts.cursor = ts.count + n_tiles + 1;
    const size_t head = (sizeof(TileAux) + (2 * n_tiles + 1) * sizeof(uint32_t) + 63) & ~(size_t)63;
    ts.used = reinterpret_cast<uint4*>(base + head);
    ts.pairs = reinterpret_cast<TileTri*>(base + head + n_tiles * sizeof(uint4));
    ts.info = c->tile_info[set];
    ts.setup = c->tile_setup[set];
    ts.tiles_x = c->tiles_x;
    ts.tiles_y = c->tiles_y;
    ts.pair_cap = c->pair_cap;
    return ts;
}

// Follow-up pass of a frame on stream `st`: the triangles the geometry kernel queued (larger than 8 x 8 candidates)
// are binned to screen tiles and rasterised per tile (k_bin_count / k_bin_scan / k_bin_fill / k_tile); what cannot be
// binned (slivers whose rows do not provably end at the bounding box), and the non-finite triangles, go to k_tail.
// Small frames skip the tile path: its four launches cost more than the few large triangles such a frame can hold.
int enqueue_tail(sloth_ctx* c, const FrameParams& p, int set, cudaStream_t st)
{
    if (!c->n_tri) return SLOTH_OK;
    const Scene sc = scene_of(c, set);
    const Queues q = make_queues(c, set);
    if (c->indexed && cone_frame(c, p) && p.image && !(p.debug & 2u)) {
        // row stamps of the super-chunks k_super_cert took off k_tri's list (reads this set's transformed vertices:
        // the transform of the frame after next waits for ev_stamped)
        // a few warps per SM walk the skip list (its length is only known on the device)
        k_super_stamp<<<std::min<uint32_t>((c->n_super + 3) / 4, (uint32_t)c->sm_count * STAMP_BLOCKS_PER_SM), 128, 0, st>>>(p, c->sc_super_ids, c->vxy[set], q);
        c->launches += 1;
        CU(cudaEventRecord(c->ev_stamped[set], st));
    }
    const bool tiles = c->tile_path && (c->tile_always || (size_t)c->W * c->H >= (size_t)512 * 256);
    if (tiles) {
        const TileState ts = tile_state(c, set);
        CU(cudaMemsetAsync(c->tile_region[set], 0, c->tile_clear_bytes, st));
        const unsigned bin_blocks = (unsigned)c->sm_count * 4u;
        k_bin_count<<<bin_blocks, 256, 0, st>>>(p, sc, q, ts);
        k_bin_scan<<<1, 1024, 0, st>>>(ts);
        k_bin_fill<<<bin_blocks, 256, 0, st>>>(q, ts);
        k_tile<<<(unsigned)c->sm_count * 8u, TILE_W * TILE_H, 0, st>>>(p, sc, c->keys[set], q, ts);
        c->launches += 4;
    }
    const uint32_t wb = (uint32_t)c->sm_count * c->tail_blocks_per_sm, ib = std::max<uint32_t>(1u, (uint32_t)c->sm_count / 2u);
    k_tail<<<wb + ib, 128, 0, st>>>(p, sc, c->keys[set], q, wb, tiles ? c->tile_info[set] : nullptr);
    c->launches += 1;
    return SLOTH_OK;
}

// Resolve half of a frame (optional z plane, key plane -> cells, key plane reset) on stream `st`.
int enqueue_resolve(sloth_ctx* c, const FrameParams& p, int set, cudaStream_t st, uint32_t* d_out, float* d_z)
{
    const Scene sc = scene_of(c, set);
    const Queues q = make_queues(c, set);
    const bool band = c->row1 != 0;
    const uint32_t rows = p.row1 - p.row0;
    const uint32_t n_tail = (c->image && !band) ? c->H : 0;
    if (d_z) {
        const uint32_t n = rows * c->W;
        // z planes are only defined for whole-frame contexts (checked by the caller)
        k_zbuffer<<<(n + 255) / 256, 256, 0, st>>>(p, c->keys[set], d_z, n);
        c->launches += 1;
    }
    if ((c->W & 1u) == 0) {
        const uint32_t n_slots = rows * p.KW;
        // grid = (segments of a row, rows): a block owns RESOLVE_SEG consecutive slots of one row (kernels.cuh)
        const dim3 grid(std::max<uint32_t>(1u, (p.KW + RESOLVE_SEG - 1) / RESOLVE_SEG), rows);
        if (rows) k_resolve_even<<<grid, CO_THREADS, 0, st>>>(p, sc, c->keys[set], q, d_out, n_slots, n_tail);
        c->launches += 1;
    } else {
        const uint32_t n_cells = rows * c->W;
        const uint32_t n = n_cells + n_tail;
        if (n) k_resolve_odd<<<(n + 255) / 256, 256, 0, st>>>(p, sc, c->keys[set], q, d_out, n_cells, n_tail, c->halo_slots);
        const uint32_t nk = (uint32_t)c->n_key_slots;
        if (nk) k_clear_keys_odd<<<(nk + 255) / 256, 256, 0, st>>>(c->keys[set], nk);
        c->launches += 2;
    }
    return SLOTH_OK;
}

// Enqueue one frame on c->stream (frame-state set 0); the cells land in d_out (device).
int enqueue_frame(sloth_ctx* c, const float rot[16], uint32_t* d_out, float* d_z, bool timed)
{
    FrameParams p;
    build_params(c, rot, p);
    cudaStream_t st = c->stream;
    const bool kt = timed && (c->stat_flags & 2u);
    if (timed) CU(cudaEventRecord(c->ev[EV_START], st));
    int rc = enqueue_geometry(c, p, 0, st, kt);
    if (rc) return rc;
    rc = enqueue_tail(c, p, 0, st);
    if (rc) return rc;
    if (kt) CU(cudaEventRecord(c->ev[EV_WALK], st));
    if (kt) CU(cudaEventRecord(c->ev[EV_RESOLVE_BEGIN], st));
    rc = enqueue_resolve(c, p, 0, st, d_out, d_z);
    if (rc) return rc;
    if (timed) CU(cudaEventRecord(c->ev[EV_END], st));
    CU(cudaGetLastError());
    c->frames += 1;
    c->last_set = 0;
    return SLOTH_OK;
}

// Frames k = 0..n-1 with the geometry of frame k+1 (issue-bound, stream G = c->stream) overlapping the
// follow-up pass and the resolve of frame k (stream R): two frame-state sets alternate.  Frame k's cells go to
// out(k) on the device; before_resolve(k) / after_resolve(k) run right before / after resolve k is
// enqueued on R (used to chain the device->host copies of the host batch).  On return everything has been enqueued and G waits for the last resolves.
template <typename BeforeFn, typename OutFn, typename AfterFn>
int enqueue_overlapped(sloth_ctx* c, const float* rots, size_t n_frames, BeforeFn before_resolve, OutFn out,
                       AfterFn after_resolve)
{
    CU(cudaEventRecord(c->ev_batch_start, c->stream));   // work enqueued on the context stream before this batch comes first
    // SLOTH_DEBUG bit 9: timeline of the first frames of the batch (timing events on all three streams), printed to stderr
    const bool trace = (c->debug & 512u) && n_frames >= 4;
    const size_t n_trace = trace ? std::min<size_t>(n_frames, 12) : 0;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
    mark(c->stream);
    for (size_t k = 0; k < n_frames; ++k) {
        const bool tr = k < n_trace;
        const int set = (int)(k & 1);
        FrameParams p;
        build_params(c, rots + 16 * k, p);
        const bool split = c->indexed && c->n_tri;
        if (split) {   // k_xform(k) on its own stream: it runs beside k_tri(k-1), as soon as k_tri(k-2) has let go of the set
            if (k >= 2) {
                CU(cudaStreamWaitEvent(c->xform_stream, c->ev_geom_done[set], 0));
                CU(cudaStreamWaitEvent(c->xform_stream, c->ev_stamped[set], 0));   // no-op unless a cone frame used the set
            } else {
                CU(cudaStreamWaitEvent(c->xform_stream, c->ev_batch_start, 0));
            }
            if (tr) mark(c->xform_stream);
            if (!(c->debug & 64u) || k < 2) enqueue_xform(c, p, set, c->xform_stream);   // bit 6: timing experiment (wrong frames)
            if (tr) mark(c->xform_stream);
            CU(cudaEventRecord(c->ev_xform[set], c->xform_stream));
            CU(cudaStreamWaitEvent(c->stream, c->ev_xform[set], 0));
        }
        if (k >= 2) CU(cudaStreamWaitEvent(c->stream, c->ev_resolved[set], 0));   // set's key plane and aux are free again
        if (tr) mark(c->stream);
        int rc = enqueue_geometry(c, p, set, c->stream, false, split);
        if (rc) return rc;
        if (tr) mark(c->stream);
        CU(cudaEventRecord(c->ev_geom_done[set], c->stream));
        CU(cudaStreamWaitEvent(c->resolve_stream, c->ev_geom_done[set], 0));
        if (tr) mark(c->resolve_stream);
        if (!(c->debug & 128u)) rc = enqueue_tail(c, p, set, c->resolve_stream);   // k_tail(k) and resolve(k) run beside k_geom3(k+1)
        if (rc) return rc;
        rc = before_resolve(k);
        if (rc) return rc;
        if (!(c->debug & 256u)) rc = enqueue_resolve(c, p, set, c->resolve_stream, out(k), nullptr);   // bits 7, 8: timing experiments
        if (rc) return rc;
        if (tr) mark(c->resolve_stream);
        CU(cudaEventRecord(c->ev_resolved[set], c->resolve_stream));
        rc = after_resolve(k);
        if (rc) return rc;
        c->frames += 1;
        c->last_set = set;
    }
    for (int set = 0; set < 2 && (size_t)set < n_frames; ++set) CU(cudaStreamWaitEvent(c->stream, c->ev_resolved[set], 0));
    CU(cudaGetLastError());
    if (trace) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaStreamSynchronize(c->resolve_stream));
        const size_t per = (c->indexed && c->n_tri) ? 6 : 4;
        std::fprintf(stderr, "[sloth trace] us since batch start: frame  xform[begin end]  geometry[begin end]  tail+resolve[begin end]\n");
        for (size_t k = 0; k < n_trace; ++k) {
            std::fprintf(stderr, "[sloth trace] %2zu", k);
            for (size_t j = 0; j < per; ++j) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, tev[0], tev[1 + k * per + j]);
                std::fprintf(stderr, " %8.1f", ms * 1e3f);
            }
            std::fprintf(stderr, "\n");
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    return SLOTH_OK;
}

size_t text_bytes_per_cell(int mode) { return mode == 0 ? 1 : (mode == 1 ? 40 : 38); }

int ensure_text_buffers(sloth_ctx* c, int mode)
{
    const size_t need = c->cells_per_frame * text_bytes_per_cell(mode) + 64;
    const size_t nb = (c->cells_per_frame + FLUSH_CELLS_PER_BLOCK - 1) / FLUSH_CELLS_PER_BLOCK + 1;
    if (!c->h_text_total) {
        CU(cudaHostAlloc(&c->h_text_total, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
        CU(cudaMalloc(&c->d_text_total, 2 * sizeof(unsigned long long)));
        for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&c->ev_text[i], cudaEventDisableTiming));
    }
    // the text buffers grow with the bytes per frame, the per-block arrays with the CELLS per frame: a resize to
    // more cells in a mode with fewer bytes per cell must still grow the latter (tracked separately)
    if (need > c->d_text_cap || nb > c->flush_blocks_cap) {
        CU(cudaStreamSynchronize(c->resolve_stream));
        CU(cudaStreamSynchronize(c->copy_stream));
        for (int i = 0; i < 2; ++i) { cudaFree(c->d_text[i]); c->d_text[i] = nullptr; }
        cudaFree(c->flush_block_sum); cudaFree(c->flush_block_off);
        c->flush_block_sum = nullptr; c->flush_block_off = nullptr;
        const size_t bytes = std::max(need, c->d_text_cap);
        c->d_text_cap = 0;            // nothing counts as allocated until every cudaMalloc below has succeeded
        c->flush_blocks_cap = 0;
        for (int i = 0; i < 2; ++i) CU(cudaMalloc(&c->d_text[i], bytes));
        CU(cudaMalloc(&c->flush_block_sum, nb * sizeof(uint32_t)));
        CU(cudaMalloc(&c->flush_block_off, nb * sizeof(unsigned long long)));
        c->d_text_cap = bytes;
        c->flush_blocks_cap = nb;
    }
    return SLOTH_OK;
}

// Context::flush on the device: d_cells (n_cells) -> d_text, total length to d_total (all on `st`).
int enqueue_flush(sloth_ctx* c, int mode, const uint32_t* d_cells, size_t n_cells, char* d_text,
                  unsigned long long* d_total, uint32_t* block_sum, unsigned long long* block_off, cudaStream_t st)
{
    const uint32_t nb = (uint32_t)((n_cells + FLUSH_CELLS_PER_BLOCK - 1) / FLUSH_CELLS_PER_BLOCK);
    if (nb == 0) {
        CU(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st));
        return SLOTH_OK;
    }
    const size_t smem = FLUSH_CELLS_PER_BLOCK * text_bytes_per_cell(mode) + 16;
    k_flush_sizes<<<nb, FLUSH_THREADS, 0, st>>>(d_cells, (uint32_t)n_cells, mode, block_sum);
    k_flush_scan<<<1, 1024, 0, st>>>(block_sum, nb, block_off, d_total);
    k_flush_write<<<nb, FLUSH_THREADS, smem, st>>>(d_cells, (uint32_t)n_cells, mode, block_off, d_text);
    c->launches += 3;
    CU(cudaGetLastError());
    return SLOTH_OK;
}

int check_ready(sloth_ctx* c)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (!c->have_scene) return fail(SLOTH_E_STATE, "sloth_scene_set has not been called");
    if (!c->sized) return fail(SLOTH_E_STATE, "sloth_ctx_resize has not been called");
    CU(cudaSetDevice(c->device));
    return SLOTH_OK;
}

void free_index(sloth_ctx* c)
{
    cudaFree(c->sc_pos); cudaFree(c->sc_rec); cudaFree(c->vxy[0]);   // both sets of (vxy, vz) live in one allocation
    cudaFree(c->sc_super); cudaFree(c->sc_super_ids); cudaFree(c->live_sc[0]); cudaFree(c->live_sc[1]); cudaFree(c->skip_sc[0]); cudaFree(c->skip_sc[1]); cudaFree(c->cone_cnt[0]); cudaFree(c->cone_cnt[1]);
    c->sc_super = nullptr; c->sc_super_ids = nullptr; c->live_sc[0] = c->live_sc[1] = c->skip_sc[0] = c->skip_sc[1] = nullptr;
    c->cone_cnt[0] = c->cone_cnt[1] = nullptr;
    c->n_super = 0;
    c->sc_pos = nullptr; c->sc_rec = nullptr;
    c->vxy[0] = c->vxy[1] = nullptr; c->vz[0] = c->vz[1] = nullptr;
    if (c->l2_window_bytes) {   // drop the residency window of the transformed vertices
        cudaStreamAttrValue attr;
        std::memset(&attr, 0, sizeof attr);
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaCtxResetPersistingL2Cache();
        c->l2_window_bytes = 0;
    }
    c->indexed = false;
    c->n_vert = 0;
    c->pos_stride = 0;
}

// Room for n_vert unique vertices (+ the sentinel) and the records of n_tri triangles (padded to 64: k_tri2 takes
// two chunks of 32 per turn).
int alloc_index(sloth_ctx* c, size_t n_vert, size_t n_tri)
{
    c->pos_stride = (n_vert + 1 + 63) & ~(size_t)63;
    const size_t n_padded = (n_tri + 63) & ~(size_t)63;
    CU(cudaMalloc(&c->sc_pos, 3 * c->pos_stride * sizeof(float)));
    CU(cudaMalloc(&c->sc_rec, std::max<size_t>(n_padded, 64) * sizeof(uint4)));
    // (x', y') and z' of every vertex in one allocation, so that one L2 access-policy window covers both
    const size_t n_slots = (n_vert + 1 + XFORM_PER_THREAD - 1) / XFORM_PER_THREAD * XFORM_PER_THREAD;   // k_xform writes whole groups
    const size_t xy_bytes = (n_slots * sizeof(float2) + 255) & ~(size_t)255;
    const size_t all_bytes = xy_bytes + n_slots * sizeof(float);
    const size_t set_bytes = (all_bytes + 255) & ~(size_t)255;
    CU(cudaMalloc(&c->vxy[0], 2 * set_bytes));
    c->vxy[1] = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(c->vxy[0]) + set_bytes);
    for (int i = 0; i < 2; ++i) c->vz[i] = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(c->vxy[i]) + xy_bytes);
    c->n_vert = (uint32_t)n_vert;
    // Keep the transformed vertices resident in L2 between k_xform (writes them) and k_tri (gathers them): the
    // triangle records stream past them at 16 B/triangle and would otherwise push half of them out to HBM
    // (ncu: 50 % of the gather sectors missed L2).  Persisting lines live in a set-aside part of L2; when the
    // set-aside is smaller than the window, hitRatio keeps only that fraction persisting (no thrash).
    c->l2_window_bytes = 0;
    if (c->l2_persist && c->l2_persist_max && c->l2_window_max && all_bytes >= (1u << 20)) {
        const size_t window = std::min(all_bytes, c->l2_window_max);
        const size_t carve = std::min(window, c->l2_persist_max);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
            cudaStreamAttrValue attr;
            std::memset(&attr, 0, sizeof attr);
            attr.accessPolicyWindow.base_ptr = c->vxy[0];
            attr.accessPolicyWindow.num_bytes = window;
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)window);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess)
                c->l2_window_bytes = window;
        }
        cudaGetLastError();   // residency control is an optimisation: never fatal
    }
    return SLOTH_OK;
}

// Deduplicate the corners of the resident soup by bit pattern (index.cuh) and decide whether the scene renders
// through the indexed path: SLOTH_PATH_AUTO takes it when triangles share vertices (at most 1.5 unique vertices
// per triangle; a soup without sharing has 3 and is better off with k_geom3).
int build_index(sloth_ctx* c, size_t n_tri)
{
    free_index(c);
    if (!n_tri || c->path_pref == SLOTH_PATH_SOUP) return SLOTH_OK;
    const Scene sc = scene_of(c, 0);
    const size_t n_corners = 3 * n_tri;
    size_t cap = 64;
    while (cap < 2 * n_corners) cap <<= 1;
    const uint32_t n_blocks = (uint32_t)((n_corners + ix::SCAN_BLOCK - 1) / ix::SCAN_BLOCK);
    uint32_t *table = nullptr, *rep = nullptr, *block_sum = nullptr;
    auto drop = [&]() { cudaFree(table); cudaFree(rep); cudaFree(block_sum); };
#define CU_IX(call)                                                                                \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            drop();                                                                                \
            free_index(c);                                                                         \
            return fail(SLOTH_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                          \
    } while (0)
    CU_IX(cudaMalloc(&table, cap * sizeof(uint32_t)));
    CU_IX(cudaMalloc(&rep, n_corners * sizeof(uint32_t)));
    CU_IX(cudaMalloc(&block_sum, (n_blocks + 1) * sizeof(uint32_t)));
    CU_IX(cudaMemsetAsync(table, 0xFF, cap * sizeof(uint32_t), c->stream));
    const unsigned cb = (unsigned)((n_corners + 255) / 256);
    ix::k_ix_insert<<<cb, 256, 0, c->stream>>>(sc, (uint32_t)n_corners, table, (uint32_t)(cap - 1));
    ix::k_ix_lookup<<<cb, 256, 0, c->stream>>>(sc, (uint32_t)n_corners, table, (uint32_t)(cap - 1), rep);
    ix::k_ix_flag_sums<<<n_blocks, 256, 0, c->stream>>>(rep, (uint32_t)n_corners, block_sum);
    ix::k_ix_scan_sums<<<1, 1024, 0, c->stream>>>(block_sum, n_blocks, block_sum + n_blocks);
    c->launches += 4;
    uint32_t n_vert = 0;
    CU_IX(cudaMemcpyAsync(&n_vert, block_sum + n_blocks, sizeof n_vert, cudaMemcpyDeviceToHost, c->stream));
    CU_IX(cudaStreamSynchronize(c->stream));
    CU_IX(cudaGetLastError());
    const bool take = c->path_pref == SLOTH_PATH_INDEXED || (size_t)n_vert * 2 <= n_tri * 3;
    if (take) {
        const int rc = alloc_index(c, n_vert, n_tri);
        if (rc) { drop(); free_index(c); return rc; }
        uint32_t* rank = table;   // the table is not needed any more: its memory holds the ranks
        float* px = c->sc_pos;
        ix::k_ix_rank<<<n_blocks, 256, 0, c->stream>>>(sc, rep, (uint32_t)n_corners, block_sum, rank, px, px + c->pos_stride,
                                                      px + 2 * c->pos_stride);
        const size_t n_padded = (n_tri + 63) & ~(size_t)63;
        ix::k_ix_records<<<(unsigned)((n_padded + 255) / 256), 256, 0, c->stream>>>(rep, rank, (uint32_t)n_tri, (uint32_t)n_padded, n_vert,
                                                                                  c->sc_rec);
        ix::k_ix_connectivity<<<(unsigned)((n_padded + 255) / 256), 256, 0, c->stream>>>(c->sc_rec, (uint32_t)n_tri, (uint32_t)(n_padded / 32));
        c->launches += 3;
        c->n_super = (uint32_t)(n_tri / ix::SC_TRIS);
        if (c->n_super) {
            CU_IX(cudaMalloc(&c->sc_super, (size_t)c->n_super * sizeof(ix::SuperChunk)));
            CU_IX(cudaMalloc(&c->sc_super_ids, (size_t)c->n_super * ix::SC_IDS * sizeof(uint32_t)));
            // k_tri<CONE>'s flat work list: the chunks behind the last full super-chunk come first and never change
            const uint32_t n_chunks_all = (uint32_t)((n_tri + 31) / 32), tail_first = c->n_super * ix::SC_CHUNKS;
            std::vector<uint32_t> tail(n_chunks_all - tail_first);
            for (size_t i = 0; i < tail.size(); ++i) tail[i] = tail_first + (uint32_t)i;
            for (int i = 0; i < 2; ++i) {
                // + room for k_tri's look-ahead past the end of the list (4 iterations of every resident warp), zeroed:
                // every entry is a valid chunk index at all times
                CU_IX(cudaMalloc(&c->live_sc[i], ((size_t)n_chunks_all + LIVE_SLACK) * sizeof(uint32_t)));
                CU_IX(cudaMemsetAsync(c->live_sc[i], 0, ((size_t)n_chunks_all + LIVE_SLACK) * sizeof(uint32_t), c->stream));
                // (pageable source: staged before the call returns; ordered on c->stream, which is synchronised below)
                if (!tail.empty())
                    CU_IX(cudaMemcpyAsync(c->live_sc[i], tail.data(), tail.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
            }
            for (int i = 0; i < 2; ++i) CU_IX(cudaMalloc(&c->skip_sc[i], (size_t)c->n_super * sizeof(uint32_t)));
            for (int i = 0; i < 2; ++i) CU_IX(cudaMalloc(&c->cone_cnt[i], sizeof(ConeCounts)));
            ix::k_ix_super<<<c->n_super, ix::SC_TRIS, 0, c->stream>>>(c->sc_rec, px, px + c->pos_stride, px + 2 * c->pos_stride, c->sc_super,
                                                              c->sc_super_ids);
            c->launches += 1;
        }
        CU_IX(cudaGetLastError());
        CU_IX(cudaStreamSynchronize(c->stream));
        c->indexed = true;
    }
#undef CU_IX
    drop();
    return SLOTH_OK;
}

// Drop the resident scene and allocate room for n_tri triangles (plus the per-scene queues).
int alloc_scene(sloth_ctx* c, size_t n_tri)
{
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->resolve_stream));
    free_index(c);
    cudaFree(c->sc_a); cudaFree(c->sc_b); cudaFree(c->sc_z3); cudaFree(c->sc_rgb); cudaFree(c->sc_chunks); cudaFree(c->sc_bounds);
    c->sc_chunks = nullptr;
    c->sc_bounds = nullptr;
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->walk_tri[i]); cudaFree(c->walk_base[i]); cudaFree(c->irr_tri[i]); cudaFree(c->tile_info[i]); cudaFree(c->tile_setup[i]);
        c->walk_tri[i] = c->irr_tri[i] = nullptr; c->walk_base[i] = nullptr; c->tile_info[i] = nullptr; c->tile_setup[i] = nullptr;
    }
    c->sc_a = c->sc_b = nullptr; c->sc_z3 = nullptr; c->sc_rgb = nullptr;
    c->have_scene = false;
    c->n_tri = 0;
    const size_t n = n_tri ? n_tri : 1;
    CU(cudaMalloc(&c->sc_a, n * sizeof(float4)));
    CU(cudaMalloc(&c->sc_b, n * sizeof(float4)));
    CU(cudaMalloc(&c->sc_z3, n * sizeof(float)));
    CU(cudaMalloc(&c->sc_rgb, n * sizeof(uint32_t)));
    const size_t n_padded = (n + 31) & ~(size_t)31;
    if (c->tma_feed) CU(cudaMalloc(&c->sc_chunks, n_padded / 32 * CHUNK_BYTES));   // second copy only for the TMA feed
    CU(cudaMalloc(&c->sc_bounds, (n_padded / 32 + 1) * sizeof(float4)));
    for (int i = 0; i < 2; ++i) {
        CU(cudaMalloc(&c->walk_tri[i], n * sizeof(uint32_t)));
        CU(cudaMalloc(&c->walk_base[i], n * sizeof(unsigned long long)));
        CU(cudaMalloc(&c->irr_tri[i], n * sizeof(uint32_t)));
        CU(cudaMalloc(&c->tile_info[i], n * sizeof(unsigned long long)));
        CU(cudaMalloc(&c->tile_setup[i], n * sizeof(TileTri)));
    }
    return SLOTH_OK;
}

// After sc_a/sc_b/sc_z3/sc_rgb have been filled on c->stream: derived copies, then wait.
int finish_scene(sloth_ctx* c, size_t n_tri)
{
    if (c->tma_feed && n_tri) {
        const size_t n_padded = (n_tri + 31) & ~(size_t)31;
        k_pack_chunks<<<(unsigned)((n_padded + 255) / 256), 256, 0, c->stream>>>(c->sc_a, c->sc_b, c->sc_z3, (uint32_t)n_tri, (uint32_t)n_padded, c->sc_chunks);
        c->launches += 1;
    }
    c->scene_absmax = 0.0f;
    if (n_tri) {
        uint32_t* d_absmax = reinterpret_cast<uint32_t*>(c->sc_bounds + ((n_tri + 31) / 32));   // spare slot after the last sphere
        CU(cudaMemsetAsync(d_absmax, 0, sizeof(uint32_t), c->stream));
        const unsigned n_chunks = (unsigned)((n_tri + 31) / 32);
        k_chunk_bounds<<<(n_chunks + 127) / 128, 128, 0, c->stream>>>(c->sc_a, c->sc_b, c->sc_z3, (uint32_t)n_tri, c->sc_bounds, d_absmax);
        c->launches += 1;
        CU(cudaMemcpyAsync(&c->scene_absmax, d_absmax, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return build_index(c, n_tri);
}

}  // namespace

#include "loader_api.inl"

extern "C" {

// ---- span wire format: device side in spans.cuh, host side in host/wire.hpp -------------------------------------
struct WireSync {   // shared by the thread that drives the GPU and the pool's expansion jobs
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> free_slots;
    int cuda_error = 0;   // first CUDA error an expansion job ran into
};

namespace {

void free_wire_buffers(sloth_ctx* c)
{
    for (int i = 0; i < 2; ++i) { cudaFree(c->d_runs[i]); c->d_runs[i] = nullptr; }
    for (int i = 0; i < sloth_ctx::WIRE_SLOTS; ++i) {
        if (c->h_runs[i]) cudaFreeHost(c->h_runs[i]);
        c->h_runs[i] = nullptr;
    }
    c->runs_cap = 0;
}

// Buffers of the span path for the current frame size: the run lists hold at most cells / 4 runs (half the bytes of
// the plain cells) -- frames with more runs go out as plain cells.
int ensure_wire_buffers(sloth_ctx* c)
{
    int rc = ensure_text_buffers(c, 0);   // block sums / offsets, the per-frame totals and their events
    if (rc) return rc;
    const size_t cap = std::max<size_t>(c->cells_per_frame / 4, 16);
    if (!c->wire_pool) {
        c->wire_pool = new sloth::WirePool(sloth::wire_default_threads());
        c->wire_sync = new WireSync;
        for (int i = 0; i < sloth_ctx::WIRE_SLOTS; ++i) CU(cudaEventCreateWithFlags(&c->ev_staged[i], cudaEventDisableTiming));
    }
    if (cap != c->runs_cap) {
        CU(cudaStreamSynchronize(c->resolve_stream));
        CU(cudaStreamSynchronize(c->copy_stream));
        c->wire_pool->wait_idle();
        free_wire_buffers(c);
        for (int i = 0; i < 2; ++i) CU(cudaMalloc(&c->d_runs[i], cap * sizeof(uint2)));
        for (int i = 0; i < sloth_ctx::WIRE_SLOTS; ++i) CU(cudaHostAlloc(&c->h_runs[i], cap * sizeof(uint2), cudaHostAllocDefault));
        c->runs_cap = cap;
        std::lock_guard<std::mutex> lk(c->wire_sync->mu);
        c->wire_sync->free_slots.clear();
        for (int i = 0; i < sloth_ctx::WIRE_SLOTS; ++i) c->wire_sync->free_slots.push_back(i);
    }
    return SLOTH_OK;
}

// run list of the frame in `d_cells` -> d_runs, its length -> *d_total (all on `st`)
int enqueue_spans(sloth_ctx* c, const uint32_t* d_cells, size_t n_cells, uint2* d_runs, unsigned long long* d_total, cudaStream_t st)
{
    const uint32_t nb = (uint32_t)((n_cells + FLUSH_CELLS_PER_BLOCK - 1) / FLUSH_CELLS_PER_BLOCK);
    k_span_count<<<nb, FLUSH_THREADS, 0, st>>>(d_cells, (uint32_t)n_cells, c->flush_block_sum);
    k_flush_scan<<<1, 1024, 0, st>>>(c->flush_block_sum, nb, c->flush_block_off, d_total);
    k_span_write<<<nb, FLUSH_THREADS, 0, st>>>(d_cells, (uint32_t)n_cells, c->flush_block_off, d_total, (unsigned long long)c->runs_cap, d_runs);
    c->launches += 3;
    CU(cudaGetLastError());
    return SLOTH_OK;
}

// sloth_render_batch over the span wire format: per frame the device sends the run list (or, for a frame without
// long runs, the plain cells); pool threads rebuild the caller's 4-byte cells while the next frames render.
int render_batch_spans(sloth_ctx* c, const float* rots, size_t n_frames, uint32_t* cells_out)
{
    int rc = ensure_wire_buffers(c);
    if (rc) return rc;
    const size_t cpf = c->cells_per_frame;
    WireSync* ws = c->wire_sync;
    sloth::WirePool* pool = c->wire_pool;
    const unsigned pieces = std::max(1u, std::min(pool->size(), 8u));
    const int device = c->device;
    // frame j's run list is complete on the device once ev_text[j&1] fires: read its length, start its copy, queue the expansion
    auto finish = [&](size_t j) -> int {
        const int b = (int)(j & 1);
        CU(cudaEventSynchronize(c->ev_text[b]));
        const size_t n_runs = (size_t)c->h_text_total[b];
        uint32_t* dest = cells_out + j * cpf;
        c->wire_frames += 1;
        if (n_runs > c->runs_cap) {   // no runs worth sending: the plain cells
            CU(cudaMemcpyAsync(dest, c->d_cells[b], cpf * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->copy_stream));
            CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
            c->wire_plain_frames += 1;
            c->wire_d2h_bytes += cpf * sizeof(uint32_t);
            return SLOTH_OK;
        }
        int slot;
        {
            std::unique_lock<std::mutex> lk(ws->mu);
            ws->cv.wait(lk, [&] { return !ws->free_slots.empty(); });
            slot = ws->free_slots.back();
            ws->free_slots.pop_back();
        }
        const sloth::Run* runs = reinterpret_cast<const sloth::Run*>(c->h_runs[slot]);
        CU(cudaMemcpyAsync(c->h_runs[slot], c->d_runs[b], n_runs * sizeof(uint2), cudaMemcpyDeviceToHost, c->copy_stream));
        CU(cudaEventRecord(c->ev_staged[slot], c->copy_stream));
        CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
        c->wire_d2h_bytes += n_runs * sizeof(uint2) + sizeof(unsigned long long);
        cudaEvent_t staged = c->ev_staged[slot];
        auto left = std::make_shared<std::atomic<unsigned>>(pieces);
        for (unsigned piece = 0; piece < pieces; ++piece) {
            const size_t c0 = cpf * piece / pieces, c1 = cpf * (piece + 1) / pieces;
            pool->submit([=] {
                cudaSetDevice(device);
                const cudaError_t e = cudaEventSynchronize(staged);
                if (e == cudaSuccess) sloth::expand_runs(runs, n_runs, c0, c1, dest, cpf);
                const bool last = left->fetch_sub(1) == 1;
                if (last || e != cudaSuccess) {
                    std::lock_guard<std::mutex> lk(ws->mu);
                    if (e != cudaSuccess && !ws->cuda_error) ws->cuda_error = (int)e;
                    if (last) ws->free_slots.push_back(slot);
                }
                if (last) ws->cv.notify_all();
            });
        }
        return SLOTH_OK;
    };
    CU(cudaEventRecord(c->ev[EV_START], c->stream));
    rc = enqueue_overlapped(
        c, rots, n_frames,
        [&](size_t k) -> int {   // cell and run buffers k&1 must have been copied out (frame k-2)
            if (k >= 2) CU(cudaStreamWaitEvent(c->resolve_stream, c->ev_copied[k & 1], 0));
            return SLOTH_OK;
        },
        [&](size_t k) { return c->d_cells[k & 1]; },
        [&](size_t k) -> int {
            const int b = (int)(k & 1);
            int r = enqueue_spans(c, c->d_cells[b], cpf, c->d_runs[b], c->d_text_total + b, c->resolve_stream);
            if (r) return r;
            CU(cudaMemcpyAsync(c->h_text_total + b, c->d_text_total + b, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                               c->resolve_stream));
            CU(cudaEventRecord(c->ev_text[b], c->resolve_stream));
            return k >= 1 ? finish(k - 1) : SLOTH_OK;
        });
    if (!rc) rc = finish(n_frames - 1);
    if (!rc) {
        cudaError_t e = cudaEventRecord(c->ev[EV_END], c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->resolve_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->copy_stream);
        if (e != cudaSuccess) rc = fail(SLOTH_E_CUDA, "span batch failed: %s", cudaGetErrorString(e));
    }
    pool->wait_idle();   // every frame of this call is in the caller's buffer (also on the error paths: jobs hold pointers into it)
    if (rc) return rc;
    if (ws->cuda_error) {
        const int e = ws->cuda_error;
        ws->cuda_error = 0;
        return fail(SLOTH_E_CUDA, "span expansion failed: %s", cudaGetErrorString((cudaError_t)e));
    }
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, c->ev[EV_START], c->ev[EV_END]));
    c->batch_ms_per_frame = ms / (float)n_frames;
    c->last_was_batch = true;
    c->ev_valid = true;
    c->ev_kernels_valid = false;
    return SLOTH_OK;
}

}  // namespace

const char* sloth_last_error(void) { return g_err; }

int sloth_ctx_create(int device, int image_mode, sloth_ctx** out)
{
    if (!out) return fail(SLOTH_E_ARG, "out is null");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(SLOTH_E_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(SLOTH_E_ARG, "device %d out of range [0,%d)", device, n_dev);
    CU(cudaSetDevice(device));
    sloth_ctx* c = new (std::nothrow) sloth_ctx();
    if (!c) return fail(SLOTH_E_ARG, "out of host memory");
    c->device = device;
    c->image = image_mode != 0;
    std::memcpy(c->thr, k_default_thr, sizeof c->thr);
    std::memcpy(c->glyph, k_default_glyph, sizeof c->glyph);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->l2_persist_max = (size_t)std::max(0, prop.persistingL2CacheMaxSize);
    c->l2_window_max = (size_t)std::max(0, prop.accessPolicyMaxWindowSize);
    if (const char* g = std::getenv("SLOTH_DEBUG")) c->debug = (uint32_t)std::atoi(g);
    if (const char* g = std::getenv("SLOTH_TMA")) c->tma_feed = std::atoi(g) != 0;
    if (const char* g = std::getenv("SLOTH_CARVEOUT")) c->carveout_override = std::atoi(g);   // profiling knob
    if (const char* g = std::getenv("SLOTH_TAIL")) c->tail_blocks_per_sm = (uint32_t)std::max(1, std::atoi(g));
    if (const char* g = std::getenv("SLOTH_BATCH")) c->batch_max = (uint32_t)std::min(16, std::max(1, std::atoi(g)));
    if (const char* g = std::getenv("SLOTH_GRID")) c->geom_blocks_per_sm = (uint32_t)std::max(1, std::atoi(g));
    if (const char* g = std::getenv("SLOTH_TRI2")) c->tri_pairs = std::atoi(g) != 0;
    if (const char* g = std::getenv("SLOTH_CONE")) c->cone = std::atoi(g) != 0;
    if (const char* g = std::getenv("SLOTH_TGRID")) c->tri_blocks_per_sm = (uint32_t)std::min((int)T_BLOCKS_PER_SM, std::max(1, std::atoi(g)));
    if (const char* g = std::getenv("SLOTH_L2PERSIST")) c->l2_persist = std::atoi(g) != 0;
    if (const char* g = std::getenv("SLOTH_TILES")) { c->tile_path = std::atoi(g) != 0; c->tile_always = std::atoi(g) == 2; }
    if (const char* g = std::getenv("SLOTH_PF")) c->pf_chunks = (uint32_t)std::min(64, std::max(0, std::atoi(g)));
    if (const char* g = std::getenv("SLOTH_PATH")) c->path_pref = std::min(2, std::max(0, std::atoi(g)));
    {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        // The geometry kernel of frame k+1 must reach the SMs the moment frame k's ends: it is the critical path of a
        // batch, the other streams' kernels (k_tail / resolve of frame k, k_xform of frame k+2) fill in beside its
        // persistent blocks.  Measured on the 10 M-triangle frame: geometry highest / others lowest 142 us per frame,
        // all lowest 154 us, resolve highest (round 1's choice for the soup kernel) 168 us.
        const char* pr = std::getenv("SLOTH_GEOM_PRIO");   // profiling knob: 0 = lowest priority instead
        CU(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, (pr && std::atoi(pr) == 0) ? lo : hi));
    }
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const char* pr = std::getenv("SLOTH_RESOLVE_PRIO");   // profiling knob: 1 = highest priority instead of lowest
        CU(cudaStreamCreateWithPriority(&c->resolve_stream, cudaStreamNonBlocking, (pr && std::atoi(pr) == 1) ? hi : lo));
    }
    {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const char* pr = std::getenv("SLOTH_XFORM_PRIO");   // profiling knob: 1 = highest priority for the k_xform stream
        CU(cudaStreamCreateWithPriority(&c->xform_stream, cudaStreamNonBlocking, (pr && std::atoi(pr) == 1) ? hi : lo));
    }
    CU(cudaEventCreateWithFlags(&c->ev_batch_start, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&c->ev_xform[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&c->ev_stamped[i], cudaEventDisableTiming));
    for (int i = 0; i < EV_N; ++i) CU(cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&c->ev_rendered[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
        // A timing event on purpose.  Recording it costs the geometry stream ~2 us after k_geom3(k), which is
        // what lets k_tail(k) / resolve(k) reach the SMs ahead of k_geom3(k+1); with a cudaEventDisableTiming
        // event about half of the B200s measured ran the overlapped batch at 185 us/frame instead of 162
        // (profiles/README.md, "overlap variance").
        CU(cudaEventCreateWithFlags(&c->ev_geom_done[i], cudaEventDefault));
        CU(cudaEventCreateWithFlags(&c->ev_resolved[i], cudaEventDisableTiming));
    }
    {   // dynamic shared memory of the soup kernel (TMA rings + per-block row stamps) can exceed the default limit
        const cudaFuncAttribute a = cudaFuncAttributeMaxDynamicSharedMemorySize;
        const int lim = 64 * 1024;
        CU(cudaFuncSetAttribute(k_geom3<false, false, false>, a, lim));
        CU(cudaFuncSetAttribute(k_geom3<false, true, false>, a, lim));
        CU(cudaFuncSetAttribute(k_geom3<true, false, false>, a, lim));
        CU(cudaFuncSetAttribute(k_geom3<true, true, false>, a, lim));
        CU(cudaFuncSetAttribute(k_geom3<false, false, true>, a, lim));
        CU(cudaFuncSetAttribute(k_geom3<false, true, true>, a, lim));
        CU(cudaFuncSetAttribute(k_geom3<true, false, true>, a, lim));
        CU(cudaFuncSetAttribute(k_geom3<true, true, true>, a, lim));
        // k_tri: cp.async rings + fragment rings (6.4 KB per warp) + up to 33 KB of row stamps
        const int tlim = (int)(sizeof(TWarpSmem) * T_WARPS) + 40 * 1024;
        CU(cudaFuncSetAttribute(k_tri<false, false, true>, a, tlim));
        CU(cudaFuncSetAttribute(k_tri<false, true, true>, a, tlim));
        CU(cudaFuncSetAttribute(k_tri<true, false, true>, a, tlim));
        CU(cudaFuncSetAttribute(k_tri<true, true, true>, a, tlim));
        CU(cudaFuncSetAttribute(k_tri<false, false, false>, a, tlim));
        CU(cudaFuncSetAttribute(k_tri<false, true, false>, a, tlim));
        CU(cudaFuncSetAttribute(k_tri<true, false, false>, a, tlim));
        CU(cudaFuncSetAttribute(k_tri<true, true, false>, a, tlim));
        CU(cudaFuncSetAttribute(k_tri<false, false, true, true>, a, tlim));
        CU(cudaFuncSetAttribute(k_tri<false, false, false, true>, a, tlim));
        const int tlim2 = (int)(sizeof(TWarpSmem2) * T_WARPS) + 40 * 1024;
        CU(cudaFuncSetAttribute(k_tri2<false, true>, a, tlim2));
        CU(cudaFuncSetAttribute(k_tri2<true, true>, a, tlim2));
        CU(cudaFuncSetAttribute(k_tri2<false, false>, a, tlim2));
        CU(cudaFuncSetAttribute(k_tri2<true, false>, a, tlim2));
    }
    *out = c;
    return SLOTH_OK;
}

int sloth_ctx_destroy(sloth_ctx* c)
{
    if (!c) return SLOTH_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->resolve_stream);
    free_frame_state(c);
    if (c->wire_pool) {
        c->wire_pool->wait_idle();
        delete c->wire_pool;
        delete c->wire_sync;
        for (int i = 0; i < sloth_ctx::WIRE_SLOTS; ++i) if (c->ev_staged[i]) cudaEventDestroy(c->ev_staged[i]);
    }
    free_wire_buffers(c);
    for (int i = 0; i < 2; ++i) { cudaFree(c->d_text[i]); if (c->ev_text[i]) cudaEventDestroy(c->ev_text[i]); }
    cudaFree(c->flush_block_sum); cudaFree(c->flush_block_off); cudaFree(c->d_text_total);
    if (c->h_text_total) cudaFreeHost(c->h_text_total);
    cudaFree(c->sc_a);
    cudaFree(c->sc_b);
    cudaFree(c->sc_z3);
    cudaFree(c->sc_rgb);
    cudaFree(c->sc_chunks);
    cudaFree(c->sc_bounds);
    free_index(c);
    for (int i = 0; i < 2; ++i) { cudaFree(c->walk_tri[i]); cudaFree(c->walk_base[i]); cudaFree(c->irr_tri[i]); cudaFree(c->tile_info[i]); cudaFree(c->tile_setup[i]); }
    for (int i = 0; i < EV_N; ++i) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(c->ev_rendered[i]);
        cudaEventDestroy(c->ev_copied[i]);
        cudaEventDestroy(c->ev_geom_done[i]);
        cudaEventDestroy(c->ev_resolved[i]);
    }
    if (c->loader) {
        c->loader->clear();
        delete c->loader;
    }
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->resolve_stream);
    if (c->xform_stream) cudaStreamDestroy(c->xform_stream);
    if (c->ev_batch_start) cudaEventDestroy(c->ev_batch_start);
    for (int i = 0; i < 2; ++i) if (c->ev_xform[i]) cudaEventDestroy(c->ev_xform[i]);
    for (int i = 0; i < 2; ++i) if (c->ev_stamped[i]) cudaEventDestroy(c->ev_stamped[i]);
    delete c;
    return SLOTH_OK;
}

int sloth_scene_set(sloth_ctx* c, const float* xyz, const uint8_t* rgb, size_t n_tri, float scene_max)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (n_tri && (!xyz || !rgb)) return fail(SLOTH_E_ARG, "xyz/rgb is null");
    if (n_tri > MAX_TRIS) return fail(SLOTH_E_TOO_LARGE, "%zu triangles; the depth key holds a 27-bit index (max %u)", n_tri, MAX_TRIS);
    CU(cudaSetDevice(c->device));
    int rc = alloc_scene(c, n_tri);
    if (rc) return rc;
    if (n_tri) {
        float* d_xyz = nullptr;
        uint8_t* d_rgb = nullptr;
        CU(cudaMalloc(&d_xyz, n_tri * 9 * sizeof(float)));
        CU(cudaMalloc(&d_rgb, n_tri * 3));
        CU(cudaMemcpyAsync(d_xyz, xyz, n_tri * 9 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_rgb, rgb, n_tri * 3, cudaMemcpyHostToDevice, c->stream));
        k_pack_scene<<<(unsigned)((n_tri + 255) / 256), 256, 0, c->stream>>>(d_xyz, d_rgb, (uint32_t)n_tri, c->sc_a, c->sc_b, c->sc_z3, c->sc_rgb);
        c->launches += 1;
        rc = finish_scene(c, n_tri);
        cudaFree(d_xyz);
        cudaFree(d_rgb);
        if (rc) return rc;
    }
    {   // finite and |v| <= 2^20 everywhere?  (NaN fails the comparison)
        bool clean = true;
        for (size_t i = 0; i < n_tri * 9 && clean; ++i) clean = std::fabs(xyz[i]) <= 1048576.0f;
        c->scene_clean = clean;
    }
    c->n_tri = (uint32_t)n_tri;
    c->scene_max = scene_max;
    c->have_scene = true;
    return SLOTH_OK;
}

int sloth_scene_set_indexed(sloth_ctx* c, const float* positions, size_t n_vert, const uint32_t* indices, const uint8_t* rgb,
                            size_t n_tri, float scene_max)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (n_tri && (!positions || !indices || !rgb || !n_vert)) return fail(SLOTH_E_ARG, "positions/indices/rgb is null or n_vert is 0");
    if (n_tri > MAX_TRIS) return fail(SLOTH_E_TOO_LARGE, "%zu triangles; the depth key holds a 27-bit index (max %u)", n_tri, MAX_TRIS);
    if (n_vert >= 0xFFFFFFFFull) return fail(SLOTH_E_TOO_LARGE, "%zu vertices; vertex ids are 32-bit", n_vert);
    CU(cudaSetDevice(c->device));
    int rc = alloc_scene(c, n_tri);
    if (rc) return rc;
    if (n_tri) {
        float* d_pos = nullptr;
        uint32_t* d_idx = nullptr;
        uint8_t* d_rgb = nullptr;
        uint32_t* d_bad = nullptr;
        auto drop = [&]() { cudaFree(d_pos); cudaFree(d_idx); cudaFree(d_rgb); cudaFree(d_bad); };
        cudaError_t e = cudaMalloc(&d_pos, n_vert * 3 * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&d_idx, n_tri * 3 * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&d_rgb, n_tri * 3);
        if (e == cudaSuccess) e = cudaMalloc(&d_bad, sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, sizeof(uint32_t), c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_pos, positions, n_vert * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, indices, n_tri * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_rgb, rgb, n_tri * 3, cudaMemcpyHostToDevice, c->stream);
        uint32_t bad = 0;
        if (e == cudaSuccess) {
            ix::k_ix_expand_input<<<(unsigned)((n_tri + 255) / 256), 256, 0, c->stream>>>(d_pos, (uint32_t)n_vert, d_idx, d_rgb, (uint32_t)n_tri,
                                                                                        c->sc_a, c->sc_b, c->sc_z3, c->sc_rgb, d_bad);
            c->launches += 1;
            e = cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, c->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        drop();
        if (e != cudaSuccess) return fail(SLOTH_E_CUDA, "sloth_scene_set_indexed: %s", cudaGetErrorString(e));
        if (bad) return fail(SLOTH_E_ARG, "%u triangles reference a vertex id >= n_vert (%zu)", bad, n_vert);
        rc = finish_scene(c, n_tri);
        if (rc) return rc;
    }
    {   // finite and |v| <= 2^20 everywhere?  Unreferenced vertices only make the answer conservative.
        bool clean = true;
        for (size_t i = 0; i < n_vert * 3 && clean; ++i) clean = std::fabs(positions[i]) <= 1048576.0f;
        c->scene_clean = clean || n_tri == 0;
    }
    c->n_tri = (uint32_t)n_tri;
    c->scene_max = scene_max;
    c->have_scene = true;
    return SLOTH_OK;
}

int sloth_ctx_set_path(sloth_ctx* c, int path)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (path != SLOTH_PATH_AUTO && path != SLOTH_PATH_SOUP && path != SLOTH_PATH_INDEXED)
        return fail(SLOTH_E_ARG, "path must be SLOTH_PATH_AUTO, _SOUP or _INDEXED");
    c->path_pref = path;   // takes effect at the next sloth_scene_set / _set_indexed / loader commit
    return SLOTH_OK;
}

int sloth_ctx_resize(sloth_ctx* c, uint32_t W, uint32_t H)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (W == 0 || H == 0) return fail(SLOTH_E_ARG, "width and height must be >= 1 (got %ux%u)", W, H);
    if (W > 65535u || H > 65535u)
        return fail(SLOTH_E_TOO_LARGE, "width/height above 65535 (the reference takes the size `as u16`, context.rs:99)");
    if ((unsigned long long)W * H + H >= (1ull << 31)) return fail(SLOTH_E_TOO_LARGE, "W*H+H must stay below 2^31");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->resolve_stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    c->W = W;
    c->H = H;
    c->row0 = c->row1 = 0;
    return alloc_frame_state(c);
}

int sloth_ctx_set_band(sloth_ctx* c, uint32_t row0, uint32_t row1)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (c->W == 0) return fail(SLOTH_E_STATE, "sloth_ctx_resize has not been called");
    if (!(row0 == 0 && row1 == 0) && (row0 >= row1 || row1 > c->H))
        return fail(SLOTH_E_ARG, "band [%u,%u) is not inside [0,%u)", row0, row1, c->H);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->resolve_stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    c->row0 = row0;
    c->row1 = row1;
    return alloc_frame_state(c);
}

size_t sloth_cells_per_frame(const sloth_ctx* c) { return c ? c->cells_per_frame : 0; }

int sloth_render(sloth_ctx* c, const float rot[16], uint32_t* cells_out, float* z_out)
{
    int rc = check_ready(c);
    if (rc) return rc;
    if (!rot || !cells_out) return fail(SLOTH_E_ARG, "rot/cells_out is null");
    if (z_out && c->row1 != 0) return fail(SLOTH_E_ARG, "z_out is not available in band mode");
    if (!z_out && c->wire == SLOTH_WIRE_SPANS) {   // one frame, its cells rebuilt by all pool threads
        rc = render_batch_spans(c, rot, 1, cells_out);
        c->last_was_batch = false;
        return rc;
    }
    if (z_out && !c->d_z) CU(cudaMalloc(&c->d_z, (size_t)c->W * c->H * sizeof(float)));
    rc = enqueue_frame(c, rot, c->d_cells[0], z_out ? c->d_z : nullptr, true);
    if (rc) return rc;
    c->ev_valid = true;
    c->ev_kernels_valid = (c->stat_flags & 2u) != 0;
    c->last_was_batch = false;
    CU(cudaMemcpyAsync(cells_out, c->d_cells[0], c->cells_per_frame * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (z_out)
        CU(cudaMemcpyAsync(z_out, c->d_z, (size_t)c->W * c->H * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SLOTH_OK;
}

int sloth_render_device(sloth_ctx* c, const float rot[16], void* d_cells)
{
    int rc = check_ready(c);
    if (rc) return rc;
    if (!rot || !d_cells) return fail(SLOTH_E_ARG, "rot/d_cells is null");
    if (((uintptr_t)d_cells & 15u) != 0) return fail(SLOTH_E_ARG, "d_cells must be 16-byte aligned");
    rc = enqueue_frame(c, rot, static_cast<uint32_t*>(d_cells), nullptr, true);
    if (rc) return rc;
    c->ev_valid = true;
    c->ev_kernels_valid = (c->stat_flags & 2u) != 0;
    c->last_was_batch = false;
    return SLOTH_OK;
}

void* sloth_ctx_stream(sloth_ctx* c) { return c ? (void*)c->stream : nullptr; }

int sloth_device_alloc(int device, size_t bytes, void** d_ptr_out)
{
    if (!d_ptr_out || !bytes) return fail(SLOTH_E_ARG, "d_ptr_out is null or bytes is 0");
    CU(cudaSetDevice(device));
    CU(cudaMalloc(d_ptr_out, bytes));   // cudaMalloc, not a pool: only such allocations can be exported
    return SLOTH_OK;
}

int sloth_device_free(int device, void* d_ptr)
{
    CU(cudaSetDevice(device));
    CU(cudaFree(d_ptr));
    return SLOTH_OK;
}

int sloth_ipc_export(int device, const void* d_ptr, unsigned char handle_out[SLOTH_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == SLOTH_IPC_HANDLE_BYTES, "handle size");
    if (!d_ptr || !handle_out) return fail(SLOTH_E_ARG, "d_ptr/handle_out is null");
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    std::memcpy(handle_out, &h, sizeof h);
    return SLOTH_OK;
}

int sloth_ipc_open(int device, const unsigned char handle[SLOTH_IPC_HANDLE_BYTES], void** d_ptr_out)
{
    if (!handle || !d_ptr_out) return fail(SLOTH_E_ARG, "handle/d_ptr_out is null");
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    CU(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return SLOTH_OK;
}

int sloth_ipc_close(int device, void* d_ptr)
{
    CU(cudaSetDevice(device));
    CU(cudaIpcCloseMemHandle(d_ptr));
    return SLOTH_OK;
}

int sloth_device_write(sloth_ctx* c, void* d_ptr, const void* host_in, size_t bytes)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (bytes && (!d_ptr || !host_in)) return fail(SLOTH_E_ARG, "d_ptr/host_in is null");
    CU(cudaSetDevice(c->device));
    if (bytes) CU(cudaMemcpyAsync(d_ptr, host_in, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SLOTH_OK;
}

int sloth_device_read(sloth_ctx* c, const void* d_ptr, void* host_out, size_t bytes)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (bytes && (!d_ptr || !host_out)) return fail(SLOTH_E_ARG, "d_ptr/host_out is null");
    CU(cudaSetDevice(c->device));
    if (bytes) CU(cudaMemcpyAsync(host_out, d_ptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SLOTH_OK;
}

int sloth_ctx_sync(sloth_ctx* c)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->resolve_stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    return SLOTH_OK;
}


int sloth_ctx_set_wire(sloth_ctx* c, int wire)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (wire != SLOTH_WIRE_CELLS && wire != SLOTH_WIRE_SPANS) return fail(SLOTH_E_ARG, "wire must be SLOTH_WIRE_CELLS or SLOTH_WIRE_SPANS");
    c->wire = wire;
    c->wire_d2h_bytes = c->wire_frames = c->wire_plain_frames = 0;
    return SLOTH_OK;
}

int sloth_wire_stats(const sloth_ctx* c, uint64_t out[4])
{
    if (!c || !out) return fail(SLOTH_E_ARG, "null argument");
    out[0] = c->wire_frames;
    out[1] = c->wire_plain_frames;
    out[2] = c->wire_d2h_bytes;
    out[3] = c->wire_pool ? c->wire_pool->size() : 0;
    return SLOTH_OK;
}

int sloth_expand_spans(const uint32_t* runs, size_t n_runs, uint32_t* cells_out, size_t n_cells)
{
    if ((!runs && n_runs) || (!cells_out && n_cells)) return fail(SLOTH_E_ARG, "null argument");
    if (n_cells && (n_runs == 0 || runs[0] != 0)) return fail(SLOTH_E_ARG, "the first run must start at cell 0");
    for (size_t i = 1; i < n_runs; ++i)
        if (runs[2 * i] <= runs[2 * (i - 1)] || runs[2 * i] >= n_cells) return fail(SLOTH_E_ARG, "run %zu: starts must ascend inside the frame", i);
    sloth::expand_runs(reinterpret_cast<const sloth::Run*>(runs), n_runs, 0, n_cells, cells_out, n_cells);
    return SLOTH_OK;
}

int sloth_render_batch(sloth_ctx* c, const float* rots, size_t n_frames, uint32_t* cells_out)
{
    int rc = check_ready(c);
    if (rc) return rc;
    if (n_frames == 0) return SLOTH_OK;
    if (!rots || !cells_out) return fail(SLOTH_E_ARG, "rots/cells_out is null");
    if (c->wire == SLOTH_WIRE_SPANS) return render_batch_spans(c, rots, n_frames, cells_out);
    const size_t cpf = c->cells_per_frame;
    CU(cudaEventRecord(c->ev[EV_START], c->stream));
    rc = enqueue_overlapped(
        c, rots, n_frames,
        [&](size_t k) -> int {   // cell buffer k&1 must have been copied out (frame k-2)
            if (k >= 2) CU(cudaStreamWaitEvent(c->resolve_stream, c->ev_copied[k & 1], 0));
            return SLOTH_OK;
        },
        [&](size_t k) { return c->d_cells[k & 1]; },
        [&](size_t k) -> int {
            const int b = (int)(k & 1);
            CU(cudaStreamWaitEvent(c->copy_stream, c->ev_resolved[b], 0));
            CU(cudaMemcpyAsync(cells_out + k * cpf, c->d_cells[b], cpf * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->copy_stream));
            CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
            return SLOTH_OK;
        });
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[EV_END], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->resolve_stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, c->ev[EV_START], c->ev[EV_END]));
    c->batch_ms_per_frame = ms / (float)n_frames;
    c->last_was_batch = true;
    c->ev_valid = true;
    c->ev_kernels_valid = false;
    return SLOTH_OK;
}

int sloth_render_device_batch(sloth_ctx* c, const float* rots, size_t n_frames, void* d_cells, size_t frame_stride_cells)
{
    int rc = check_ready(c);
    if (rc) return rc;
    if (n_frames == 0) return SLOTH_OK;
    if (!rots || !d_cells) return fail(SLOTH_E_ARG, "rots/d_cells is null");
    if (((uintptr_t)d_cells & 15u) != 0 || (frame_stride_cells & 3u) != 0)
        return fail(SLOTH_E_ARG, "d_cells must be 16-byte aligned and the frame stride a multiple of 4 cells");
    uint32_t* base = static_cast<uint32_t*>(d_cells);
    CU(cudaEventRecord(c->ev[EV_START], c->stream));
    rc = enqueue_overlapped(
        c, rots, n_frames, [&](size_t) -> int { return SLOTH_OK; },
        [&](size_t k) { return base + k * frame_stride_cells; }, [&](size_t) -> int { return SLOTH_OK; });
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[EV_END], c->stream));   // the context stream has joined the last resolves
    c->batch_ms_per_frame = 0.0f;
    c->batch_frames = n_frames;
    c->last_was_batch = true;
    c->batch_pending = true;
    c->ev_valid = true;
    c->ev_kernels_valid = false;
    return SLOTH_OK;
}

size_t sloth_text_capacity(const sloth_ctx* c, int mode)
{
    if (!c || mode < 0 || mode > 2) return 0;
    return c->cells_per_frame * text_bytes_per_cell(mode);
}

int sloth_flush_device(sloth_ctx* c, int mode, const void* d_cells, size_t n_cells, void* d_text, size_t cap, size_t* len_out)
{
    if (!c || !d_cells || !d_text || !len_out) return fail(SLOTH_E_ARG, "null argument");
    if (mode < 0 || mode > 2) return fail(SLOTH_E_ARG, "mode must be 0 (plain), 1 (ANSI) or 2 (webify)");
    if (cap < n_cells * text_bytes_per_cell(mode)) return fail(SLOTH_E_ARG, "text buffer smaller than the worst case (%zu bytes)", n_cells * text_bytes_per_cell(mode));
    CU(cudaSetDevice(c->device));
    uint32_t* bs = nullptr;
    unsigned long long* bo = nullptr;
    unsigned long long* tot = nullptr;
    const size_t nb = (n_cells + FLUSH_CELLS_PER_BLOCK - 1) / FLUSH_CELLS_PER_BLOCK + 1;
    CU(cudaMalloc(&bs, nb * sizeof(uint32_t)));
    CU(cudaMalloc(&bo, nb * sizeof(unsigned long long)));
    CU(cudaMalloc(&tot, sizeof(unsigned long long)));
    int rc = enqueue_flush(c, mode, static_cast<const uint32_t*>(d_cells), n_cells, static_cast<char*>(d_text), tot, bs, bo, c->stream);
    unsigned long long h = 0;
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(&h, tot, sizeof h, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(SLOTH_E_CUDA, "flush failed: %s", cudaGetErrorString(e));
    }
    cudaFree(bs); cudaFree(bo); cudaFree(tot);
    *len_out = (size_t)h;
    return rc;
}

int sloth_render_text_batch(sloth_ctx* c, const float* rots, size_t n_frames, int mode, char* text_out, size_t frame_stride,
                            size_t* lens_out)
{
    int rc = check_ready(c);
    if (rc) return rc;
    if (n_frames == 0) return SLOTH_OK;
    if (!rots || !text_out || !lens_out) return fail(SLOTH_E_ARG, "null argument");
    if (mode < 0 || mode > 2) return fail(SLOTH_E_ARG, "mode must be 0 (plain), 1 (ANSI) or 2 (webify)");
    if (frame_stride < sloth_text_capacity(c, mode)) return fail(SLOTH_E_ARG, "frame_stride below sloth_text_capacity()");
    rc = ensure_text_buffers(c, mode);
    if (rc) return rc;
    const size_t cpf = c->cells_per_frame;
    // frame j's text is complete on the device once ev_text[j&1] fires: read its length, start its copy
    auto finish = [&](size_t j) -> int {
        const int b = (int)(j & 1);
        CU(cudaEventSynchronize(c->ev_text[b]));
        const size_t len = (size_t)c->h_text_total[b];
        lens_out[j] = len;
        CU(cudaMemcpyAsync(text_out + j * frame_stride, c->d_text[b], len, cudaMemcpyDeviceToHost, c->copy_stream));
        CU(cudaEventRecord(c->ev_copied[b], c->copy_stream));
        return SLOTH_OK;
    };
    rc = enqueue_overlapped(
        c, rots, n_frames,
        [&](size_t k) -> int {   // buffers k&1 (cells and text) must have been copied out (frame k-2)
            if (k >= 2) CU(cudaStreamWaitEvent(c->resolve_stream, c->ev_copied[k & 1], 0));
            return SLOTH_OK;
        },
        [&](size_t k) { return c->d_cells[k & 1]; },
        [&](size_t k) -> int {
            const int b = (int)(k & 1);
            int r = enqueue_flush(c, mode, c->d_cells[b], cpf, c->d_text[b], c->d_text_total + b, c->flush_block_sum,
                                  c->flush_block_off, c->resolve_stream);
            if (r) return r;
            CU(cudaMemcpyAsync(c->h_text_total + b, c->d_text_total + b, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                               c->resolve_stream));
            CU(cudaEventRecord(c->ev_text[b], c->resolve_stream));
            return k >= 1 ? finish(k - 1) : SLOTH_OK;
        });
    if (rc) return rc;
    rc = finish(n_frames - 1);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->resolve_stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    c->last_was_batch = true;
    c->ev_valid = false;
    return SLOTH_OK;
}

int sloth_render_text(sloth_ctx* c, const float rot[16], int mode, char* text_out, size_t cap, size_t* len_out)
{
    if (!len_out) return fail(SLOTH_E_ARG, "len_out is null");
    return sloth_render_text_batch(c, rot, 1, mode, text_out, cap, len_out);
}

int sloth_shader_set(sloth_ctx* c, const float thr[9], const char glyph[10])
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (!thr || !glyph) {
        std::memcpy(c->thr, k_default_thr, sizeof c->thr);
        std::memcpy(c->glyph, k_default_glyph, sizeof c->glyph);
        return SLOTH_OK;
    }
    std::memcpy(c->thr, thr, sizeof c->thr);
    std::memcpy(c->glyph, glyph, 10);
    return SLOTH_OK;
}

int sloth_stats_enable(sloth_ctx* c, uint32_t flags)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    c->stat_flags = flags;
    return SLOTH_OK;
}

int sloth_stats_get(sloth_ctx* c, sloth_stats* out)
{
    if (!c || !out) return fail(SLOTH_E_ARG, "null argument");
    std::memset(out, 0, sizeof *out);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->resolve_stream));
    out->frames = c->frames;
    out->kernel_launches = c->launches;
    out->n_tri = c->n_tri;
    out->load_read_ms = c->load_ms[0];
    out->load_parse_ms = c->load_ms[1];
    out->load_commit_ms = c->load_ms[2];
    out->l2_window_bytes = (uint64_t)c->l2_window_bytes;
    out->l2_persist_max = (uint64_t)c->l2_persist_max;
    out->n_vert = c->indexed ? c->n_vert : 0u;
    out->geom_path = c->indexed ? (uint32_t)SLOTH_PATH_INDEXED : (uint32_t)SLOTH_PATH_SOUP;
    if (c->sized && c->aux_region[c->last_set]) {
        FrameAux aux;
        CU(cudaMemcpy(&aux, c->aux_region[c->last_set] + c->rowmax_bytes, sizeof aux, cudaMemcpyDeviceToHost));
        out->fragments = aux.frag_counter;
        out->walk_tris = (uint32_t)(aux.walk_counter >> ITEM_BITS);
        out->walk_items = (uint32_t)(aux.walk_counter & ITEM_MASK);
        out->irregular_tris = aux.irr_count;
        out->stamp_fixups = aux.stamp_exact;
        out->chunks_processed = aux.chunks_done;
        if (c->tile_region[c->last_set]) {
            TileAux ta;
            CU(cudaMemcpy(&ta, c->tile_region[c->last_set], sizeof ta, cudaMemcpyDeviceToHost));
            out->tile_tris = ta.binned_tris;
            out->tile_pairs = (uint32_t)std::min<unsigned long long>(ta.pairs_total, 0xFFFFFFFFull);
            out->tiles_used = ta.n_tiles_used;
        }
    }
    if (c->ev_valid) {
        if (c->last_was_batch && c->batch_pending) {
            float ms = 0.0f;
            CU(cudaEventElapsedTime(&ms, c->ev[EV_START], c->ev[EV_END]));
            c->batch_ms_per_frame = ms / (float)std::max<size_t>(c->batch_frames, 1);
            c->batch_pending = false;
        }
        if (c->last_was_batch) out->last_frame_ms = c->batch_ms_per_frame;
        else CU(cudaEventElapsedTime(&out->last_frame_ms, c->ev[EV_START], c->ev[EV_END]));
        if (c->ev_kernels_valid && !c->last_was_batch) {
            if (c->indexed && c->n_tri) {
                CU(cudaEventElapsedTime(&out->xform_ms, c->ev[EV_START], c->ev[EV_XFORM]));
                CU(cudaEventElapsedTime(&out->geom_ms, c->ev[EV_XFORM], c->ev[EV_GEOM]));
            } else {
                CU(cudaEventElapsedTime(&out->geom_ms, c->ev[EV_START], c->ev[EV_GEOM]));
            }
            CU(cudaEventElapsedTime(&out->walk_ms, c->ev[EV_GEOM], c->ev[EV_WALK]));
            CU(cudaEventElapsedTime(&out->resolve_ms, c->ev[EV_RESOLVE_BEGIN], c->ev[EV_END]));
        }
    }
    return SLOTH_OK;
}

int sloth_pinned_alloc(size_t bytes, void** out)
{
    if (!out) return fail(SLOTH_E_ARG, "out is null");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return SLOTH_OK;
}

int sloth_pinned_free(void* ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return SLOTH_OK;
}

int sloth_host_register(void* ptr, size_t bytes)
{
    if (!ptr || !bytes) return fail(SLOTH_E_ARG, "ptr is null or bytes is 0");
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return SLOTH_OK;
}

int sloth_host_unregister(void* ptr)
{
    if (ptr) CU(cudaHostUnregister(ptr));
    return SLOTH_OK;
}

/* ---- host-side helpers ---- */

void sloth_utransform(uint32_t W, uint32_t H, float scene_max, float out[16])
{
    // Context::blank starts from identity (context.rs:25-27); update() only
    // replaces it when the u16 size differs from (0,0) (context.rs:104).
    std::memset(out, 0, 16 * sizeof(float));
    at(out, 0, 0) = at(out, 1, 1) = at(out, 2, 2) = at(out, 3, 3) = 1.0f;
    const uint16_t w16 = (uint16_t)W, h16 = (uint16_t)H;
    if (w16 == 0 && h16 == 0) return;
    const float fw = (float)w16, fh = (float)h16;
    const float scale = std::fmin(fh, fw / 2.0f) / scene_max / 2.0f;  // context.rs:114
    at(out, 0, 0) = scale;
    at(out, 0, 3) = fw / 4.0f;
    at(out, 1, 1) = -scale;
    at(out, 1, 3) = fh / 2.0f;
    at(out, 2, 2) = scale;
}

void sloth_rotation_from_euler(float roll, float pitch, float yaw, float out[16])
{
    const float sr = sinf(roll), cr = cosf(roll);
    const float sp = sinf(pitch), cp = cosf(pitch);
    const float sy = sinf(yaw), cy = cosf(yaw);
    std::memset(out, 0, 16 * sizeof(float));
    at(out, 0, 0) = cy * cp;
    at(out, 0, 1) = cy * sp * sr - sy * cr;
    at(out, 0, 2) = cy * sp * cr + sy * sr;
    at(out, 1, 0) = sy * cp;
    at(out, 1, 1) = sy * sp * sr + cy * cr;
    at(out, 1, 2) = sy * sp * cr - cy * sr;
    at(out, 2, 0) = -sp;
    at(out, 2, 1) = cp * sr;
    at(out, 2, 2) = cp * cr;
    at(out, 3, 3) = 1.0f;
}

size_t sloth_turntable_pitches(float y_arg, uint32_t n_frames, float* out, size_t cap)
{
    const float pi = 3.14159265358979323846f;
    float pitch = y_arg;
    pitch += pi;                                           // inputs.rs:148
    const float step = (2.0f * pi) * (1.0f / (float)n_frames);  // main.rs:57
    size_t count = 0;
    long long frame_count = 0;
    for (;;) {
        if (out && count < cap) out[count] = pitch;
        ++count;
        pitch += step;                                     // main.rs:92-96
        if (pitch > 9.42477f || (long long)n_frames - 1 == frame_count) break;  // main.rs:99
        ++frame_count;
    }
    return count;
}

}  // extern "C"
