// raster_core.cuh -- device-side arithmetic of the sloth raster path (sm_100a).
//
// Every floating-point operation that decides coverage, depth or glyph is
// written with the round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/
// __fdiv_rn/__fsqrt_rn), which nvcc never contracts into FMAs, in exactly the
// operation order of the reference:
//   Triangle::mul           src/geometry.rs:43-48   (nalgebra gemv order)
//   Triangle::aabb          src/geometry.rs:37-42
//   bounds / 1/area         src/rasterizer.rs:58-67
//   orient                  src/rasterizer.rs:30-32
//   shade / z / id / test   src/rasterizer.rs:72-86
//   Triangle::normal        src/geometry.rs:49-56
//   default_shader          src/rasterizer.rs:5-27
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sloth {

#define SLOTH_DEV __device__ __forceinline__

static constexpr unsigned long long KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
static constexpr uint32_t MAX_TRIS = (1u << 27) - 1;  // 27-bit triangle index in the key
static constexpr float REGULAR_LIMIT = 1099511627776.0f;  // 2^40

// Per-frame constants, passed by value to every kernel.
struct FrameParams {
    float m[12];       // rows 0..2 of M = utransform * rot  (m[r*4 + c]); row 3 is (0,0,0,1)
    float thr[9];      // shader thresholds (`<=`)
    float wm1, hm1;    // (W-1) as f32, (H-1) as f32          rasterizer.rs:64-65
    uint32_t W, H;     // Context.width / height
    uint32_t KW;       // key slots per row: W/2 (W even) or W (W odd)
    uint32_t XS;       // key step per pixel x: 1 (W even) or 2 (W odd)
    uint32_t row0, row1;  // destination-row band owned by this context ([0,H) = whole frame)
    uint32_t krow0;    // first row of the key plane: row0, or row0-1 (halo) for odd W bands
    uint32_t srow0, srow1;  // rows whose newline stamp (cell y*W+1) lands in this context's cells: [row0,row1) --
                       // except for W == 1, where that cell is (row y+1, column 0): [row0-1, row1-1)
    uint32_t n_tri;
    uint32_t image;    // Context.image
    uint32_t count_frags;
    uint32_t debug;    // profiling experiments only (SLOTH_DEBUG): 1 = skip key atomics, 2 = skip row stamps, 8 = k_tri parks but
                       // never emits, 16 = k_tri stops every chunk after the back-face proof
    char glyph[12];    // 10 glyphs (+pad)
    // band contexts: whole chunks of 32 triangles are skipped when their bounding sphere (Scene::bounds) cannot
    // reach the band's rows.  cull_scale = |row 1 of M| (rounded up), cull_pad = bound on the rounding error of
    // the computed y' of any vertex or sphere centre (both set by the host per frame); 0 = no culling.
    uint32_t cull_on;
    float cull_scale, cull_pad;
    // indexed path (k_tri), bounded scenes only: 2^-18 * D_frame, D_frame >= the distance bound D of every
    // triangle's back-face proof (backface_proven), rounded up on the host
    float bf_k;
    uint32_t pf_chunks;   // k_tri: L2 prefetch distance of the record stream in chunks per warp (SLOTH_PF, 0 = off)
    // indexed path, bounded whole-frame scenes: the per-frame half of the super-chunk back-face certificate
    // (k_super_cert, tri_kernel.cuh).  cone_c = (row 0 of M) x (row 1 of M): the doubled screen area of a triangle is
    // cone_c . ((V1-V3) x (V2-V1)); cone_cnorm = |cone_c|, cone_s = max(|row 0|, |row 1|), cone_ev = bound on the
    // rounding error of any computed x' or y' (the last three rounded up).  cone_on = 0: no super-chunk is skipped.
    uint32_t cone_on;
    uint32_t cone_n_super;   // full super-chunks of the scene: chunks from 8 * cone_n_super on are not covered by the list
    float cone_c[3], cone_cnorm, cone_s, cone_ev;
};

// Resident scene: 40 B per triangle in four coalesced streams; the geometry kernels read the first
// three (36 B), only resolve reads the colours.
struct Scene {
    const float4* __restrict__ a;     // v1.x v1.y v1.z v2.x
    const float4* __restrict__ b;     // v2.y v2.z v3.x v3.y
    const float* __restrict__ z3;     // v3.z
    const uint32_t* __restrict__ rgb; // r | g<<8 | b<<16
    const float4* __restrict__ bounds; // per chunk of 32 triangles: bounding sphere (centre.xyz, radius), object space
    // indexed path (index.cuh): per-triangle vertex ids, and the per-frame output of k_xform
    const uint4* __restrict__ rec;     // (i0, i1, i2, triangle << 1 | connected) per triangle, padded to a multiple of 32 with the sentinel vertex
    const float2* __restrict__ vxy;    // (x', y') per unique vertex
    const float* __restrict__ vz;      // z' per unique vertex
    const uint32_t* __restrict__ super_ids;   // unique vertex ids of every super-chunk of SC_TRIS triangles (index.cuh), SC_IDS each
};

// Transformed triangle + everything hoistable out of the per-candidate loop.
struct Setup {
    float x1, y1, z1, x2, y2, z2, x3, y3, z3;
    float dx0, dy0, dx1, dy1, dx2, dy2;  // edge deltas: e0 = v2->v3, e1 = v3->v1, e2 = v1->v2
    float mx0;                           // aabb.max.x (for the tight column estimate)
    uint32_t minx, maxx, miny, maxy;     // the reference's scan domain (rasterizer.rs:59-66)
    bool regular;                        // all coordinates finite and |v| <= 2^40
};

SLOTH_DEV float mul(float a, float b) { return __fmul_rn(a, b); }
SLOTH_DEV float add(float a, float b) { return __fadd_rn(a, b); }
SLOTH_DEV float sub(float a, float b) { return __fsub_rn(a, b); }

// (M * v)[r] with v.w = 1: ((m0*x + m1*y) + m2*z) + m3*1   -- gemv/axcpy order
SLOTH_DEV float xform_row(const float* r, float x, float y, float z)
{
    return add(add(add(mul(r[0], x), mul(r[1], y)), mul(r[2], z)), r[3]);
}

SLOTH_DEV void load_tri(const Scene& sc, uint32_t t, float (&v)[9])
{
    const float4 A = __ldg(sc.a + t);
    const float4 B = __ldg(sc.b + t);
    const float C = __ldg(sc.z3 + t);
    v[0] = A.x; v[1] = A.y; v[2] = A.z;
    v[3] = A.w; v[4] = B.x; v[5] = B.y;
    v[6] = B.z; v[7] = B.w; v[8] = C;
}

SLOTH_DEV bool in_limit(float v) { return fabsf(v) <= REGULAR_LIMIT; }  // false for NaN

// Transform + bounds.  Bit-identical to draw_triangle's prologue.
SLOTH_DEV void setup_tri(const FrameParams& p, const float (&v)[9], Setup& s)
{
    s.x1 = xform_row(p.m + 0, v[0], v[1], v[2]);
    s.y1 = xform_row(p.m + 4, v[0], v[1], v[2]);
    s.z1 = xform_row(p.m + 8, v[0], v[1], v[2]);
    s.x2 = xform_row(p.m + 0, v[3], v[4], v[5]);
    s.y2 = xform_row(p.m + 4, v[3], v[4], v[5]);
    s.z2 = xform_row(p.m + 8, v[3], v[4], v[5]);
    s.x3 = xform_row(p.m + 0, v[6], v[7], v[8]);
    s.y3 = xform_row(p.m + 4, v[6], v[7], v[8]);
    s.z3 = xform_row(p.m + 8, v[6], v[7], v[8]);

    // aabb (f32::min/max == fminf/fmaxf: NaN-ignoring)
    const float mn0 = fminf(s.x1, fminf(s.x2, s.x3));
    const float mn1 = fminf(s.y1, fminf(s.y2, s.y3));
    const float mx0 = fmaxf(s.x1, fmaxf(s.x2, s.x3));
    const float mx1 = fmaxf(s.y1, fmaxf(s.y2, s.y3));
    s.mx0 = mx0;
    // `as usize` saturates and maps NaN to 0; cvt.rzi.u32.f32 does the same
    // (values above 2^32 clamp to 0xFFFFFFFF, which is still "past the end").
    s.minx = __float2uint_rz(ceilf(fmaxf(mn0, 1.0f)));
    s.miny = __float2uint_rz(ceilf(fmaxf(mn1, 1.0f)));
    s.maxx = __float2uint_rz(ceilf(fminf(mul(mx0, 2.0f), p.wm1)));
    s.maxy = __float2uint_rz(ceilf(fminf(mx1, p.hm1)));

    s.dx0 = sub(s.x3, s.x2); s.dy0 = sub(s.y3, s.y2);
    s.dx1 = sub(s.x1, s.x3); s.dy1 = sub(s.y1, s.y3);
    s.dx2 = sub(s.x2, s.x1); s.dy2 = sub(s.y2, s.y1);

    s.regular = in_limit(s.x1) && in_limit(s.y1) && in_limit(s.z1) && in_limit(s.x2) && in_limit(s.y2) &&
                in_limit(s.z2) && in_limit(s.x3) && in_limit(s.y3) && in_limit(s.z3);
}

// Per-triangle shading constants: a = 1/orient(v1,v2,v3), k = normal().z * a.
struct Shade {
    float a, k, dz1, dz2;
};

SLOTH_DEV void shade_setup(const Setup& s, Shade& sh)
{
    const float e1x = s.dx2, e1y = s.dy2, e1z = sub(s.z2, s.z1);  // v2 - v1
    const float e2x = sub(s.x3, s.x1), e2y = sub(s.y3, s.y1), e2z = sub(s.z3, s.z1);  // v3 - v1
    // orient(v1,v2,v3) = (v2x-v1x)*(v3y-v1y) - (v2y-v1y)*(v3x-v1x); == normal's raw z
    const float area = sub(mul(e1x, e2y), mul(e1y, e2x));
    sh.a = __fdiv_rn(1.0f, area);
    const float nx = sub(mul(e1y, e2z), mul(e1z, e2y));
    const float ny = sub(mul(e1z, e2x), mul(e1x, e2z));
    const float nz = area;
    // nalgebra 4-lane dot: (l0 + l2) + (l1 + l3), l3 = 0*0
    const float n2 = add(add(mul(nx, nx), mul(nz, nz)), add(mul(ny, ny), 0.0f));
    const float n = __fsqrt_rn(add(0.0f, n2));
    const float nzu = __fdiv_rn(nz, n);
    sh.k = mul(nzu, sh.a);
    sh.dz1 = e1z;
    sh.dz2 = e2z;
}

// Row constants: c_i = dx_i * (py - base_i.y)
struct RowC {
    float c0, c1, c2;
};

SLOTH_DEV RowC row_setup(const Setup& s, uint32_t y)
{
    const float py = (float)y;
    RowC r;
    r.c0 = mul(s.dx0, sub(py, s.y2));
    r.c1 = mul(s.dx1, sub(py, s.y3));
    r.c2 = mul(s.dx2, sub(py, s.y1));
    return r;
}

SLOTH_DEV void edge_eval(const Setup& s, const RowC& r, uint32_t x, float& w0, float& w1, float& w2)
{
    const float px = (float)x;
    w0 = sub(r.c0, mul(s.dy0, sub(px, s.x2)));
    w1 = sub(r.c1, mul(s.dy1, sub(px, s.x3)));
    w2 = sub(r.c2, mul(s.dy2, sub(px, s.x1)));
}

// True when an edge whose y-delta is not negative fails at this candidate.  For
// such an edge the subtracted term dy*(px - ax) is non-decreasing in px (rounding
// is monotone), so the edge also fails at every candidate further right on this
// row: the row is finished.  Requires a regular triangle (no NaN/inf anywhere).
SLOTH_DEV bool row_closed(const Setup& s, float w0, float w1, float w2)
{
    return (w0 < 0.0f && !(s.dy0 < 0.0f)) || (w1 < 0.0f && !(s.dy1 < 0.0f)) || (w2 < 0.0f && !(s.dy2 < 0.0f));
}

SLOTH_DEV uint32_t glyph_index(const FrameParams& p, float shade)
{
    uint32_t g = 9;
#pragma unroll
    for (int i = 8; i >= 0; --i)
        if (shade <= p.thr[i]) g = i;
    return g;
}

// Emit one covered fragment: depth, glyph, packed key, atomicMin into the key plane.
//   key = orderable(z) << 32 | tri << 5 | direct << 4 | glyph
// min over keys == the sequential strict-`<` depth test with first-wins ties
// (rasterizer.rs:81): smaller z wins; equal z -> smaller triangle index; same
// triangle hitting one id twice (row wrap) -> the wrapped fragment of row y-1
// precedes the direct fragment of row y.
// LEAN (k_tri's whole-frame instantiations): the context owns all rows, so krow0 = 0, and the SLOTH_DEBUG bit that
// skips the atomic is not looked at.
template <bool LEAN = false>
SLOTH_DEV void emit_fragment(const FrameParams& p, const Setup& s, const Shade& sh, uint32_t tri, uint32_t x,
                             uint32_t y, float w0, float w1, float w2, unsigned long long* __restrict__ keys)
{
    const float z = add(s.z1, mul(sh.a, add(mul(w1, sh.dz1), mul(w2, sh.dz2))));
    if (!(z < 3.40282347e+38f)) return;  // z_buffer starts at f32::MAX; NaN never wins
    const uint32_t kx = x * p.XS;
    const uint32_t direct = kx < p.KW ? 1u : 0u;
    const uint32_t row = y + 1u - direct;
    if ((!LEAN && row < p.krow0) || row >= p.row1) return;
    const float shade = mul(sh.k, add(add(w0, w1), w2));
    const uint32_t g = glyph_index(p, shade);
    uint32_t zb = __float_as_uint(z);
    if (zb == 0x80000000u) zb = 0u;  // -0.0 < +0.0 is false in the reference
    const uint32_t ord = (zb & 0x80000000u) ? ~zb : (zb | 0x80000000u);
    const unsigned long long key =
        ((unsigned long long)ord << 32) | (unsigned long long)((tri << 5) | (direct << 4) | g);
    const uint32_t L = LEAN ? y * p.KW + kx : y * p.KW + kx - p.krow0 * p.KW;
    if (LEAN || !(p.debug & 1u)) atomicMin(keys + L, key);
}

// Number of rows one row-band work item of the walk kernel covers.
SLOTH_DEV uint32_t walk_rows_per_item(uint32_t tight_w)
{
    const uint32_t w = tight_w < 32u ? 32u : tight_w;
    uint32_t r = 2048u / w;
    return r < 1u ? 1u : r;
}

// Tight column count [minx, min(maxx, floor(max_x)+1)) -- only a size estimate
// for work distribution, never used to skip candidates.
SLOTH_DEV uint32_t tight_width(const Setup& s)
{
    const uint32_t f = __float2uint_rz(floorf(s.mx0));
    const uint32_t te = f >= s.maxx ? s.maxx : f + 1u;
    return te > s.minx ? te - s.minx : 0u;
}

}  // namespace sloth
