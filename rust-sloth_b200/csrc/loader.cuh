// loader.cuh -- model loading on the device, SURVEY.md 8(f) "next-2".
//
// Replaces the load-time half of the reference for large files: src/inputs.rs:95-129 (match_meshes) hands each
// file to tobj 3.2.2 / stl_io 0.4.2 and src/geometry.rs:83-189 (to_meshes) turns the result into a triangle soup
// with colours and a bounding box.  Here the file's bytes are copied to the GPU once and parsed there:
//
//   k_count_newlines / k_line_starts   newline positions -> line table
//   k_obj_classify                     one thread per line: key, token syntax, per-line counts
//   k_obj_usemtl                       usemtl name -> material state (table built on the host from the .mtl files)
//   k_scan_*                           exclusive scan over the lines: vertex index, vertex-colour offset,
//                                      triangle offset, material state in force
//   k_obj_vertices / k_obj_faces       decimal -> f32 (correctly rounded, see dec_to_f32), fan triangulation,
//                                      colour rules, straight into the staging soup
//   k_stl_ascii_* / k_stl_binary       the two STL encodings
//   k_soup_scan                        max coordinate (context.rs:106-113) and the "clean scene" flag
//
// Rules follow host/mesh_io.cpp line for line (that file restates the crates' behaviour and is the checker for
// this one: tests/test_gpu_loader.py compares the two bit for bit).  Tokens outside the grammar handled here
// (inf/nan, more than 19 significant digits that straddle a rounding boundary, |decimal exponent| > 27) are
// reported as LD_UNSUPPORTED instead of being guessed at.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dec_float.cuh"

namespace sloth {
namespace ld {

// error codes written to the device error words (low byte; the line index sits above it)
enum : uint32_t {
    LD_OK = 0,
    LD_POSITION = 1,      // "position parse error"
    LD_FACE = 2,          // "face parse error"
    LD_MATERIAL = 3,      // "material parse error"
    LD_UNSUPPORTED = 4,   // token outside the device grammar
    LD_NO_MATERIAL = 5,   // model without material although materials exist (geometry.rs:110 unwrap)
    LD_MISSING_VERTEX = 6,
    LD_VCOL_RANGE = 7,
    LD_STL_VERTEX = 8,    // "bad vertex"
    LD_TOO_MANY_MTLLIB = 9,
};

static constexpr uint32_t LD_TILE = 8192;            // bytes of text per block in the newline passes
static constexpr uint32_t LD_MAX_MTLLIB = 256;
static constexpr uint32_t MAT_PRESENT = 0x80000000u; // material state word: a usemtl statement was seen ...
static constexpr uint32_t MAT_FOUND = 0x40000000u;   // ... and its name is in the table; low bits = material id

struct Material {
    uint32_t name_off, name_len;   // into the names blob
    uint32_t defined_at;           // byte offset of the mtllib statement that loaded it
    uint32_t rgb;                  // (Kd*255) as u8, packed r | g<<8 | b<<16
};

struct ObjTotals {                 // written by the scan
    uint32_t n_vertices, n_vcol, n_tris, final_mat;
};

__device__ __forceinline__ bool is_ws(unsigned char c) { return c == ' ' || (c >= 9u && c <= 13u); }

__device__ __forceinline__ void report(unsigned long long* err, uint32_t line, uint32_t code)
{
    atomicMin(err, ((unsigned long long)line << 8) | code);
}

// next whitespace-delimited token in [p, end); returns false when none is left
__device__ __forceinline__ bool next_token(const unsigned char*& p, const unsigned char* end, const unsigned char*& tb,
                                           const unsigned char*& te)
{
    while (p < end && is_ws(*p)) ++p;
    if (p >= end) return false;
    tb = p;
    while (p < end && !is_ws(*p)) ++p;
    te = p;
    return true;
}

__device__ __forceinline__ bool token_is(const unsigned char* tb, const unsigned char* te, const char* s, int n)
{
    if (te - tb != n) return false;
    for (int i = 0; i < n; ++i)
        if (tb[i] != (unsigned char)s[i]) return false;
    return true;
}

// strtol(first part of a face corner, base 10) consuming everything up to the first '/'
__device__ inline bool parse_corner(const unsigned char* tb, const unsigned char* te, long long n_pos, long long& vi)
{
    const unsigned char* e = tb;
    while (e < te && *e != '/') ++e;
    const unsigned char* p = tb;
    bool neg = false;
    if (p < e && (*p == '+' || *p == '-')) { neg = *p == '-'; ++p; }
    if (p >= e) return false;
    long long x = 0;
    for (; p < e; ++p) {
        if (!is_digit(*p)) return false;
        if (x < (1ll << 56)) x = x * 10 + (long long)(*p - '0');
    }
    vi = neg ? n_pos - x : x - 1;
    return true;
}

// ---- line table ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count_newlines(const unsigned char* __restrict__ text, uint32_t len,
                                                        uint32_t* __restrict__ tile_count)
{
    __shared__ uint32_t s_warp[8];
    const uint32_t base = blockIdx.x * LD_TILE + threadIdx.x * 32u;
    uint32_t n = 0;
    if (base + 32u <= len) {
        const uint4* q = reinterpret_cast<const uint4*>(text + base);   // text is 256-byte aligned, base % 32 == 0
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint4 v = q[k];
            const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t x = wds[j] ^ 0x0A0A0A0Au;                           // zero byte where '\n'
                n += __popc(~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu));   // exact zero-byte mask
            }
        }
    } else {
        for (uint32_t i = base; i < len && i < base + 32u; ++i) n += text[i] == '\n';
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, d);
    if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += s_warp[w];
        tile_count[blockIdx.x] = t;
    }
}

// exclusive scan of up to a few 10^5 tile counts with one block; total to *total
__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t* __restrict__ tile, uint32_t n, uint32_t* __restrict__ total)
{
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (n + 1023u) / 1024u;
    const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += tile[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024u; d <<= 1) {
        const uint32_t v = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - sum;
    for (uint32_t i = lo; i < hi; ++i) {
        const uint32_t c = tile[i];
        tile[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) *total = s_part[1023];
}

// line_start[0] = 0, line_start[k] = 1 + position of the k-th '\n', line_start[n_lines] = len + 1
__global__ void __launch_bounds__(256) k_line_starts(const unsigned char* __restrict__ text, uint32_t len,
                                                     const uint32_t* __restrict__ tile_off, uint32_t n_newlines,
                                                     uint32_t* __restrict__ line_start)
{
    __shared__ uint32_t s_warp[8];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * LD_TILE + threadIdx.x * 32u;
    uint32_t bits = 0;
    for (uint32_t i = 0; i < 32u && base + i < len; ++i) bits |= (uint32_t)(text[base + i] == '\n') << i;
    const uint32_t mine = __popc(bits);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t off = tile_off[blockIdx.x] + incl - mine;
    for (uint32_t w = 0; w < warp; ++w) off += s_warp[w];
    while (bits) {
        const uint32_t i = __ffs((int)bits) - 1u;
        bits &= bits - 1u;
        line_start[++off] = base + i + 1u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        line_start[0] = 0u;
        line_start[n_newlines + 1u] = len + 1u;
    }
}

// ---- OBJ ----------------------------------------------------------------------------------------
// rec[line] = (is vertex line, vertex-colour floats pushed, triangles emitted, material state set by the line)
__global__ void __launch_bounds__(256) k_obj_classify(const unsigned char* __restrict__ text, const uint32_t* __restrict__ line_start,
                                                      uint32_t n_lines, uint4* __restrict__ rec, uint32_t* __restrict__ mtllib_lines,
                                                      uint32_t* __restrict__ mtllib_count, unsigned long long* __restrict__ err)
{
    const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines) return;
    const unsigned char* p = text + line_start[line];
    const unsigned char* const end = text + line_start[line + 1] - 1u;
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    const unsigned char *tb, *te;
    if (next_token(p, end, tb, te)) {
        if (token_is(tb, te, "v", 1)) {
            // tobj parse_floatn(.., 3): three tokens, each a float; then up to three colour floats, stopping
            // silently at the first one that does not parse
            uint32_t got = 0, status = TOK_OK;
            float f;
            for (; got < 3u && next_token(p, end, tb, te); ++got)
                if ((status = parse_float(tb, te, f)) != TOK_OK) break;
            if (status == TOK_UNSUPPORTED) report(err, line, LD_UNSUPPORTED);
            else if (status != TOK_OK || got < 3u) report(err, line, LD_POSITION);
            else {
                uint32_t ncol = 0;
                for (; ncol < 3u && next_token(p, end, tb, te); ++ncol) {
                    status = parse_float(tb, te, f);
                    if (status == TOK_UNSUPPORTED) report(err, line, LD_UNSUPPORTED);
                    if (status != TOK_OK) break;
                }
                r.x = 1u;
                r.y = ncol;
            }
        } else if (token_is(tb, te, "f", 1) || token_is(tb, te, "l", 1)) {
            uint32_t corners = 0;
            bool ok = true;
            long long vi;
            while (next_token(p, end, tb, te)) {
                ok = ok && parse_corner(tb, te, 0, vi);
                ++corners;
            }
            if (!ok || corners == 0u) report(err, line, LD_FACE);
            else r.z = corners >= 3u ? corners - 2u : 0u;   // points and lines are dropped (GPU_LOAD_OPTIONS)
        } else if (token_is(tb, te, "mtllib", 6)) {
            if (!next_token(p, end, tb, te)) report(err, line, LD_MATERIAL);
            else {
                const uint32_t k = atomicAdd(mtllib_count, 1u);
                if (k < LD_MAX_MTLLIB) mtllib_lines[k] = line;
                else report(err, line, LD_TOO_MANY_MTLLIB);
            }
        } else if (token_is(tb, te, "usemtl", 6)) {
            if (!next_token(p, end, tb, te)) report(err, line, LD_MATERIAL);   // empty name
            else r.w = MAT_PRESENT;                                             // resolved by k_obj_usemtl
        }
    }
    rec[line] = r;
}

// usemtl: the rest of the line after the keyword, trimmed, looked up among the materials loaded by mtllib
// statements that precede the line (the map is filled as the file is read; a later definition of a name wins)
__global__ void __launch_bounds__(256) k_obj_usemtl(const unsigned char* __restrict__ text, const uint32_t* __restrict__ line_start,
                                                    uint32_t n_lines, uint4* __restrict__ rec, const Material* __restrict__ mats,
                                                    uint32_t n_mats, const unsigned char* __restrict__ names)
{
    const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines || rec[line].w == 0u) return;
    const unsigned char* p = text + line_start[line];
    const unsigned char* end = text + line_start[line + 1] - 1u;
    const unsigned char *tb, *te;
    next_token(p, end, tb, te);   // the keyword
    while (p < end && is_ws(*p)) ++p;
    while (end > p && is_ws(end[-1])) --end;
    const uint32_t n = (uint32_t)(end - p);
    uint32_t state = MAT_PRESENT;
    for (uint32_t m = 0; m < n_mats; ++m) {
        if (mats[m].name_len != n || mats[m].defined_at >= line_start[line]) continue;
        bool same = true;
        for (uint32_t i = 0; i < n && same; ++i) same = names[mats[m].name_off + i] == p[i];
        if (same) state = MAT_PRESENT | MAT_FOUND | m;   // keep going: the last definition wins
    }
    rec[line].w = state;
}

// ---- exclusive scan over the line records: (sum, sum, sum, last non-zero) --------------------------
static constexpr uint32_t SCAN_PER_THREAD = 4, SCAN_THREADS = 256, SCAN_PER_BLOCK = SCAN_PER_THREAD * SCAN_THREADS;

__device__ __forceinline__ uint4 rec_combine(const uint4& a, const uint4& b)   // a before b
{
    return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, b.w ? b.w : a.w);
}

__device__ __forceinline__ uint4 shfl_up4(const uint4& v, int d)
{
    return make_uint4(__shfl_up_sync(0xFFFFFFFFu, v.x, d), __shfl_up_sync(0xFFFFFFFFu, v.y, d),
                      __shfl_up_sync(0xFFFFFFFFu, v.z, d), __shfl_up_sync(0xFFFFFFFFu, v.w, d));
}

// block-wide scan of one value per thread: returns the combination of all values of lower-numbered threads
// (the exclusive prefix); the block total goes to `total`
__device__ inline uint4 block_scan4(uint4 v, uint4* s_warp, uint4& total)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint4 u = shfl_up4(v, d);
        if ((int)lane >= d) v = rec_combine(u, v);
    }
    if (lane == 31) s_warp[warp] = v;
    uint4 excl = shfl_up4(v, 1);
    if (lane == 0) excl = zero;
    __syncthreads();
    uint4 before = zero;
    total = zero;
    for (uint32_t w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp) before = rec_combine(before, s_warp[w]);
        total = rec_combine(total, s_warp[w]);
    }
    __syncthreads();   // s_warp may be reused by the caller's next round
    return rec_combine(before, excl);
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint4* __restrict__ rec, uint32_t n, uint4* __restrict__ block_sum)
{
    __shared__ uint4 s_warp[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_PER_BLOCK + threadIdx.x * SCAN_PER_THREAD;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (uint32_t k = 0; k < SCAN_PER_THREAD; ++k)
        if (base + k < n) v = rec_combine(v, rec[base + k]);
    uint4 total;
    block_scan4(v, s_warp, total);
    if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}

// one block: exclusive scan of the block sums in place, grand total to *totals
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocks(uint4* __restrict__ block_sum, uint32_t n_blocks, ObjTotals* __restrict__ totals)
{
    __shared__ uint4 s_warp[SCAN_THREADS / 32];
    uint4 carry = make_uint4(0u, 0u, 0u, 0u);   // same value in every thread
    for (uint32_t base = 0; base < n_blocks; base += SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint4 v = i < n_blocks ? block_sum[i] : make_uint4(0u, 0u, 0u, 0u);
        uint4 total;
        const uint4 excl = block_scan4(v, s_warp, total);
        if (i < n_blocks) block_sum[i] = rec_combine(carry, excl);
        carry = rec_combine(carry, total);
    }
    if (threadIdx.x == 0) {
        totals->n_vertices = carry.x; totals->n_vcol = carry.y; totals->n_tris = carry.z; totals->final_mat = carry.w;
    }
}

// pre[line] = combination of all records before the line
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint4* __restrict__ rec, uint32_t n, const uint4* __restrict__ block_pre,
                                                             uint4* __restrict__ pre)
{
    __shared__ uint4 s_warp[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_PER_BLOCK + threadIdx.x * SCAN_PER_THREAD;
    uint4 r[SCAN_PER_THREAD];
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (uint32_t k = 0; k < SCAN_PER_THREAD; ++k) {
        r[k] = base + k < n ? rec[base + k] : make_uint4(0u, 0u, 0u, 0u);
        v = rec_combine(v, r[k]);
    }
    uint4 total;
    const uint4 excl = block_scan4(v, s_warp, total);
    uint4 run = rec_combine(block_pre[blockIdx.x], excl);
#pragma unroll
    for (uint32_t k = 0; k < SCAN_PER_THREAD; ++k) {
        if (base + k < n) pre[base + k] = run;
        run = rec_combine(run, r[k]);
    }
}

__global__ void __launch_bounds__(256) k_obj_vertices(const unsigned char* __restrict__ text, const uint32_t* __restrict__ line_start,
                                                      uint32_t n_lines, const uint4* __restrict__ rec, const uint4* __restrict__ pre,
                                                      float* __restrict__ pos, float* __restrict__ vcol)
{
    const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines || rec[line].x == 0u) return;
    const unsigned char* p = text + line_start[line];
    const unsigned char* const end = text + line_start[line + 1] - 1u;
    const unsigned char *tb, *te;
    next_token(p, end, tb, te);   // "v"
    const uint4 before = pre[line];
    for (uint32_t d = 0; d < 3u; ++d) {
        float f = 0.0f;
        next_token(p, end, tb, te);
        parse_float(tb, te, f);   // validated by k_obj_classify
        pos[(size_t)before.x * 3u + d] = f;
    }
    const uint32_t ncol = rec[line].y;
    for (uint32_t d = 0; d < ncol; ++d) {
        float f = 0.0f;
        next_token(p, end, tb, te);
        parse_float(tb, te, f);
        vcol[(size_t)before.y + d] = f;
    }
}

// Rust `f32 as u8`: truncate toward zero, saturate, NaN -> 0 (geometry.rs:111-124)
__device__ __forceinline__ uint32_t f32_as_u8(float v)
{
    if (!(v > 0.0f)) return 0u;
    if (v >= 255.0f) return 255u;
    return (uint32_t)v;
}

// fan triangulation (0, i-1, i) of every face into the soup, colour rules of geometry.rs:91-126
__global__ void __launch_bounds__(256) k_obj_faces(const unsigned char* __restrict__ text, const uint32_t* __restrict__ line_start,
                                                   uint32_t n_lines, const uint4* __restrict__ rec, const uint4* __restrict__ pre,
                                                   const float* __restrict__ pos, const float* __restrict__ vcol, ObjTotals tot,
                                                   const Material* __restrict__ mats, uint32_t n_mats, float* __restrict__ soup_xyz,
                                                   uint8_t* __restrict__ soup_rgb, unsigned long long* __restrict__ err)
{
    const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines || rec[line].z == 0u) return;
    const unsigned char* p = text + line_start[line];
    const unsigned char* const end = text + line_start[line + 1] - 1u;
    const unsigned char *tb, *te;
    next_token(p, end, tb, te);   // "f" / "l"
    const uint4 before = pre[line];
    uint32_t base_rgb = 0x010101u;   // geometry.rs:91
    if (n_mats) {
        if (!(before.w & MAT_FOUND)) { report(err, line, LD_NO_MATERIAL); return; }
        base_rgb = mats[before.w & 0x3FFFFFFFu].rgb;
    }
    long long first = 0, prev = 0, cur = 0;
    uint32_t corner = 0;
    size_t tri = before.z;
    while (next_token(p, end, tb, te)) {
        parse_corner(tb, te, (long long)before.x, cur);
        if (corner == 0u) first = cur;
        if (corner >= 2u) {
            const long long idx[3] = {first, prev, cur};
            bool ok = true;
            for (int k = 0; k < 3; ++k) ok = ok && idx[k] >= 0 && idx[k] < (long long)tot.n_vertices;
            if (!ok) { report(err, line, LD_MISSING_VERTEX); return; }
            for (int k = 0; k < 3; ++k)
                for (int d = 0; d < 3; ++d) soup_xyz[tri * 9u + k * 3 + d] = pos[(size_t)idx[k] * 3u + d];
            uint32_t rgb = base_rgb;
            if (n_mats && tot.n_vcol) {   // first corner's vertex colour (geometry.rs:117-126)
                const size_t ci = (size_t)first * 3u;
                if (ci + 2u >= tot.n_vcol) { report(err, line, LD_VCOL_RANGE); return; }
                rgb = f32_as_u8(__fmul_rn(vcol[ci], 255.0f)) | f32_as_u8(__fmul_rn(vcol[ci + 1u], 255.0f)) << 8 |
                      f32_as_u8(__fmul_rn(vcol[ci + 2u], 255.0f)) << 16;
            }
            soup_rgb[tri * 3u] = (uint8_t)rgb;
            soup_rgb[tri * 3u + 1u] = (uint8_t)(rgb >> 8);
            soup_rgb[tri * 3u + 2u] = (uint8_t)(rgb >> 16);
            ++tri;
        }
        prev = cur;
        ++corner;
    }
}

// ---- STL ----------------------------------------------------------------------------------------
// ASCII: a line of exactly four tokens starting with "vertex" is a vertex (stl_io reads facets of three)
__global__ void __launch_bounds__(256) k_stl_ascii_classify(const unsigned char* __restrict__ text, const uint32_t* __restrict__ line_start,
                                                            uint32_t n_lines, uint4* __restrict__ rec, unsigned long long* __restrict__ err)
{
    const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines) return;
    const unsigned char* p = text + line_start[line];
    const unsigned char* const end = text + line_start[line + 1] - 1u;
    const unsigned char *tb, *te, *t1b = nullptr, *t1e = nullptr, *fb[3], *fe[3];
    uint32_t n = 0;
    while (next_token(p, end, tb, te)) {
        if (n == 0) { t1b = tb; t1e = te; }
        else if (n <= 3) { fb[n - 1] = tb; fe[n - 1] = te; }
        ++n;
    }
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (n == 4u && token_is(t1b, t1e, "vertex", 6)) {
        bool ok = true;
        for (int d = 0; d < 3; ++d) {
            float f;
            const uint32_t s = parse_float(fb[d], fe[d], f);
            if (s == TOK_UNSUPPORTED) report(err, line, LD_UNSUPPORTED);
            else if (s != TOK_OK) report(err, line, LD_STL_VERTEX);
            ok = ok && s == TOK_OK;
        }
        r.x = ok ? 1u : 0u;
    }
    rec[line] = r;
}

__global__ void __launch_bounds__(256) k_stl_ascii_vertices(const unsigned char* __restrict__ text, const uint32_t* __restrict__ line_start,
                                                            uint32_t n_lines, const uint4* __restrict__ rec, const uint4* __restrict__ pre,
                                                            float* __restrict__ soup_xyz)
{
    const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines || rec[line].x == 0u) return;
    const unsigned char* p = text + line_start[line];
    const unsigned char* const end = text + line_start[line + 1] - 1u;
    const unsigned char *tb, *te;
    next_token(p, end, tb, te);   // "vertex"
    for (uint32_t d = 0; d < 3u; ++d) {
        float f = 0.0f;
        next_token(p, end, tb, te);
        parse_float(tb, te, f);
        soup_xyz[(size_t)pre[line].x * 3u + d] = f;
    }
}

// binary: 80-byte header, u32 count, 50-byte records (normal, 3 vertices, attribute); one thread per float
__global__ void __launch_bounds__(256) k_stl_binary(const unsigned char* __restrict__ bytes, uint32_t n_tri, float* __restrict__ soup_xyz)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_tri * 9u) return;
    const size_t t = i / 9u, k = i % 9u;
    const unsigned short* h = reinterpret_cast<const unsigned short*>(bytes + 84u + t * 50u + 12u + k * 4u);   // 2-byte aligned
    soup_xyz[i] = __uint_as_float((uint32_t)h[0] | (uint32_t)h[1] << 16);
}

__global__ void __launch_bounds__(256) k_fill_rgb(uint8_t* __restrict__ rgb, size_t n_tri, uint32_t colour)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tri * 3u) return;
    rgb[i] = (uint8_t)(colour >> (8u * (uint32_t)(i % 3u)));
}

// ---- soup statistics ------------------------------------------------------------------------------
// out[0] = bits of max(0, max coordinate) (fold of fmax from 0.0, context.rs:106-113: NaN never wins),
// out[1] = 1 when some coordinate is NaN or beyond 2^20 in magnitude (the "regular scene" shortcut is off then)
__global__ void __launch_bounds__(256) k_soup_scan(const float* __restrict__ xyz, size_t n, uint32_t* __restrict__ out)
{
    float m = 0.0f;
    bool dirty = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = xyz[i];
        m = fmaxf(m, v);
        dirty |= !(fabsf(v) <= 1048576.0f);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
    const bool any_dirty = __any_sync(0xFFFFFFFFu, dirty);
    if ((threadIdx.x & 31u) == 0u) {
        if (m > 0.0f) atomicMax(out, __float_as_uint(m));   // positive floats order like their bit patterns
        if (any_dirty) out[1] = 1u;
    }
}

// inverse of k_pack_scene for sloth_scene_get
__global__ void __launch_bounds__(256) k_unpack_scene(const float4* __restrict__ a, const float4* __restrict__ b, const float* __restrict__ z3,
                                                      const uint32_t* __restrict__ rgb, uint32_t n, float* __restrict__ xyz,
                                                      uint8_t* __restrict__ rgb_out)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float4 A = a[t], B = b[t];
    float* o = xyz + (size_t)t * 9u;
    o[0] = A.x; o[1] = A.y; o[2] = A.z; o[3] = A.w; o[4] = B.x; o[5] = B.y; o[6] = B.z; o[7] = B.w; o[8] = z3[t];
    const uint32_t c = rgb[t];
    rgb_out[(size_t)t * 3u] = (uint8_t)c;
    rgb_out[(size_t)t * 3u + 1u] = (uint8_t)(c >> 8);
    rgb_out[(size_t)t * 3u + 2u] = (uint8_t)(c >> 16);
}

}  // namespace ld
}  // namespace sloth
