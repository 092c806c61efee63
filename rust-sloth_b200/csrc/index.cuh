// index.cuh -- indexed scene for the per-vertex transform stage (sm_100a).
//
// The reference's meshes are indexed until to_simple_mesh de-indexes them (geometry.rs:99-107 reads
// positions[indices[..]]), and Triangle::mul (geometry.rs:43-48) is a pure function of one vertex: the
// transformed vertex is bit-identical whichever triangle asks for it.  At scene-set time the soup's 3N corners
// are therefore deduplicated by exact bit pattern (so -0.0 / +0.0 and NaN payloads stay distinct) into
//   pos    unique vertices, SoA x[] y[] z[], numbered in order of first appearance in the draw order
//          (neighbours in the soup stay neighbours in memory)
//   rec    one uint4 per triangle: (i0, i1, i2, 0), padded to a multiple of 32 triangles with a sentinel vertex
// and every frame runs k_xform (one thread per unique vertex, same xform_row order as the soup path) before the
// triangle kernel k_tri gathers (x', y') -- and (z') only for covering triangles.
//
// Build passes (once per scene; none of this is on the per-frame path):
//   k_ix_insert   open-addressing table of corner indices; a slot's vertex identity never changes, its value
//                 converges to the smallest corner index with those bits (atomicMin)
//   k_ix_lookup   rep[c] = that smallest corner index
//   k_ix_flag_sums / k_ix_scan_sums / k_ix_emit   exclusive scan of "corner is its own representative" ->
//                 vertex id = rank of the first appearance; writes pos and rec
#pragma once
#include "raster_core.cuh"

namespace sloth {
namespace ix {

static constexpr uint32_t EMPTY = 0xFFFFFFFFu;
static constexpr uint32_t SCAN_BLOCK = 1024;   // corners per block of the flag scan

SLOTH_DEV uint32_t hash3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t h = a * 0x9E3779B1u;
    h = (h ^ (h >> 15)) + b * 0x85EBCA77u;
    h = (h ^ (h >> 13)) + c * 0xC2B2AE3Du;
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    return h;
}

// bits of corner c (= 3 * triangle + k) from the resident soup streams
SLOTH_DEV void corner_bits(const Scene& sc, uint32_t corner, uint32_t& bx, uint32_t& by, uint32_t& bz)
{
    const uint32_t t = corner / 3u, k = corner - 3u * t;
    const float4 A = sc.a[t];
    const float4 B = sc.b[t];
    float x, y, z;
    if (k == 0u) { x = A.x; y = A.y; z = A.z; }
    else if (k == 1u) { x = A.w; y = B.x; z = B.y; }
    else { x = B.z; y = B.w; z = sc.z3[t]; }
    bx = __float_as_uint(x); by = __float_as_uint(y); bz = __float_as_uint(z);
}

__global__ void __launch_bounds__(256) k_ix_insert(const Scene sc, uint32_t n_corners, uint32_t* __restrict__ table, uint32_t mask)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_corners) return;
    uint32_t bx, by, bz;
    corner_bits(sc, c, bx, by, bz);
    uint32_t h = hash3(bx, by, bz) & mask;
    for (;;) {
        uint32_t cur = *reinterpret_cast<volatile uint32_t*>(table + h);
        if (cur == EMPTY) {
            cur = atomicCAS(table + h, EMPTY, c);
            if (cur == EMPTY) return;   // this corner founded the slot
        }
        uint32_t ox, oy, oz;
        corner_bits(sc, cur, ox, oy, oz);
        if (ox == bx && oy == by && oz == bz) {
            if (c < cur) atomicMin(table + h, c);
            return;
        }
        h = (h + 1u) & mask;
    }
}

__global__ void __launch_bounds__(256) k_ix_lookup(const Scene sc, uint32_t n_corners, const uint32_t* __restrict__ table, uint32_t mask,
                                                   uint32_t* __restrict__ rep)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_corners) return;
    uint32_t bx, by, bz;
    corner_bits(sc, c, bx, by, bz);
    uint32_t h = hash3(bx, by, bz) & mask;
    for (;;) {
        const uint32_t cur = table[h];   // never EMPTY before the match: this corner was inserted
        uint32_t ox, oy, oz;
        corner_bits(sc, cur, ox, oy, oz);
        if (ox == bx && oy == by && oz == bz) { rep[c] = cur; return; }
        h = (h + 1u) & mask;
    }
}

// per block of SCAN_BLOCK corners: how many are their own representative
__global__ void __launch_bounds__(256) k_ix_flag_sums(const uint32_t* __restrict__ rep, uint32_t n_corners, uint32_t* __restrict__ block_sum)
{
    __shared__ uint32_t warp_sum[8];
    const uint32_t base = blockIdx.x * SCAN_BLOCK;
    uint32_t cnt = 0;
#pragma unroll
    for (uint32_t k = 0; k < SCAN_BLOCK / 256u; ++k) {
        const uint32_t c = base + k * 256u + threadIdx.x;
        cnt += (c < n_corners && rep[c] == c) ? 1u : 0u;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
    if ((threadIdx.x & 31u) == 0u) warp_sum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < 8; ++w) s += warp_sum[w];
        block_sum[blockIdx.x] = s;
    }
}

// one block: exclusive scan of the block sums in place, total to *total
__global__ void __launch_bounds__(1024) k_ix_scan_sums(uint32_t* __restrict__ block_sum, uint32_t n_blocks, uint32_t* __restrict__ total)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_sum[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t nn = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((int)(threadIdx.x & 31u) >= d) inc += nn;
        }
        if ((threadIdx.x & 31u) == 31u) warp_tot[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32u) {
            uint32_t w = warp_tot[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t nn = __shfl_up_sync(0xFFFFFFFFu, w, d);
                if ((int)threadIdx.x >= d) w += nn;
            }
            warp_tot[threadIdx.x] = w;   // inclusive over warps
        }
        __syncthreads();
        const uint32_t warp_excl = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1u] : 0u;
        if (i < n_blocks) block_sum[i] = carry + warp_excl + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// rank[c] = vertex id of corner c if it is a first appearance (exclusive scan of the flags), written for those
// corners only; their positions go to pos.
__global__ void __launch_bounds__(256) k_ix_rank(const Scene sc, const uint32_t* __restrict__ rep, uint32_t n_corners,
                                                 const uint32_t* __restrict__ block_pre, uint32_t* __restrict__ rank,
                                                 float* __restrict__ px, float* __restrict__ py, float* __restrict__ pz)
{
    __shared__ uint32_t warp_sum[8];
    const uint32_t base = blockIdx.x * SCAN_BLOCK;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // thread owns SCAN_BLOCK/256 consecutive corners so that ranks follow the corner order
    constexpr uint32_t PER = SCAN_BLOCK / 256u;
    const uint32_t c0 = base + threadIdx.x * PER;
    bool first[PER];
    uint32_t cnt = 0;
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) {
        const uint32_t c = c0 + k;
        first[k] = c < n_corners && rep[c] == c;
        cnt += first[k] ? 1u : 0u;
    }
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t nn = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if ((int)lane >= d) inc += nn;
    }
    if (lane == 31u) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t pre = block_pre[blockIdx.x] + inc - cnt;
    for (uint32_t w = 0; w < warp; ++w) pre += warp_sum[w];
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) {
        if (!first[k]) continue;
        const uint32_t c = c0 + k;
        uint32_t bx, by, bz;
        corner_bits(sc, c, bx, by, bz);
        rank[c] = pre;
        px[pre] = __uint_as_float(bx);
        py[pre] = __uint_as_float(by);
        pz[pre] = __uint_as_float(bz);
        ++pre;
    }
}

// rec[t] = (vertex ids of the three corners, 0); triangles [n_tri, n_padded) reference the sentinel vertex
__global__ void __launch_bounds__(256) k_ix_records(const uint32_t* __restrict__ rep, const uint32_t* __restrict__ rank, uint32_t n_tri,
                                                    uint32_t n_padded, uint32_t sentinel, uint4* __restrict__ rec)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_padded) return;
    if (t >= n_tri) { rec[t] = make_uint4(sentinel, sentinel, sentinel, 0u); return; }
    rec[t] = make_uint4(rank[rep[3u * t]], rank[rep[3u * t + 1u]], rank[rep[3u * t + 2u]], 0u);
}

// Connectivity flag of every chunk of 32 triangles (bit 0 of rec[..].w, same value in all 32 records): set when the
// chunk is full and its triangles form one component under "share a vertex id".  k_tri then stamps the chunk's rows
// as one interval.  One warp per chunk, lane = triangle: labels start as the lane index and take the minimum over
// every triangle they share a vertex with, until nothing changes.
__global__ void __launch_bounds__(256) k_ix_connectivity(uint4* __restrict__ rec, uint32_t n_tri, uint32_t n_chunks)
{
    const uint32_t chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (chunk >= n_chunks) return;
    const uint32_t t = chunk * 32u + lane;
    uint4 r = rec[t];
    const bool full = chunk * 32u + 32u <= n_tri;
    uint32_t label = lane;
    for (int round = 0; round < 32; ++round) {
        const uint32_t before = label;
        for (int s = 0; s < 32; ++s) {
            const uint32_t ax = __shfl_sync(0xFFFFFFFFu, r.x, s), ay = __shfl_sync(0xFFFFFFFFu, r.y, s), az = __shfl_sync(0xFFFFFFFFu, r.z, s);
            const uint32_t ls = __shfl_sync(0xFFFFFFFFu, label, s);
            const bool share = r.x == ax || r.x == ay || r.x == az || r.y == ax || r.y == ay || r.y == az || r.z == ax || r.z == ay ||
                               r.z == az;
            if (share && ls < label) label = ls;
            // the other direction: triangle s learns this lane's label in the same step
            const uint32_t back = __reduce_min_sync(0xFFFFFFFFu, share ? label : 0xFFFFFFFFu);
            if ((int)lane == s && back < label) label = back;
        }
        if (!__any_sync(0xFFFFFFFFu, label != before)) break;
    }
    const bool one = full && __all_sync(0xFFFFFFFFu, label == 0u);
    r.w = one ? 1u : 0u;
    rec[t] = r;
}

// sloth_scene_set_indexed: positions[indices[..]] (geometry.rs:99-107) -> the resident soup streams and colours;
// ids are checked against n_vert (*bad counts the triangles that fail, they become all-zero triangles).
__global__ void __launch_bounds__(256) k_ix_expand_input(const float* __restrict__ pos, uint32_t n_vert, const uint32_t* __restrict__ idx,
                                                         const uint8_t* __restrict__ rgb, uint32_t n_tri, float4* __restrict__ a,
                                                         float4* __restrict__ b, float* __restrict__ z3, uint32_t* __restrict__ col,
                                                         uint32_t* __restrict__ bad)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tri) return;
    const uint32_t i0 = idx[3u * t], i1 = idx[3u * t + 1u], i2 = idx[3u * t + 2u];
    col[t] = (uint32_t)rgb[(size_t)t * 3] | ((uint32_t)rgb[(size_t)t * 3 + 1] << 8) | ((uint32_t)rgb[(size_t)t * 3 + 2] << 16);
    if (i0 >= n_vert || i1 >= n_vert || i2 >= n_vert) {
        atomicAdd(bad, 1u);
        a[t] = b[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        z3[t] = 0.f;
        return;
    }
    const float* p0 = pos + (size_t)i0 * 3u;
    const float* p1 = pos + (size_t)i1 * 3u;
    const float* p2 = pos + (size_t)i2 * 3u;
    a[t] = make_float4(p0[0], p0[1], p0[2], p1[0]);
    b[t] = make_float4(p1[1], p1[2], p2[0], p2[1]);
    z3[t] = p2[2];
}

}  // namespace ix

// ---------------------------------------------------------------------------------
// k_xform: Triangle::mul (geometry.rs:43-48) once per unique vertex, gemv/axcpy order (xform_row), so every
// x', y', z' is bit-identical to what the soup path computes per corner.  Four consecutive vertices per thread:
// three 16-byte SoA loads, (x', y') as two 16-byte stores (what k_tri gathers), z' as one (gathered for covering
// triangles only) -- few instructions and enough bytes in flight per thread to run beside the previous frame's
// k_tri at one block per SM.  Slot n_vert is the sentinel the padding records point at: far off-screen, no rows,
// no candidates.  The position arrays and the outputs are padded to whole groups of four.
// ---------------------------------------------------------------------------------
static constexpr uint32_t XFORM_PER_THREAD = 4;

__global__ void __launch_bounds__(256) k_xform(const __grid_constant__ FrameParams p, const float* __restrict__ px,
                                               const float* __restrict__ py, const float* __restrict__ pz, uint32_t n_vert,
                                               float2* __restrict__ vxy, float* __restrict__ vz)
{
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * XFORM_PER_THREAD;
    if (i > n_vert) return;
    const float4 X = __ldcs(reinterpret_cast<const float4*>(px + i));
    const float4 Y = __ldcs(reinterpret_cast<const float4*>(py + i));
    const float4 Z = __ldcs(reinterpret_cast<const float4*>(pz + i));
    const float x[4] = {X.x, X.y, X.z, X.w}, y[4] = {Y.x, Y.y, Y.z, Y.w}, z[4] = {Z.x, Z.y, Z.z, Z.w};
    float ox[4], oy[4], oz[4];
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k) {
        ox[k] = xform_row(p.m + 0, x[k], y[k], z[k]);
        oy[k] = xform_row(p.m + 4, x[k], y[k], z[k]);
        oz[k] = xform_row(p.m + 8, x[k], y[k], z[k]);
        if (i + k >= n_vert) { ox[k] = oy[k] = -1.0e30f; oz[k] = 0.0f; }   // the sentinel (and the padding after it)
    }
    float4* oxy = reinterpret_cast<float4*>(vxy + i);
    oxy[0] = make_float4(ox[0], oy[0], ox[1], oy[1]);
    oxy[1] = make_float4(ox[2], oy[2], ox[3], oy[3]);
    *reinterpret_cast<float4*>(vz + i) = make_float4(oz[0], oz[1], oz[2], oz[3]);
}

}  // namespace sloth
