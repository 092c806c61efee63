// index.cuh -- indexed scene for the per-vertex transform stage (sm_100a).
//
// The reference's meshes are indexed until to_simple_mesh de-indexes them (geometry.rs:99-107 reads
// positions[indices[..]]), and Triangle::mul (geometry.rs:43-48) is a pure function of one vertex: the
// transformed vertex is bit-identical whichever triangle asks for it.  At scene-set time the soup's 3N corners
// are therefore deduplicated by exact bit pattern (so -0.0 / +0.0 and NaN payloads stay distinct) into
//   pos    unique vertices, SoA x[] y[] z[], numbered in order of first appearance in the draw order
//          (neighbours in the soup stay neighbours in memory)
//   rec    one uint4 per triangle: (i0, i1, i2, triangle << 1 | chunk-connected flag), padded to a multiple of 32
//          triangles with a sentinel vertex -- the record names its own triangle, so a parked fragment is the
//          record itself (one 16-byte shared-memory store) and k_tri carries no chunk index through its pipeline
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

// rec[t] = (vertex ids of the three corners, t << 1); triangles [n_tri, n_padded) reference the sentinel vertex
__global__ void __launch_bounds__(256) k_ix_records(const uint32_t* __restrict__ rep, const uint32_t* __restrict__ rank, uint32_t n_tri,
                                                    uint32_t n_padded, uint32_t sentinel, uint4* __restrict__ rec)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_padded) return;
    if (t >= n_tri) { rec[t] = make_uint4(sentinel, sentinel, sentinel, t << 1); return; }
    rec[t] = make_uint4(rank[rep[3u * t]], rank[rep[3u * t + 1u]], rank[rep[3u * t + 2u]], t << 1);
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
    r.w = (t << 1) | (one ? 1u : 0u);   // t < 2^27 (sloth_scene_set refuses larger scenes)
    rec[t] = r;
}

// ---------------------------------------------------------------------------------
// Super-chunks: SC_CHUNKS consecutive chunks (4 = 128 triangles by default).  Per super-chunk, once per scene:
//   * the cone of its triangles' object-space normals n_t = (V1-V3) x (V2-V1) (axis, half-angle), the smallest |n_t|
//     and the longest edge -- the rotation-independent half of the certificate "every triangle of this super-chunk
//     is back-facing by a margin" that k_super_cert (tri_kernel.cuh) completes per frame;
//   * the list of its unique vertex ids (at most SC_IDS), from which the row range the super-chunk stamps is read
//     off without touching its triangles.
// A super-chunk is certifiable only when it is full, every triangle has a non-zero normal, the cone is narrower than
// 60 degrees, it has at most SC_IDS unique vertices, and every triangle but the first shares a vertex with an earlier
// one (then the row ranges of its triangles form one interval, see k_ix_connectivity).  One block per super-chunk.
// ---------------------------------------------------------------------------------
// Size of a super-chunk in chunks of 32 triangles: 2, 4 or 8.  Smaller super-chunks have narrower normal cones and
// reach closer to the silhouette (bench workload, k_tri alone: 102.3 us with 8, 97.9 us with 4, 95.0 us with 2), but
// every certified one costs k_super_stamp a latency chain (23 / 28 / 54 us of k_tail + k_super_stamp): 4 is the best
// frame.  (profiles/r02/r02_superchunk_size_ab.log)
#ifndef SLOTH_SC_CHUNKS
#define SLOTH_SC_CHUNKS 4
#endif
static constexpr uint32_t SC_CHUNKS = SLOTH_SC_CHUNKS;
static constexpr uint32_t SC_TRIS = SC_CHUNKS * 32u;
static constexpr uint32_t SC_WARPS = SC_CHUNKS;  // k_ix_super: one thread per triangle
static constexpr uint32_t SC_IDS = SC_TRIS + SC_TRIS / 2u;   // room for 1.5 unique vertices per triangle; 12 (6, 3) per lane of the warp
                                                              // that reads them
static constexpr uint32_t SC_SORT = 4u * SC_TRIS;             // 3 keys per triangle, padded to a power of two

struct SuperChunk {        // 32 bytes
    float ax, ay, az;      // unit cone axis times cos(half-angle)      (half-angle rounded up)
    float sin_theta;       // sin(half-angle), rounded up; 2 = not certifiable
    float inv_nmin;        // 1 / min |n_t|, rounded up
    float emax;            // longest edge of any triangle, rounded up
    uint32_t n_ids, pad;
};

__device__ __forceinline__ double block_reduce(double v, int op, double* s_part)   // op 0 sum, 1 min, 2 max; SC_WARPS warps
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
        v = op == 0 ? v + o : (op == 1 ? fmin(v, o) : fmax(v, o));
    }
    __syncthreads();
    if ((threadIdx.x & 31u) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = s_part[0];
    for (uint32_t w = 1; w < SC_WARPS; ++w) r = op == 0 ? r + s_part[w] : (op == 1 ? fmin(r, s_part[w]) : fmax(r, s_part[w]));
    return r;
}

__global__ void __launch_bounds__(SC_TRIS) k_ix_super(const uint4* __restrict__ rec, const float* __restrict__ px,
                                                      const float* __restrict__ py, const float* __restrict__ pz,
                                                      SuperChunk* __restrict__ out, uint32_t* __restrict__ ids)
{
    __shared__ unsigned long long s_key[SC_SORT];
    __shared__ double s_part[SC_WARPS];
    __shared__ uint32_t s_warp[SC_WARPS];
    __shared__ int s_fail;
    const uint32_t sc = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint4 r = rec[(size_t)sc * SC_TRIS + tid];
    const uint32_t id[3] = {r.x, r.y, r.z};
#pragma unroll
    for (int j = 0; j < 3; ++j) s_key[tid * 3u + j] = ((unsigned long long)id[j] << 32) | tid;
    s_key[3u * SC_TRIS + tid] = ~0ull;
    if (tid == 0) s_fail = 0;
    __syncthreads();
    for (uint32_t k = 2; k <= SC_SORT; k <<= 1)      // bitonic sort of (vertex id, triangle) pairs
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < SC_SORT; i += SC_TRIS) {
                const uint32_t x = i ^ j;
                if (x > i) {
                    const unsigned long long a = s_key[i], b = s_key[x];
                    if ((a > b) == ((i & k) == 0u)) { s_key[i] = b; s_key[x] = a; }
                }
            }
            __syncthreads();
        }
    // unique ids in ascending order: position i starts a new id
    uint32_t first[3], mine = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const uint32_t i = tid * 3u + j;
        first[j] = (i == 0u || (uint32_t)(s_key[i] >> 32) != (uint32_t)(s_key[i - 1u] >> 32)) ? 1u : 0u;
        mine += first[j];
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t off = incl - mine, n_ids = 0;
    for (uint32_t w = 0; w < SC_WARPS; ++w) {
        if (w < warp) off += s_warp[w];
        n_ids += s_warp[w];
    }
    uint32_t* my_ids = ids + (size_t)sc * SC_IDS;
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (first[j]) {
            if (off < SC_IDS) my_ids[off] = (uint32_t)(s_key[tid * 3u + j] >> 32);
            ++off;
        }
    const uint32_t id0 = (uint32_t)(s_key[0] >> 32);
    for (uint32_t i = n_ids + tid; i < SC_IDS; i += SC_TRIS) my_ids[i] = id0;   // padding: a vertex that is in the list anyway
    // every triangle but the first shares a vertex with an earlier one (smallest triangle index per id = the low
    // word of the id's first sorted pair)
    bool linked = tid == 0u;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        uint32_t lo = 0, hi = 3u * SC_TRIS;
        const unsigned long long want = (unsigned long long)id[j] << 32;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_key[mid] < want) lo = mid + 1u; else hi = mid;
        }
        if ((uint32_t)s_key[lo] < tid) linked = true;
    }
    // normal cone, in double
    const double x1 = px[id[0]], y1 = py[id[0]], z1 = pz[id[0]];
    const double x2 = px[id[1]], y2 = py[id[1]], z2 = pz[id[1]];
    const double x3 = px[id[2]], y3 = py[id[2]], z3 = pz[id[2]];
    const double ax = x1 - x3, ay = y1 - y3, az = z1 - z3, bx = x2 - x1, by = y2 - y1, bz = z2 - z1;
    const double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const double len = sqrt(nx * nx + ny * ny + nz * nz);
    const double e3x = x3 - x2, e3y = y3 - y2, e3z = z3 - z2;
    const double emax_t = sqrt(fmax(fmax(ax * ax + ay * ay + az * az, bx * bx + by * by + bz * bz), e3x * e3x + e3y * e3y + e3z * e3z));
    const bool good = linked && len > 0.0 && isfinite(len) && isfinite(emax_t);
    if (!good) s_fail = 1;   // benign race: every writer stores 1
    const double ux = good ? nx / len : 0.0, uy = good ? ny / len : 0.0, uz = good ? nz / len : 0.0;
    double sx = block_reduce(ux, 0, s_part), sy = block_reduce(uy, 0, s_part), sz = block_reduce(uz, 0, s_part);
    const double sl = sqrt(sx * sx + sy * sy + sz * sz);
    if (sl > 0.0) { sx /= sl; sy /= sl; sz /= sl; }
    const double cmin = block_reduce(good ? ux * sx + uy * sy + uz * sz : -1.0, 1, s_part);
    const double nmin = block_reduce(good ? len : 0.0, 1, s_part);
    const double emax = block_reduce(good ? emax_t : 0.0, 2, s_part);
    __syncthreads();
    if (tid == 0) {
        SuperChunk o;
        const double cs = cmin - 1.0e-6;   // cosine of the half-angle, rounded down
        const bool ok = !s_fail && sl > 0.0 && n_ids <= SC_IDS && cs >= 0.5 && nmin > 0.0;
        const double sn = ok ? sqrt(fmax(0.0, 1.0 - cs * cs)) + 1.0e-6 : 2.0;
        o.ax = (float)(sx * cs); o.ay = (float)(sy * cs); o.az = (float)(sz * cs);
        o.sin_theta = ok ? __double2float_ru(sn) : 2.0f;
        o.inv_nmin = ok ? __double2float_ru(1.000001 / nmin) : 0.0f;
        o.emax = ok ? __double2float_ru(emax * 1.000001) : 0.0f;
        o.n_ids = n_ids;
        o.pad = 0u;
        if (!isfinite(o.inv_nmin) || !isfinite(o.emax)) o.sin_theta = 2.0f;
        out[sc] = o;
    }
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
#ifndef XFORM_VERTS_PER_THREAD
#define XFORM_VERTS_PER_THREAD 4
#endif
static constexpr uint32_t XFORM_PER_THREAD = XFORM_VERTS_PER_THREAD;   // a multiple of 4

__global__ void __launch_bounds__(256) k_xform(const __grid_constant__ FrameParams p, const float* __restrict__ px,
                                               const float* __restrict__ py, const float* __restrict__ pz, uint32_t n_vert,
                                               float2* __restrict__ vxy, float* __restrict__ vz)
{
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * XFORM_PER_THREAD;
    if (i0 > n_vert) return;
    float4 X[XFORM_PER_THREAD / 4], Y[XFORM_PER_THREAD / 4], Z[XFORM_PER_THREAD / 4];
#pragma unroll
    for (uint32_t g = 0; g < XFORM_PER_THREAD / 4; ++g) {   // all loads first: the kernel is latency, not arithmetic
        X[g] = __ldcs(reinterpret_cast<const float4*>(px + i0 + 4u * g));
        Y[g] = __ldcs(reinterpret_cast<const float4*>(py + i0 + 4u * g));
        Z[g] = __ldcs(reinterpret_cast<const float4*>(pz + i0 + 4u * g));
    }
#pragma unroll
    for (uint32_t g = 0; g < XFORM_PER_THREAD / 4; ++g) {
        const uint32_t i = i0 + 4u * g;
        const float x[4] = {X[g].x, X[g].y, X[g].z, X[g].w}, y[4] = {Y[g].x, Y[g].y, Y[g].z, Y[g].w}, z[4] = {Z[g].x, Z[g].y, Z[g].z, Z[g].w};
        float ox[4], oy[4], oz[4];
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {
            ox[k] = xform_row(p.m + 0, x[k], y[k], z[k]);
            oy[k] = xform_row(p.m + 4, x[k], y[k], z[k]);
            oz[k] = xform_row(p.m + 8, x[k], y[k], z[k]);
        }
        if (i + 4u > n_vert) {   // only the thread that holds the end of the array: the sentinel (and the padding after it)
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k)
                if (i + k >= n_vert) { ox[k] = oy[k] = -1.0e30f; oz[k] = 0.0f; }
        }
        float4* oxy = reinterpret_cast<float4*>(vxy + i);
        oxy[0] = make_float4(ox[0], oy[0], ox[1], oy[1]);
        oxy[1] = make_float4(ox[2], oy[2], ox[3], oy[3]);
        *reinterpret_cast<float4*>(vz + i) = make_float4(oz[0], oz[1], oz[2], oz[3]);
    }
}

}  // namespace sloth
