// tile_kernels.cuh -- the binned path for triangles larger than the in-register footprints (sm_100a).
//
// The geometry kernel queues every triangle whose scan domain exceeds 8 x 8 candidates (Queues::walk_tri).  Those
// triangles -- few, but they carry almost all fragments of a model that fills the screen (config 3, "suzy suzy" at
// 4K: 1,088 triangles, 3.0 M fragments) -- are rasterised per screen tile instead of per triangle:
//
//   k_bin_count   warp per queued triangle: transform + bounds (setup_tri), then the proof that nothing can be
//                 covered right of the bounding box: on every row the first candidate column right of the box must
//                 fail a closing edge (row_closed; the same argument that ends rows in the footprint tiers and in the
//                 walk kernel).  A triangle that passes is binned over the TILE_W x TILE_H tiles of its tight domain:
//                 one warp-aggregated atomic reserves its pairs, one atomicAdd per overlapped tile counts them.  A
//                 triangle that does not (slivers), or whose pairs do not fit the pair buffer, stays with the walk
//                 kernel (k_tail), which then skips everything that was binned.
//   k_bin_scan    one block: exclusive scan of the tile counters -> pair offsets, and the compacted list of non-empty
//                 tiles.
//   k_bin_fill    warp per binned triangle: writes its index into the pair list of every tile it overlaps.
//   k_tile        persistent blocks over the non-empty tiles, thread = candidate pixel, warp = one row of the tile:
//                 the pixel's depth / glyph / order key lives in a register (the same 64-bit key as everywhere
//                 else); the triangle's transform, 1/area and normal were computed once per triangle by k_bin_count
//                 (64-byte TileTri records); the block stages the records of its tile's pairs in shared memory, 64
//                 at a time with all 256 threads fetching (one memory round trip per batch), and every warp skips,
//                 as a whole, the triangles whose domain misses its 32 pixels; every pixel takes the minimum over the triangles that cover it and
//                 the tile leaves with ONE 64-bit atomicMin per touched pixel (other paths write the same key plane,
//                 and a wrapped pixel shares its slot with the next row's).
//
// Exactness: candidates are tested with the reference's own per-pixel edge functions (row_setup / edge_eval: no
// stepping), inside the reference's scan domain [minx, maxx) x [miny, maxy) cut to the columns left of the closing
// column, which the proof above shows is all that can be covered.
#pragma once
#include "kernels.cuh"

namespace sloth {

static constexpr uint32_t TILE_W = 32, TILE_H = 8;           // candidates (pixels) per tile: one block of 256 threads

struct __align__(16) TileTri {   // one queued triangle, set up once per frame (64 bytes)
    float x1, y1, x2, y2;      // words 0-3
    float x3, y3, z1, a;       // words 4-7
    float k, dz1, dz2, padf;   // words 8-11
    uint32_t xr;               // words 12-15: minx | maxx << 16   (maxx already cut to the closing column)
    uint32_t yr;               //              miny | maxy << 16
    uint32_t tri;
    uint32_t pad;
};
static_assert(sizeof(TileTri) == 64, "TileTri is read as four 16-byte loads");

struct TileAux {
    unsigned long long pairs_total;   // (tile, triangle) pairs reserved so far
    uint32_t n_tiles_used;            // non-empty tiles (k_bin_scan)
    uint32_t binned_tris;             // statistics
    uint32_t pad[12];
};

struct TileState {
    uint32_t* __restrict__ count;       // [n_tiles + 1]  pairs per tile, then (k_bin_scan) exclusive offsets
    uint32_t* __restrict__ cursor;      // [n_tiles]      fill cursors
    uint4* __restrict__ used;           // [n_tiles]      compacted non-empty tiles: (tile, first pair, last pair, 0)
    struct TileTri* __restrict__ pairs; // [pair_cap]     triangle records, grouped by tile (a copy per pair: one contiguous
                                        //                run of 64-byte records per tile, no indirection when rasterising)
    unsigned long long* __restrict__ info;   // [n_tri] per queue slot: tx0 | tx1 << 16 | ty0 << 32 | ty1 << 48, or ~0 = not binned
    struct TileTri* __restrict__ setup;      // [n_tri] per queue slot: the triangle set up once for all its tiles
    TileAux* __restrict__ aux;
    uint32_t tiles_x, tiles_y, pair_cap;
};

static constexpr unsigned long long TILE_NOT_BINNED = ~0ull;

// ---- J1: setup + bin ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bin_count(const __grid_constant__ FrameParams p, const Scene sc, const Queues q,
                                                   const TileState ts)
{
    const uint32_t n_slots = (uint32_t)(q.aux->walk_counter >> ITEM_BITS);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); slot < n_slots; slot += n_warps) {
        const uint32_t t = q.walk_tri[slot];
        float v[9];
        load_tri(sc, t, v);
        Setup s;
        setup_tri(p, v, s);
        // first candidate column right of the bounding box (if the scan domain reaches that far)
        const uint32_t f = __float2uint_rz(floorf(s.mx0));
        const uint32_t xc = f >= s.maxx ? s.maxx : f + 1u;      // tight domain: [minx, xc)
        bool open = false;
        if (xc < s.maxx) {
            for (uint32_t y = s.miny + lane; y < s.maxy; y += 32u) {
                const RowC rc = row_setup(s, y);
                float w0, w1, w2;
                edge_eval(s, rc, xc, w0, w1, w2);
                if (!row_closed(s, w0, w1, w2)) open = true;
            }
        }
        open = __any_sync(0xFFFFFFFFu, open);
        unsigned long long info = TILE_NOT_BINNED;
        if (!open && xc > s.minx && s.regular) {
            const uint32_t tx0 = s.minx / TILE_W, tx1 = (xc - 1u) / TILE_W;
            const uint32_t ty0 = s.miny / TILE_H, ty1 = (s.maxy - 1u) / TILE_H;
            const uint32_t n = (tx1 - tx0 + 1u) * (ty1 - ty0 + 1u);
            unsigned long long before = 0;
            if (lane == 0) before = atomicAdd(&ts.aux->pairs_total, (unsigned long long)n);
            before = __shfl_sync(0xFFFFFFFFu, before, 0);
            if (before + n <= ts.pair_cap) {
                info = (unsigned long long)tx0 | ((unsigned long long)tx1 << 16) | ((unsigned long long)ty0 << 32) |
                       ((unsigned long long)ty1 << 48);
                const uint32_t nx = tx1 - tx0 + 1u;
                for (uint32_t i = lane; i < n; i += 32u) atomicAdd(ts.count + (ty0 + i / nx) * ts.tiles_x + tx0 + i % nx, 1u);
                if (lane == 0) atomicAdd(&ts.aux->binned_tris, 1u);
            }
        }
        if (lane == 0) {
            ts.info[slot] = info;
            if (info != TILE_NOT_BINNED) {
                Shade sh;
                shade_setup(s, sh);
                TileTri o;
                o.x1 = s.x1; o.y1 = s.y1; o.x2 = s.x2; o.y2 = s.y2; o.x3 = s.x3; o.y3 = s.y3;
                o.z1 = s.z1; o.a = sh.a; o.k = sh.k; o.dz1 = sh.dz1; o.dz2 = sh.dz2;
                o.xr = s.minx | (xc << 16);   // the closing column and beyond cannot be covered
                o.yr = s.miny | (s.maxy << 16);
                o.tri = t;
                o.padf = 0.0f; o.pad = 0u;
                ts.setup[slot] = o;
            }
        }
    }
}

// one block of 1024 threads: exclusive scan of count[0..n_tiles) in place (count[n_tiles] = total), cursors zeroed,
// non-empty tiles compacted into used[] (ascending).  Warp w owns the contiguous range [w R, (w+1) R) and walks it 32
// tiles at a time with shuffle scans (coalesced, no block barrier inside the loop); one barrier to scan the 32 range
// totals, then a second walk adds the range offsets.
__global__ void __launch_bounds__(1024) k_bin_scan(const TileState ts)
{
    __shared__ uint32_t range_pairs[32], range_used[32];
    const uint32_t n_tiles = ts.tiles_x * ts.tiles_y;
    if (ts.aux->binned_tris == 0u) {   // nothing was binned (every counter is still zero): k_tile has no tile to visit
        if (threadIdx.x == 0) ts.aux->n_tiles_used = 0u;
        return;
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t R = ((n_tiles + 31u) / 32u + 31u) & ~31u;   // range length, a multiple of 32
    const uint32_t r0 = warp * R, r1 = min(n_tiles, r0 + R);
    uint32_t run_p = 0, run_u = 0;
    for (uint32_t base = r0; base < r1; base += 256u) {   // eight independent loads in flight per lane
        uint32_t v[8];
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k) {
            const uint32_t i = base + 32u * k + lane;
            v[k] = i < r1 ? ts.count[i] : 0u;
        }
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k) {
            run_p += v[k];
            run_u += v[k] ? 1u : 0u;
        }
    }
    run_p = __reduce_add_sync(0xFFFFFFFFu, run_p);
    run_u = __reduce_add_sync(0xFFFFFFFFu, run_u);
    if (lane == 0) { range_pairs[warp] = run_p; range_used[warp] = run_u; }
    __syncthreads();
    if (warp == 0) {
        uint32_t a = range_pairs[lane], b = range_used[lane];
        const uint32_t ta = a, tb = b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, a, d), y = __shfl_up_sync(0xFFFFFFFFu, b, d);
            if ((int)lane >= d) { a += x; b += y; }
        }
        range_pairs[lane] = a - ta; range_used[lane] = b - tb;   // exclusive over ranges
        if (lane == 31u) { ts.count[n_tiles] = a; ts.aux->n_tiles_used = b; }
    }
    __syncthreads();
    uint32_t pre_p = range_pairs[warp], pre_u = range_used[warp];
    for (uint32_t base8 = r0; base8 < r1; base8 += 256u) {
      uint32_t v8[8];
#pragma unroll
      for (uint32_t k = 0; k < 8u; ++k) {
          const uint32_t i = base8 + 32u * k + lane;
          v8[k] = i < r1 ? ts.count[i] : 0u;   // (L2 hits: the first walk just read them)
      }
#pragma unroll
      for (uint32_t k = 0; k < 8u; ++k) {
        const uint32_t i = base8 + 32u * k + lane;
        const uint32_t v = v8[k];
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((int)lane >= d) inc += x;
        }
        const unsigned nz = __ballot_sync(0xFFFFFFFFu, v != 0u);
        if (i < r1) {
            const uint32_t off = pre_p + inc - v;
            ts.count[i] = off;
            ts.cursor[i] = 0u;
            if (v) ts.used[pre_u + __popc(nz & ((1u << lane) - 1u))] = make_uint4(i, off, off + v, 0u);
        }
        pre_p += __shfl_sync(0xFFFFFFFFu, inc, 31);
        pre_u += __popc(nz);
      }
    }
}

__global__ void __launch_bounds__(256) k_bin_fill(const Queues q, const TileState ts)
{
    const uint32_t n_slots = (uint32_t)(q.aux->walk_counter >> ITEM_BITS);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); slot < n_slots; slot += n_warps) {
        const unsigned long long info = ts.info[slot];
        if (info == TILE_NOT_BINNED) continue;
        const uint32_t tx0 = (uint32_t)info & 0xFFFFu, tx1 = (uint32_t)(info >> 16) & 0xFFFFu;
        const uint32_t ty0 = (uint32_t)(info >> 32) & 0xFFFFu, ty1 = (uint32_t)(info >> 48) & 0xFFFFu;
        const uint32_t nx = tx1 - tx0 + 1u, n = nx * (ty1 - ty0 + 1u);
        const uint4* src = reinterpret_cast<const uint4*>(ts.setup + slot);
        const uint4 A = src[0], B = src[1], C = src[2], D = src[3];
        for (uint32_t i = lane; i < n; i += 32u) {
            const uint32_t tile = (ty0 + i / nx) * ts.tiles_x + tx0 + i % nx;
            uint4* dst = reinterpret_cast<uint4*>(ts.pairs + ts.count[tile] + atomicAdd(ts.cursor + tile, 1u));
            dst[0] = A; dst[1] = B; dst[2] = C; dst[3] = D;
        }
    }
}

// ---- J2: per-tile raster -----------------------------------------------------------------------------------
static constexpr uint32_t TILE_BATCH = 64;   // triangle records staged per round (4 KB of shared memory)

SLOTH_DEV void tile_stage(uint4* dst, const TileTri* src, uint32_t nb)
{
    // all 256 threads copy: thread i moves 16-byte word i of the (contiguous) run of nb records
    if ((threadIdx.x >> 2) < nb)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + threadIdx.x)),
                     "l"(reinterpret_cast<const uint4*>(src) + threadIdx.x)
                     : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(TILE_W * TILE_H, 4) k_tile(const __grid_constant__ FrameParams p, const Scene sc,
                                                             unsigned long long* __restrict__ keys, const Queues q,
                                                             const TileState ts)
{
    // two staging buffers: while a tile is rasterised from one, the first batch of the block's next tile lands in
    // the other (cp.async), so the per-tile memory round trips overlap with arithmetic
    __shared__ uint4 batch[2][TILE_BATCH * 4];   // TileTri records as four 16-byte words each
    const uint32_t n_used = ts.aux->n_tiles_used;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t nfrag = 0;
    uint32_t buf = 0;
    uint4 U = make_uint4(0u, 0u, 0u, 0u);
    if (blockIdx.x < n_used) {
        U = ts.used[blockIdx.x];
        tile_stage(batch[0], ts.pairs + U.y, min(TILE_BATCH, U.z - U.y));
    }
    for (uint32_t u = blockIdx.x; u < n_used; u += gridDim.x) {
        const uint32_t tile = U.x, first = U.y, last = U.z;
        // the next tile of this block: its descriptor now, its first batch as soon as the other buffer is free
        const uint32_t un = u + gridDim.x;
        uint4 Un = make_uint4(0u, 0u, 0u, 0u);
        if (un < n_used) Un = ts.used[un];
        const uint32_t tx = tile % ts.tiles_x, ty = tile / ts.tiles_x;
        const uint32_t x0w = tx * TILE_W;                                 // this warp's candidates: columns x0w .. x0w + 31,
        const uint32_t x = x0w + lane, y = ty * TILE_H + warp;            // row y
        const float px = (float)x, py = (float)y;
        unsigned long long best = KEY_EMPTY;
        for (uint32_t b0 = first; b0 < last; b0 += TILE_BATCH) {
            const uint32_t nb = min(TILE_BATCH, last - b0);
            if (b0 != first) {   // further batches of a crowded tile: staged on demand
                __syncthreads();
                tile_stage(batch[buf], ts.pairs + b0, nb);
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            if (b0 + TILE_BATCH >= last && un < n_used) tile_stage(batch[buf ^ 1u], ts.pairs + Un.y, min(TILE_BATCH, Un.z - Un.y));
            const uint4* cur = batch[buf];
            for (uint32_t j = 0; j < nb; ++j) {
                const uint4 D = cur[j * 4u + 3u];
                const uint32_t minx = D.x & 0xFFFFu, maxx = D.x >> 16, miny = D.y & 0xFFFFu, maxy = D.y >> 16;
                // the whole warp skips a triangle whose domain misses its 32 candidates
                if (y < miny || y >= maxy || x0w + TILE_W <= minx || x0w >= maxx) continue;
                if (x < minx || x >= maxx) continue;
                const uint4 A = cur[j * 4u], B = cur[j * 4u + 1u], C = cur[j * 4u + 2u];
                const float x1 = __uint_as_float(A.x), y1 = __uint_as_float(A.y), x2 = __uint_as_float(A.z), y2 = __uint_as_float(A.w);
                const float x3 = __uint_as_float(B.x), y3 = __uint_as_float(B.y), z1 = __uint_as_float(B.z), ia = __uint_as_float(B.w);
                const float kk = __uint_as_float(C.x), dz1 = __uint_as_float(C.y), dz2 = __uint_as_float(C.z);
                // orient() with the reference's base vertices (rasterizer.rs:72-74), no stepping
                const float w0 = sub(mul(sub(x3, x2), sub(py, y2)), mul(sub(y3, y2), sub(px, x2)));
                const float w1 = sub(mul(sub(x1, x3), sub(py, y3)), mul(sub(y1, y3), sub(px, x3)));
                const float w2 = sub(mul(sub(x2, x1), sub(py, y1)), mul(sub(y2, y1), sub(px, x1)));
                if (!(w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f)) continue;
                ++nfrag;
                const float z = add(z1, mul(ia, add(mul(w1, dz1), mul(w2, dz2))));
                if (!(z < 3.40282347e+38f)) continue;   // z_buffer starts at f32::MAX; NaN never wins
                const float shade = mul(kk, add(add(w0, w1), w2));
                const uint32_t g = glyph_index(p, shade);
                uint32_t zb = __float_as_uint(z);
                if (zb == 0x80000000u) zb = 0u;
                const uint32_t ord = (zb & 0x80000000u) ? ~zb : (zb | 0x80000000u);
                const uint32_t direct = x * p.XS < p.KW ? 1u : 0u;
                const unsigned long long key = ((unsigned long long)ord << 32) | (unsigned long long)((D.z << 5) | (direct << 4) | g);
                best = key < best ? key : best;
            }
        }
        if (best != KEY_EMPTY) {   // the single write-out of this pixel (same slot arithmetic as emit_fragment)
            const uint32_t kx = x * p.XS;
            const uint32_t direct = kx < p.KW ? 1u : 0u;
            const uint32_t row = y + 1u - direct;
            if (row >= p.krow0 && row < p.row1 && !(p.debug & 1u)) atomicMin(keys + (y * p.KW + kx - p.krow0 * p.KW), best);
        }
        buf ^= 1u;
        U = Un;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (p.count_frags) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nfrag += __shfl_xor_sync(0xFFFFFFFFu, nfrag, d);
        if (lane == 0 && nfrag) atomicAdd(&q.aux->frag_counter, (unsigned long long)nfrag);
    }
}

}  // namespace sloth
