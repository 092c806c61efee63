// kernels.cuh -- the per-frame kernels of the sloth raster path (sm_100a).
//
// Frame pipeline (one stream, no host round-trips):
//   k_geom     one thread per triangle: coalesced 16+16+8 B loads, transform, bounds,
//              image-mode row stamps; sub-tile triangles are rasterised in place
//              (64-bit atomicMin into the key plane), larger ones are queued as
//              row-band work items with one warp-aggregated atomic per warp,
//              non-finite ones are queued for the brute-force kernel
//   k_walk     one warp per row-band item: 32 candidates per step, ballot-driven
//              row termination, atomicMin into the key plane
//   k_irregular one block per queued triangle: the reference's whole scan domain
//   k_resolve  key plane -> cell buffer (glyph + colour, double-cell write,
//              newline stamps), vectorised stores, resets the key plane
//   k_stampfix_* rare: order between a newline stamp and a wrapped fragment that
//              lands on the same cell
#pragma once
#include "raster_core.cuh"

namespace sloth {

static constexpr uint32_t TINY_ROWS = 4;   // rows x columns handled inside k_geom
static constexpr uint32_t TINY_COLS = 4;
static constexpr uint32_t TINY_MAX_STEPS = TINY_COLS + 3;  // per row before handing over to k_walk
static constexpr int ITEM_BITS = 37;       // packed queue counter: slots << 37 | items
static constexpr unsigned long long ITEM_MASK = (1ull << ITEM_BITS) - 1ull;

// Small per-frame device state (cleared with one memset per frame).
struct FrameAux {
    unsigned long long walk_counter;   // slots << 37 | items
    unsigned long long frag_counter;   // covered fragments (only when count_frags)
    uint32_t irr_count;
    uint32_t fix_count;
    uint32_t pad[10];
};

struct Queues {
    uint32_t* __restrict__ walk_tri;             // [n_tri]
    unsigned long long* __restrict__ walk_base;  // [n_tri] first item id of that triangle (ascending)
    uint32_t* __restrict__ irr_tri;              // [n_tri]
    uint32_t* __restrict__ rowbits;              // [(H+31)/32] image-mode: rows stamped by some triangle
    uint32_t* __restrict__ fix_rows;             // [H] rows whose column-1 cell needs the order check
    uint32_t* __restrict__ fix_tri;              // [H] triangle of the wrapped winner on that cell
    uint32_t* __restrict__ fix_newline;          // [H] 1 if a later stamp overrides it
    FrameAux* __restrict__ aux;
};

// rows [y0,y1) of a triangle get the '\n' marker at column 1 (rasterizer.rs:89-91).
SLOTH_DEV void stamp_rows(const FrameParams& p, const Queues& q, uint32_t y0, uint32_t y1)
{
    y0 = max(y0, p.row0);
    y1 = min(y1, p.row1);
    if (y0 >= y1) return;
    for (uint32_t w = y0 >> 5; w <= (y1 - 1) >> 5; ++w) {
        const uint32_t lo = max(y0, w << 5) & 31u, hi = (min(y1, (w + 1) << 5) - 1u) & 31u;
        const uint32_t m = (0xFFFFFFFFu >> (31u - hi)) & (0xFFFFFFFFu << lo);
        if ((q.rowbits[w] & m) != m) atomicOr(q.rowbits + w, m);
    }
}

__global__ void __launch_bounds__(256) k_geom(const __grid_constant__ FrameParams p, const Scene sc,
                                              unsigned long long* __restrict__ keys, const Queues q)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t walk_items = 0;   // > 0: queue this triangle for k_walk
    bool irregular = false;
    uint32_t nfrag = 0;

    if (t < p.n_tri) {
        float v[9];
        uint32_t rgb;
        load_tri(sc, t, v, rgb);
        Setup s;
        setup_tri(p, v, s);
        // destination rows are y (direct) or y+1 (wrapped): skip triangles outside the band
        const bool rows_ok = s.miny < s.maxy && s.miny < p.row1 && s.maxy + 1u > p.krow0;
        if (rows_ok) {
            if (p.image) stamp_rows(p, q, s.miny, s.maxy);
            if (s.minx < s.maxx) {
                const uint32_t rows = s.maxy - s.miny;
                const uint32_t tw = tight_width(s);
                if (!s.regular) {
                    irregular = true;
                } else if (rows <= TINY_ROWS && tw <= TINY_COLS) {
                    Shade sh;
                    bool have_sh = false;
                    bool bail = false;
                    for (uint32_t y = s.miny; y < s.maxy && !bail; ++y) {
                        const RowC rc = row_setup(s, y);
                        uint32_t steps = 0;
                        for (uint32_t x = s.minx; x < s.maxx; ++x) {
                            float w0, w1, w2;
                            edge_eval(s, rc, x, w0, w1, w2);
                            if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) {
                                if (!have_sh) {
                                    shade_setup(s, sh);
                                    have_sh = true;
                                }
                                emit_fragment(p, s, sh, t, x, y, w0, w1, w2, keys);
                                ++nfrag;
                            } else if (row_closed(s, w0, w1, w2)) {
                                break;
                            }
                            if (++steps > TINY_MAX_STEPS) {  // degenerate sliver: let k_walk redo it
                                bail = true;
                                break;
                            }
                        }
                    }
                    if (bail) {
                        nfrag = 0;  // k_walk re-emits the same keys (atomicMin is idempotent) and recounts
                        walk_items = (rows + walk_rows_per_item(tw) - 1u) / walk_rows_per_item(tw);
                    }
                } else {
                    walk_items = (rows + walk_rows_per_item(tw) - 1u) / walk_rows_per_item(tw);
                }
            }
        }
    }

    // ---- warp-aggregated queue allocation: one atomic per warp -----------------------
    const unsigned need = __ballot_sync(0xFFFFFFFFu, walk_items > 0);
    if (need) {
        uint32_t incl = walk_items;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += n;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        unsigned long long old = 0;
        if (lane == 0)
            old = atomicAdd(&q.aux->walk_counter, ((unsigned long long)__popc(need) << ITEM_BITS) | total);
        old = __shfl_sync(0xFFFFFFFFu, old, 0);
        if (walk_items > 0) {
            const uint32_t slot = (uint32_t)(old >> ITEM_BITS) + __popc(need & ((1u << lane) - 1u));
            q.walk_tri[slot] = t;
            q.walk_base[slot] = (old & ITEM_MASK) + (incl - walk_items);
        }
    }
    const unsigned irr = __ballot_sync(0xFFFFFFFFu, irregular);
    if (irr) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&q.aux->irr_count, (uint32_t)__popc(irr));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (irregular) q.irr_tri[base + __popc(irr & ((1u << lane) - 1u))] = t;
    }
    if (p.count_frags) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nfrag += __shfl_xor_sync(0xFFFFFFFFu, nfrag, d);
        if (lane == 0 && nfrag) atomicAdd(&q.aux->frag_counter, (unsigned long long)nfrag);
    }
}

// One warp per row-band item.  Lanes take 32 consecutive candidates of a row;
// a row ends when any lane sees a closing edge fail (everything right of that
// lane fails as well) or at the reference's maxx.
__global__ void __launch_bounds__(128) k_walk(const __grid_constant__ FrameParams p, const Scene sc,
                                              unsigned long long* __restrict__ keys, const Queues q)
{
    const unsigned long long packed = q.aux->walk_counter;
    const unsigned long long n_items = packed & ITEM_MASK;
    const uint32_t n_slots = (uint32_t)(packed >> ITEM_BITS);
    if (n_items == 0) return;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    uint32_t nfrag = 0;
    for (unsigned long long item = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
         item < n_items; item += n_warps) {
        // last slot whose base <= item
        uint32_t lo = 0, hi = n_slots;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (q.walk_base[mid] <= item) lo = mid; else hi = mid;
        }
        const uint32_t t = q.walk_tri[lo];
        const uint32_t band = (uint32_t)(item - q.walk_base[lo]);
        float v[9];
        uint32_t rgb;
        load_tri(sc, t, v, rgb);
        Setup s;
        setup_tri(p, v, s);
        Shade sh;
        shade_setup(s, sh);
        const uint32_t rpi = walk_rows_per_item(tight_width(s));
        const uint32_t y0 = s.miny + band * rpi;
        const uint32_t y1 = min(s.maxy, y0 + rpi);
        for (uint32_t y = y0; y < y1; ++y) {
            if (y + 1u < p.krow0 || y >= p.row1) continue;
            const RowC rc = row_setup(s, y);
            for (uint32_t xb = s.minx; xb < s.maxx; xb += 32u) {
                const uint32_t x = xb + lane;
                bool closed = false;
                if (x < s.maxx) {
                    float w0, w1, w2;
                    edge_eval(s, rc, x, w0, w1, w2);
                    if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) {
                        emit_fragment(p, s, sh, t, x, y, w0, w1, w2, keys);
                        ++nfrag;
                    } else {
                        closed = row_closed(s, w0, w1, w2);
                    }
                }
                if (__any_sync(0xFFFFFFFFu, closed)) break;
            }
        }
    }
    if (p.count_frags) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nfrag += __shfl_xor_sync(0xFFFFFFFFu, nfrag, d);
        if (lane == 0 && nfrag) atomicAdd(&q.aux->frag_counter, (unsigned long long)nfrag);
    }
}

// Triangles with NaN/inf/huge coordinates: no monotonicity argument applies,
// so every candidate of the reference's scan domain is evaluated.
__global__ void __launch_bounds__(256) k_irregular(const __grid_constant__ FrameParams p, const Scene sc,
                                                   unsigned long long* __restrict__ keys, const Queues q)
{
    const uint32_t n = q.aux->irr_count;
    uint32_t nfrag = 0;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint32_t t = q.irr_tri[i];
        float v[9];
        uint32_t rgb;
        load_tri(sc, t, v, rgb);
        Setup s;
        setup_tri(p, v, s);
        Shade sh;
        shade_setup(s, sh);
        const unsigned long long w = s.maxx - s.minx, h = s.maxy - s.miny;  // both > 0 (checked in k_geom)
        for (unsigned long long c = threadIdx.x; c < w * h; c += blockDim.x) {
            const uint32_t y = s.miny + (uint32_t)(c / w), x = s.minx + (uint32_t)(c % w);
            const RowC rc = row_setup(s, y);
            float w0, w1, w2;
            edge_eval(s, rc, x, w0, w1, w2);
            if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) {
                emit_fragment(p, s, sh, t, x, y, w0, w1, w2, keys);
                ++nfrag;
            }
        }
    }
    if (p.count_frags && nfrag) atomicAdd(&q.aux->frag_counter, (unsigned long long)nfrag);
}

// ---------------------------------------------------------------------------------
// Resolve: key plane -> cells.
// ---------------------------------------------------------------------------------
SLOTH_DEV uint32_t key_tri(unsigned long long k) { return ((uint32_t)k >> 5) & MAX_TRIS; }

SLOTH_DEV uint32_t cell_of(const FrameParams& p, const Scene& sc, unsigned long long key)
{
    const uint32_t g = (uint32_t)key & 15u;
    const uint32_t rgb = __float_as_uint(__ldg(&sc.c[key_tri(key)].y));
    return (uint32_t)(uint8_t)p.glyph[g] | (rgb << 8);
}

// Column-1 cell of a stamped row on which a (wrapped) fragment also landed: the
// stamp of triangle T is written after all fragments of triangles <= T, so the
// '\n' stays unless the fragment's triangle is later than every stamping triangle.
// That needs max{T : row in T's y-range}; queue the row, k_stampfix_scan decides.
SLOTH_DEV void queue_fix(const Queues& q, uint32_t row, uint32_t tri)
{
    const uint32_t i = atomicAdd(&q.aux->fix_count, 1u);
    q.fix_rows[i] = row;
    q.fix_tri[i] = tri;
    q.fix_newline[i] = 0u;
}

// W even: ids are even, one key slot owns cells (2k, 2k+1) of its row.
__global__ void __launch_bounds__(256) k_resolve_even(const __grid_constant__ FrameParams p, const Scene sc,
                                                      unsigned long long* __restrict__ keys, const Queues q,
                                                      uint32_t* __restrict__ cells, uint32_t n_slots,
                                                      uint32_t n_tail)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t blank = (uint32_t)' ';
    if (i < n_slots) {
        const unsigned long long key = keys[i];
        keys[i] = KEY_EMPTY;  // the key plane is clean again for the next frame
        uint32_t c0 = blank, c1 = blank;
        if (key != KEY_EMPTY) c0 = c1 = cell_of(p, sc, key);
        const uint32_t row = p.row0 + i / p.KW, kx = i % p.KW;
        if (p.image && kx == 0 && ((q.rowbits[row >> 5] >> (row & 31u)) & 1u)) {
            if (key != KEY_EMPTY) queue_fix(q, row, key_tri(key));  // provisional: fragment stays
            else c1 = (uint32_t)'\n';
        }
        reinterpret_cast<uint2*>(cells)[i] = make_uint2(c0, c1);
    } else if (i < n_slots + n_tail) {
        cells[2u * n_slots + (i - n_slots)] = blank;  // image-mode tail, context.rs:38-39
    }
}

SLOTH_DEV bool frag_later(const FrameParams& p, unsigned long long ka, uint32_t ida, unsigned long long kb,
                          uint32_t idb)
{
    // sequential time of the winning fragment: (triangle, source row, source x)
    const uint32_t ta = key_tri(ka), tb = key_tri(kb);
    if (ta != tb) return ta > tb;
    const uint32_t da = ((uint32_t)ka >> 4) & 1u, db = ((uint32_t)kb >> 4) & 1u;
    const uint32_t ya = ida / p.W - (1u - da), yb = idb / p.W - (1u - db);
    if (ya != yb) return ya > yb;
    const uint32_t xa = (ida % p.W + (1u - da) * p.W) >> 1, xb = (idb % p.W + (1u - db) * p.W) >> 1;
    return xa > xb;
}

// W odd: key slot == id; a cell can be an id itself and the id+1 copy of its
// left neighbour (which may be the last cell of the row above) -- the later
// write wins.  One thread per cell.  `halo` = key slots that
// precede this band's first cell in the key plane (one extra row when the band
// does not start at row 0, so that the copy from the row above is seen).
__global__ void __launch_bounds__(256) k_resolve_odd(const __grid_constant__ FrameParams p, const Scene sc,
                                                     const unsigned long long* __restrict__ keys, const Queues q,
                                                     uint32_t* __restrict__ cells, uint32_t n_cells,
                                                     uint32_t n_tail, uint32_t halo)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t blank = (uint32_t)' ';
    if (i < n_cells) {
        const uint32_t id = p.row0 * p.W + i;  // global cell index
        const unsigned long long ka = keys[halo + i];
        const unsigned long long kb = (halo + i) > 0 ? keys[halo + i - 1] : KEY_EMPTY;
        unsigned long long kw = KEY_EMPTY;
        if (ka != KEY_EMPTY && kb != KEY_EMPTY) kw = frag_later(p, ka, id, kb, id - 1u) ? ka : kb;
        else if (ka != KEY_EMPTY) kw = ka;
        else if (kb != KEY_EMPTY) kw = kb;
        uint32_t c = blank;
        if (kw != KEY_EMPTY) c = cell_of(p, sc, kw);
        const uint32_t row = id / p.W, col = id % p.W;
        if (p.image && col == 1 && ((q.rowbits[row >> 5] >> (row & 31u)) & 1u)) {
            if (kw != KEY_EMPTY) queue_fix(q, row, key_tri(kw));
            else c = (uint32_t)'\n';
        }
        cells[i] = c;
    } else if (i < n_cells + n_tail) {
        cells[i] = blank;
    }
}

__global__ void __launch_bounds__(256) k_clear_keys_odd(unsigned long long* __restrict__ keys, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = KEY_EMPTY;
}

// For every queued row: does a triangle with index >= the fragment's triangle
// stamp that row?  Only the y-range of each triangle is needed.
__global__ void __launch_bounds__(256) k_stampfix_scan(const __grid_constant__ FrameParams p, const Scene sc,
                                                       const Queues q)
{
    const uint32_t n_fix = q.aux->fix_count;
    if (n_fix == 0) return;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < p.n_tri; t += gridDim.x * blockDim.x) {
        float v[9];
        uint32_t rgb;
        load_tri(sc, t, v, rgb);
        Setup s;
        setup_tri(p, v, s);
        if (s.miny >= s.maxy) continue;
        for (uint32_t i = 0; i < n_fix; ++i) {
            const uint32_t row = q.fix_rows[i];
            if (t >= q.fix_tri[i] && row >= s.miny && row < s.maxy) q.fix_newline[i] = 1u;
        }
    }
}

__global__ void k_stampfix_apply(const __grid_constant__ FrameParams p, const Queues q,
                                 uint32_t* __restrict__ cells)
{
    const uint32_t n_fix = q.aux->fix_count;
    for (uint32_t i = threadIdx.x; i < n_fix; i += blockDim.x)
        if (q.fix_newline[i]) cells[(q.fix_rows[i] - p.row0) * p.W + 1u] = (uint32_t)'\n';
}

// Context.z_buffer (context.rs:17) reconstructed from the key plane, before resolve.
__global__ void __launch_bounds__(256) k_zbuffer(const __grid_constant__ FrameParams p,
                                                 const unsigned long long* __restrict__ keys,
                                                 float* __restrict__ z, uint32_t n_cells)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    float out = 3.40282347e+38f;
    const bool is_slot = (p.XS == 2u) || ((i & 1u) == 0u);
    if (is_slot) {
        const unsigned long long key = keys[p.XS == 2u ? i : (i >> 1)];
        if (key != KEY_EMPTY) {
            const uint32_t ord = (uint32_t)(key >> 32);
            out = __uint_as_float((ord & 0x80000000u) ? (ord & 0x7FFFFFFFu) : ~ord);
        }
    }
    z[i] = out;
}

// Scene upload: raw soup (9 floats + 3 bytes per triangle) -> the three resident streams.
__global__ void __launch_bounds__(256) k_pack_scene(const float* __restrict__ xyz, const uint8_t* __restrict__ rgb,
                                                    uint32_t n, float4* __restrict__ a, float4* __restrict__ b,
                                                    float2* __restrict__ c)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* v = xyz + (size_t)t * 9;
    a[t] = make_float4(v[0], v[1], v[2], v[3]);
    b[t] = make_float4(v[4], v[5], v[6], v[7]);
    const uint32_t col = (uint32_t)rgb[(size_t)t * 3] | ((uint32_t)rgb[(size_t)t * 3 + 1] << 8) |
                         ((uint32_t)rgb[(size_t)t * 3 + 2] << 16);
    c[t] = make_float2(v[8], __uint_as_float(col));
}

}  // namespace sloth
