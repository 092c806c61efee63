// kernels.cuh -- the per-frame kernels of the sloth raster path (sm_100a).
//
// Frame pipeline (no host round-trips, 3 launches + one small memset; in batches the geometry of
// frame k+1 runs on one stream while the resolve of frame k runs on a second one):
//   k_geom3    persistent warps over the triangle stream: coalesced 16+16+4 B loads with register
//              prefetch, transform, bounds, image-mode row stamps, back-face proof; triangles up to
//              8x8 candidates are rasterised in place (64-bit atomicMin into the key plane), larger
//              ones are queued as row-band work items, non-finite ones for the brute-force pass
//   k_tail     walk part: one warp per row-band item, ballot-terminated rows;
//              irregular part: the reference's whole scan domain for NaN/inf/huge triangles
//   k_resolve  key plane -> cell buffer (glyph + colour, double-cell write, newline stamps and
//              their order against wrapped fragments), 8-byte stores, resets the key plane
#pragma once
#include "raster_core.cuh"

namespace sloth {

static constexpr int ITEM_BITS = 37;       // packed queue counter: slots << 37 | items
static constexpr unsigned long long ITEM_MASK = (1ull << ITEM_BITS) - 1ull;

// Small per-frame device state (cleared with one memset per frame, together with rowmax).
struct FrameAux {
    unsigned long long walk_counter;   // slots << 37 | items
    unsigned long long frag_counter;   // covered fragments (only when count_frags)
    uint32_t irr_count;
    uint32_t stamp_exact;              // newline-vs-wrapped-fragment order decided by the exact check
    uint32_t chunks_done;              // chunks of 32 triangles k_geom3 processed, i.e. not band-culled (only when count_frags)
    uint32_t pad[9];
};

// Lengths of the two super-chunk lists of a frame (k_super_cert); cleared on the transform stream, apart from FrameAux.
struct ConeCounts {
    uint32_t n_live;   // super-chunks left for k_tri (Queues::live_sc holds SC_CHUNKS chunk indices for each)
    uint32_t n_skip;   // entries of Queues::skip_sc: certified back-facing, only their row stamps remain
};

struct Queues {
    uint32_t* __restrict__ walk_tri;             // [n_tri]
    unsigned long long* __restrict__ walk_base;  // [n_tri] first item id of that triangle (ascending)
    uint32_t* __restrict__ irr_tri;              // [n_tri]
    // image mode: rowmax[y] = 1 + index of the last 32-triangle chunk that stamps row y with the
    // '\n' marker (rasterizer.rs:89-91), 0 = row never stamped.  [H + 64]
    uint32_t* __restrict__ rowmax;
    FrameAux* __restrict__ aux;
    uint32_t* __restrict__ live_sc;              // [chunks] k_tri<CONE>'s work list: the chunks behind the last full super-chunk
                                                 // (static), then SC_CHUNKS entries per super-chunk k_super_cert could not certify
    uint32_t* __restrict__ skip_sc;              // [super-chunks] the certified ones: only their row stamps remain
    ConeCounts* __restrict__ cone_cnt;
};

// ---------------------------------------------------------------------------------
// TMA feed for k_geom3: the geometry streams are also kept as 1152-byte chunks of 32 triangles
// ([32 x float4 A][32 x float4 B][32 x float v3.z], same fields as Scene), so one elected
// lane moves a whole chunk with a single cp.async.bulk into a per-warp 3-stage ring and the
// warp waits on the stage's mbarrier: no per-lane address arithmetic or predicates, no
// registers held by loads in flight, two chunks of look-ahead.
// ---------------------------------------------------------------------------------
static constexpr uint32_t CHUNK_FLOATS = 288;               // 128 + 128 + 32
static constexpr uint32_t CHUNK_BYTES = CHUNK_FLOATS * 4;   // 1152
static constexpr uint32_t TMA_STAGES = 3;

SLOTH_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

SLOTH_DEV void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

SLOTH_DEV void tma_load_chunk(float* dst, const float* src, unsigned long long* bar)
{
    // generic-proxy reads of this stage (its previous use) are ordered before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(CHUNK_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(CHUNK_BYTES), "r"(smem_u32(bar))
                 : "memory");
}

SLOTH_DEV void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
}

struct TmaRing {
    float stage[TMA_STAGES][CHUNK_FLOATS];
    unsigned long long full[TMA_STAGES];
    unsigned long long pad;
};

// ---------------------------------------------------------------------------------
// k_geom3: the geometry kernel, warp-autonomous (no block barriers).
//
// Persistent warps walk the triangle stream in chunks of 32 triangles (lane = triangle),
// chunk i on warp i mod n_warps (or in short runs of consecutive chunks, see capi.cu):
//   A  prefetched 16+16+8 B loads, x'/y' transform (z' is not needed unless a
//      fragment is found), bounds, row stamps, classification.
//   cull  a triangle whose computed orientation is negative by more than the
//      rounding slack of the edge functions cannot cover any candidate (proof at
//      backface_proven); closed meshes lose half their triangles here, and whole
//      warps skip phase B.
//   B  every remaining triangle whose scan domain fits 2 rows x (2 columns + 1
//      closing column) is evaluated in registers, all lanes in lockstep: the edge
//      functions are separable, w_i(x,y) = c_i(y) - g_i(x), so the 6 candidates
//      cost 2x3 row terms + 3x3 column terms + 18 subtractions.  Nothing can be
//      covered right of a column where a closing edge already fails (row_closed);
//      if that cannot be shown inside the footprint the triangle goes to k_walk.
//   C  covering triangles are parked in a per-warp shared-memory ring; whenever 32
//      are queued they are emitted with all lanes busy: lane = triangle, full
//      re-transform (now with z'), normal and 1/area once, then per footprint bit
//      one edge evaluation, depth, glyph and a 64-bit atomicMin into the key plane.
// ---------------------------------------------------------------------------------
static constexpr uint32_t G3_WARPS = 8;          // warps per block
#ifndef G3_BLOCKS_PER_SM
#define G3_BLOCKS_PER_SM 3
#endif
static constexpr uint32_t G3_BATCH_MAX = 16;     // upper limit of the consecutive-chunks-per-turn knob (default 1, see capi.cu)
static constexpr uint32_t G3_RING = 64;          // per-warp ring of covering triangles (power of two)

struct G3Queue {
    float raw[9][G3_RING];        // object-space vertices
    uint32_t tri[G3_RING];
    uint32_t xy[G3_RING];         // minx | miny << 16
    uint32_t mask_lo[G3_RING];    // footprint bits: bit = row*3 + col (2x3 tier) or row*8 + col (8x8 tier,
    uint32_t mask_hi[G3_RING];    // flagged by bit 31 of tri[])
};

// No candidate of the scan domain can pass all three edge tests when the computed
// orientation A_c = fl(fl(dy2*dx1) - fl(dx2*dy1)) (bit-identical to the reference's
// orient(v1,v2,v3)) satisfies  A_c < -2^-18 * L * D,  L = larger bbox side,
// D = largest |candidate - vertex| distance along an axis.  Proof sketch (u = 2^-24,
// regular triangle, DESIGN.md has the full argument): each computed edge value has
// the sign of E1(1+t1) - E2(1+t2) with |t| <= 3.01u, so "all three >= 0" implies
// A_exact >= -3.01u * sum(|E1|+|E2|) >= -18.1u L D, while A_exact <= A_c(1-u) + 6.1u L^2
// and L <= 2D; 2^-18 = 64u leaves a factor > 2 for the rounding of the bound itself.
SLOTH_DEV bool backface_proven(const FrameParams& p, float dx1, float dy1, float dx2, float dy2, float mn0,
                               float mx0, float mn1, float mx1)
{
    const float area = sub(mul(dy2, dx1), mul(dx2, dy1));
    const float L = fmaxf(sub(mx0, mn0), sub(mx1, mn1));
    const float D = fmaxf(add(fmaxf(fabsf(mn0), fabsf(mx0)), p.wm1), add(fmaxf(fabsf(mn1), fabsf(mx1)), p.hm1));
    const float T = mul(mul(L, D), 3.814697265625e-06f);   // 2^-18 L D
    // T > 1e-30 keeps the bound itself (and the edge products it stands for) out of the
    // denormal range, where the relative-error argument would not hold.
    return T > 1e-30f && area < -T;
}

// Emit the fragments of `count` queued triangles starting at ring position `head`:
// lane = triangle; the per-triangle work (full transform now including z', normal,
// 1/area) is done once, then each footprint bit costs one edge evaluation + depth +
// glyph + atomicMin.
SLOTH_DEV void g3_emit(const FrameParams& p, const G3Queue& wq, uint32_t head, uint32_t count, uint32_t lane,
                       unsigned long long* __restrict__ keys, uint32_t& nfrag_count)
{
    if (lane >= count) return;
    const uint32_t slot = (head + lane) & (G3_RING - 1u);
    float v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = wq.raw[k][slot];
    // same operations on the same inputs as phase A: bit-identical x', y'
    Setup s;
    s.x1 = xform_row(p.m + 0, v[0], v[1], v[2]); s.y1 = xform_row(p.m + 4, v[0], v[1], v[2]);
    s.z1 = xform_row(p.m + 8, v[0], v[1], v[2]);
    s.x2 = xform_row(p.m + 0, v[3], v[4], v[5]); s.y2 = xform_row(p.m + 4, v[3], v[4], v[5]);
    s.z2 = xform_row(p.m + 8, v[3], v[4], v[5]);
    s.x3 = xform_row(p.m + 0, v[6], v[7], v[8]); s.y3 = xform_row(p.m + 4, v[6], v[7], v[8]);
    s.z3 = xform_row(p.m + 8, v[6], v[7], v[8]);
    s.dx0 = sub(s.x3, s.x2); s.dy0 = sub(s.y3, s.y2);
    s.dx1 = sub(s.x1, s.x3); s.dy1 = sub(s.y1, s.y3);
    s.dx2 = sub(s.x2, s.x1); s.dy2 = sub(s.y2, s.y1);
    Shade sh;
    shade_setup(s, sh);
    const uint32_t xy = wq.xy[slot], tri_word = wq.tri[slot];
    const uint32_t tri = tri_word & 0x7FFFFFFFu;
    const bool wide = (tri_word >> 31) != 0u;   // 8x8 tier: bit = row*8 + col; else bit = row*3 + col
    unsigned long long mask = ((unsigned long long)wq.mask_hi[slot] << 32) | wq.mask_lo[slot];
    while (mask) {
        const uint32_t bit = __ffsll((long long)mask) - 1u;
        mask &= mask - 1ull;
        const uint32_t r = wide ? (bit >> 3) : (bit >= 3u ? 1u : 0u);
        const uint32_t k = wide ? (bit & 7u) : (bit >= 3u ? bit - 3u : bit);
        const uint32_t x = (xy & 0xFFFFu) + k, y = (xy >> 16) + r;
        const RowC rc = row_setup(s, y);
        float w0, w1, w2;
        edge_eval(s, rc, x, w0, w1, w2);
        emit_fragment(p, s, sh, tri, x, y, w0, w1, w2, keys);
        ++nfrag_count;
    }
}

// 72 registers: three persistent blocks per SM must leave room (8 K registers) for one block of the
// resolve / follow-up kernels of the previous frame, which run beside this kernel on a second stream.
template <bool CHECK_REGULAR, bool BAND, bool TMA>
__global__ void __maxnreg__(72) k_geom3(const __grid_constant__ FrameParams p, const Scene sc,
                                                            const float* __restrict__ chunks,
                                                            unsigned long long* __restrict__ keys, const Queues q,
                                                            const uint32_t batch_chunks, const uint32_t rowmax_shared)
{
    __shared__ G3Queue queues[G3_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    G3Queue& wq = queues[warp];
    const uint32_t n_chunks = (p.n_tri + 31u) >> 5;
    const uint32_t n_warps = gridDim.x * G3_WARPS;
    const uint32_t gw = blockIdx.x * G3_WARPS + warp;
    uint32_t q_head = 0, q_count = 0, nfrag_count = 0;   // warp-uniform ring state
    uint32_t chunks_done = 0;
    const bool do_stamps = p.image && !(p.debug & 2u);
    // Row stamps: per-block copy of rowmax in shared memory (every warp in flight stamps the same
    // few rows, and same-address traffic serialises in L2); flushed once at the end.  Frames taller
    // than the shared-memory budget (rowmax_shared == 0) stamp the global array directly.
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    TmaRing* const rings = reinterpret_cast<TmaRing*>(dyn_smem);
    uint32_t* const s_rowmax_buf = reinterpret_cast<uint32_t*>(dyn_smem + (TMA ? sizeof(TmaRing) * G3_WARPS : 0));
    const uint32_t n_rowmax = ((p.H + 31u) & ~31u) + 64u;
    uint32_t* const rowmax = rowmax_shared ? s_rowmax_buf : q.rowmax;
    if (do_stamps && rowmax_shared)
        for (uint32_t i = threadIdx.x; i < n_rowmax; i += blockDim.x) s_rowmax_buf[i] = 0u;
    if (TMA && lane == 0) {
        for (uint32_t k = 0; k < TMA_STAGES; ++k) mbar_init(&rings[warp].full[k], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // This warp's chunks: batches of `batch_chunks` consecutive chunks, batches strided by n_warps.
    // Two cursors walk that sequence without divisions: `c` (consumed now) and `pf` (being fetched).
    const uint32_t batch_jump = (n_warps - 1u) * batch_chunks + 1u;
    uint32_t c = gw * batch_chunks, c_pos = 0;
    uint32_t pf = c, pf_pos = 0;
    auto advance = [&](uint32_t& idx, uint32_t& pos) {
        if (++pos == batch_chunks) { pos = 0; idx += batch_jump; } else ++idx;
    };
    // Band contexts skip a chunk when no vertex of it can have a destination row inside the band: every computed
    // y' lies within |row 1 of M| * radius + cull_pad of the computed y' of the sphere centre, rows of a triangle
    // are ceil()s of its y' range and a wrapped fragment lands one row further down (hence the 2).  Non-finite
    // bounds compare false and are processed.
    auto culled = [&](uint32_t idx) -> bool {
        if (!BAND || TMA || !p.cull_on) return false;
        const float4 s = __ldg(sc.bounds + idx);
        const float yc = xform_row(p.m + 4, s.x, s.y, s.z);
        const float R = add(mul(s.w, p.cull_scale), p.cull_pad);
        return sub(yc, R) >= (float)p.row1 || add(add(yc, R), 2.0f) <= (float)p.krow0;
    };
    auto advance_live = [&](uint32_t& idx, uint32_t& pos) {   // next chunk of this warp that is not culled
        do advance(idx, pos); while (BAND && !TMA && idx < n_chunks && culled(idx));
    };
    if (BAND && !TMA) {
        while (c < n_chunks && culled(c)) advance(c, c_pos);
        pf = c;
        pf_pos = c_pos;
    }
    TmaRing& ring = rings[TMA ? warp : 0];
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;
    float C = 0.f;
    if (TMA) {
        for (uint32_t k = 0; k < TMA_STAGES; ++k) {
            if (lane == 0 && pf < n_chunks) tma_load_chunk(ring.stage[k], chunks + (size_t)pf * CHUNK_FLOATS, &ring.full[k]);
            advance(pf, pf_pos);
        }
    } else {
        if (c < n_chunks && c * 32u + lane < p.n_tri) { A = __ldg(sc.a + c * 32u + lane); B = __ldg(sc.b + c * 32u + lane); C = __ldg(sc.z3 + c * 32u + lane); }
        advance_live(pf, pf_pos);
    }
    uint32_t stg = 0, phase = 0;
    {
        uint32_t c_after = 0;   // register-prefetch path: the chunk whose data is in flight
        for (; c < n_chunks; TMA ? advance(c, c_pos) : (void)(c = c_after)) {
            const uint32_t t = c * 32u + lane;
            ++chunks_done;
            if (TMA) {
                mbar_wait(&ring.full[stg], phase);
                A = reinterpret_cast<const float4*>(ring.stage[stg])[lane];
                B = reinterpret_cast<const float4*>(ring.stage[stg] + 128)[lane];
                C = ring.stage[stg][256 + lane];
            }
            const float v0 = A.x, v1 = A.y, v2 = A.z, v3 = A.w, v4 = B.x, v5 = B.y, v6 = B.z, v7 = B.w, v8 = C;
            if (!TMA) {   // prefetch the next chunk of this warp into registers
                const uint32_t tn = pf * 32u + lane;
                if (pf < n_chunks && tn < p.n_tri) { A = __ldg(sc.a + tn); B = __ldg(sc.b + tn); C = __ldg(sc.z3 + tn); }
                c_after = pf;
                advance_live(pf, pf_pos);
            }
            // ---- phase A: x'/y' transform, bounds (Triangle::mul, aabb, rasterizer.rs:58-66) --
            const float y1 = xform_row(p.m + 4, v0, v1, v2), x1 = xform_row(p.m + 0, v0, v1, v2);
            const float y2 = xform_row(p.m + 4, v3, v4, v5), x2 = xform_row(p.m + 0, v3, v4, v5);
            const float y3 = xform_row(p.m + 4, v6, v7, v8), x3 = xform_row(p.m + 0, v6, v7, v8);
            const float mn1 = fminf(y1, fminf(y2, y3)), mx1 = fmaxf(y1, fmaxf(y2, y3));
            const float mn0 = fminf(x1, fminf(x2, x3)), mx0 = fmaxf(x1, fmaxf(x2, x3));
            const uint32_t miny = __float2uint_rz(ceilf(fmaxf(mn1, 1.0f)));
            const uint32_t maxy = __float2uint_rz(ceilf(fminf(mx1, p.hm1)));
            const uint32_t minx = __float2uint_rz(ceilf(fmaxf(mn0, 1.0f)));
            const uint32_t maxx = __float2uint_rz(ceilf(fminf(mul(mx0, 2.0f), p.wm1)));
            // whole-frame contexts own every row; band contexts skip triangles whose destination
            // rows (y, or y+1 after a row wrap) miss the band
            const bool has_rows = t < p.n_tri && miny < maxy && (!BAND || (miny < p.row1 && maxy + 1u > p.krow0));
            if (TMA) {
                // every lane has consumed its nine coordinates (the transform above depends on them), so
                // after the warp barrier no shared-memory read of this stage is outstanding: refill it
                __syncwarp();
                if (lane == 0 && pf < n_chunks)
                    tma_load_chunk(ring.stage[stg], chunks + (size_t)pf * CHUNK_FLOATS, &ring.full[stg]);
                advance(pf, pf_pos);
                if (++stg == TMA_STAGES) { stg = 0; phase ^= 1u; }
            }

            // ---- row stamps (rasterizer.rs:89-91) ----------------------------------------------
            // rowmax[y] = max(c + 1) over chunks c that stamp row y.  The lanes of a chunk are
            // neighbours on screen, so their rows fall into a 64-row window: two REDUX.OR give
            // the chunk's row set, then lane i owns rows i and 32+i of the window (conflict-free).
            // Branch-free: with nothing to stamp the need words are 0 and no atomic is issued.
            const uint32_t sy0 = BAND ? max(miny, p.srow0) : miny, sy1 = BAND ? min(maxy, p.srow1) : maxy;
            const bool st = do_stamps && has_rows && sy0 < sy1;
            bool tall = false;   // rows outside the 64-row window: stamped in the rare block below
            if (do_stamps) {
                const uint32_t ymin = __reduce_min_sync(0xFFFFFFFFu, st ? sy0 : 0xFFFFFFFFu);
                const uint32_t base = ymin & ~31u;
                const uint32_t lo = sy0 - base, n = sy1 - sy0;              // meaningful when st
                const bool fits = st && n <= 32u && lo + n <= 64u;          // inside the 2-word window
                tall = st && !fits;
                const unsigned long long m64 = fits ? (unsigned long long)(0xFFFFFFFFu >> ((32u - n) & 31u)) << (lo & 63u) : 0ull;
                const uint32_t need0 = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)m64);
                const uint32_t need1 = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)(m64 >> 32));
                if ((need0 >> lane) & 1u) atomicMax(rowmax + base + lane, c + 1u);
                if ((need1 >> lane) & 1u) atomicMax(rowmax + base + 32u + lane, c + 1u);
            }

            // ---- classification ---------------------------------------------------------------
            const bool live = has_rows && minx < maxx;
            bool regular = true;
            if (CHECK_REGULAR)   // x/y only: z never enters coverage, it is evaluated exactly per fragment
                regular = in_limit(x1) && in_limit(y1) && in_limit(x2) && in_limit(y2) && in_limit(x3) && in_limit(y3);
            const float dx0 = sub(x3, x2), dy0 = sub(y3, y2);
            const float dx1 = sub(x1, x3), dy1 = sub(y1, y3);
            const float dx2 = sub(x2, x1), dy2 = sub(y2, y1);
            const bool cand = live && regular && !backface_proven(p, dx1, dy1, dx2, dy2, mn0, mx0, mn1, mx1);
            const uint32_t rows = maxy - miny, span = maxx - minx;
            // tight width <= 2  <=>  span <= 2 or floor(max_x) <= minx + 1   (see tight_width)
            const bool foot = cand && rows <= 2u && (span <= 2u || __float2uint_rz(floorf(mx0)) <= minx + 1u);   // tier 1
            bool beyond = cand && !foot;   // tier 2 / 3: handled in the rare block

            // ---- tier 1: 2 x 3 footprint in registers, lockstep ---------------------------------
            uint32_t mask = 0, mask_hi = 0;
            bool wide = false;   // footprint bits are row*8 + col (tier 2) instead of row*3 + col
            if (__any_sync(0xFFFFFFFFu, foot)) {
                float cr[2][3], gc[3][3];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float py = (float)(miny + r);
                    cr[r][0] = mul(dx0, sub(py, y2));
                    cr[r][1] = mul(dx1, sub(py, y3));
                    cr[r][2] = mul(dx2, sub(py, y1));
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float px = (float)(minx + k);
                    gc[k][0] = mul(dy0, sub(px, x2));
                    gc[k][1] = mul(dy1, sub(px, x3));
                    gc[k][2] = mul(dy2, sub(px, x1));
                }
                uint32_t cov = 0;
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        // regular triangle: no NaN, so "all >= 0" == "none < 0"
                        const float w0 = sub(cr[r][0], gc[k][0]), w1 = sub(cr[r][1], gc[k][1]), w2 = sub(cr[r][2], gc[k][2]);
                        if (!(w0 < 0.0f || w1 < 0.0f || w2 < 0.0f)) cov |= 1u << (r * 3 + k);
                    }
                // candidates that exist: rows r < rows, columns k < span
                const uint32_t cm = (1u << min(span, 3u)) - 1u;
                const uint32_t valid = cm | (rows > 1u ? cm << 3 : 0u);
                // a row is finished after column 2 if a closing edge (dy >= 0) fails there
                bool open = false;
                if (span > 3u) {
                    const bool nd0 = !(dy0 < 0.0f), nd1 = !(dy1 < 0.0f), nd2 = !(dy2 < 0.0f);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const bool closed = (nd0 && sub(cr[r][0], gc[2][0]) < 0.0f) || (nd1 && sub(cr[r][1], gc[2][1]) < 0.0f) ||
                                            (nd2 && sub(cr[r][2], gc[2][2]) < 0.0f);
                        if ((uint32_t)r < rows && !closed) open = true;
                    }
                }
                if (foot) {
                    if (open) beyond = true;   // sliver: the rare block hands it to k_tail
                    else mask = cov & valid;
                }
            }

            // ---- everything uncommon under one vote: tier 2, tier 3 / irregular queues, tall stamps ----
            if (__any_sync(0xFFFFFFFFu, beyond || tall || (CHECK_REGULAR && live && !regular))) {
                if (tall)
                    for (uint32_t y = sy0; y < sy1; ++y) atomicMax(rowmax + y, c + 1u);
                uint32_t tw;
                {
                    const uint32_t f = __float2uint_rz(floorf(mx0));
                    const uint32_t te = f >= maxx ? maxx : f + 1u;
                    tw = te > minx ? te - minx : 0u;
                }
                bool walk = beyond;
                const bool mid = beyond && !foot && rows <= 8u && tw <= 6u;   // tier 2: up to 8 x 8, one lane each
                if (mid) {
                    Setup s;
                    s.x1 = x1; s.y1 = y1; s.x2 = x2; s.y2 = y2; s.x3 = x3; s.y3 = y3;
                    s.dx0 = dx0; s.dy0 = dy0; s.dx1 = dx1; s.dy1 = dy1; s.dx2 = dx2; s.dy2 = dy2;
                    unsigned long long m = 0ull;
                    bool open = false;
                    for (uint32_t r = 0; r < rows && !open; ++r) {
                        const RowC rc = row_setup(s, miny + r);
                        bool closed = false;
                        for (uint32_t k = 0; k < 8u && k < span; ++k) {
                            float w0, w1, w2;
                            edge_eval(s, rc, minx + k, w0, w1, w2);
                            if (!(w0 < 0.0f || w1 < 0.0f || w2 < 0.0f)) m |= 1ull << (r * 8u + k);
                            else if (row_closed(s, w0, w1, w2)) { closed = true; break; }
                        }
                        open = !closed && span > 8u;   // candidates remain right of the window
                    }
                    if (!open) { mask = (uint32_t)m; mask_hi = (uint32_t)(m >> 32); wide = true; walk = false; }
                }
                // tier 3: row-band work items for k_tail, one warp-aggregated atomic
                const uint32_t walk_items = walk ? (rows + walk_rows_per_item(tw) - 1u) / walk_rows_per_item(tw) : 0u;
                const unsigned need = __ballot_sync(0xFFFFFFFFu, walk_items > 0);
                if (need) {
                    uint32_t wi = walk_items;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t nn = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                        if ((int)lane >= d) wi += nn;
                    }
                    const uint32_t total = __shfl_sync(0xFFFFFFFFu, wi, 31);
                    unsigned long long old = 0;
                    if (lane == 0)
                        old = atomicAdd(&q.aux->walk_counter, ((unsigned long long)__popc(need) << ITEM_BITS) | total);
                    old = __shfl_sync(0xFFFFFFFFu, old, 0);
                    if (walk_items > 0) {
                        const uint32_t slot = (uint32_t)(old >> ITEM_BITS) + __popc(need & ((1u << lane) - 1u));
                        q.walk_tri[slot] = t;
                        q.walk_base[slot] = (old & ITEM_MASK) + (wi - walk_items);
                    }
                }
                if (CHECK_REGULAR) {
                    const unsigned irr = __ballot_sync(0xFFFFFFFFu, live && !regular);
                    if (irr) {
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(&q.aux->irr_count, (uint32_t)__popc(irr));
                        base = __shfl_sync(0xFFFFFFFFu, base, 0);
                        if (live && !regular) q.irr_tri[base + __popc(irr & ((1u << lane) - 1u))] = t;
                    }
                }
            }

            // ---- phase C: park covering triangles; emit 32 at a time --------------------------
            const unsigned cov_lanes = __ballot_sync(0xFFFFFFFFu, (mask | mask_hi) != 0u);
            if (cov_lanes) {
                if (mask | mask_hi) {
                    const uint32_t slot = (q_head + q_count + __popc(cov_lanes & ((1u << lane) - 1u))) & (G3_RING - 1u);
                    wq.raw[0][slot] = v0; wq.raw[1][slot] = v1; wq.raw[2][slot] = v2;
                    wq.raw[3][slot] = v3; wq.raw[4][slot] = v4; wq.raw[5][slot] = v5;
                    wq.raw[6][slot] = v6; wq.raw[7][slot] = v7; wq.raw[8][slot] = v8;
                    wq.tri[slot] = t | (wide ? 0x80000000u : 0u);   // bit 31: tier-2 (8 x 8) footprint
                    wq.xy[slot] = minx | (miny << 16);
                    wq.mask_lo[slot] = mask;
                    wq.mask_hi[slot] = mask_hi;
                }
                q_count += __popc(cov_lanes);
                if (q_count >= 32u) {
                    __syncwarp();
                    g3_emit(p, wq, q_head, 32u, lane, keys, nfrag_count);
                    __syncwarp();
                    q_head = (q_head + 32u) & (G3_RING - 1u);
                    q_count -= 32u;
                }
            }
        }
    }
    if (q_count) {
        __syncwarp();
        g3_emit(p, wq, q_head, q_count, lane, keys, nfrag_count);
    }
    if (do_stamps && rowmax_shared) {   // publish this block's stamps (the probe skips most atomics)
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_rowmax - 64u; i += blockDim.x) {
            const uint32_t m = s_rowmax_buf[i];
            if (m && __ldcg(q.rowmax + i) < m) atomicMax(q.rowmax + i, m);
        }
    }
    if (p.count_frags) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nfrag_count += __shfl_xor_sync(0xFFFFFFFFu, nfrag_count, d);
        if (lane == 0 && nfrag_count) atomicAdd(&q.aux->frag_counter, (unsigned long long)nfrag_count);
        if (lane == 0 && chunks_done) atomicAdd(&q.aux->chunks_done, chunks_done);
    }
}

// One warp per row-band item.  Lanes take 32 consecutive candidates of a row (or
// 2 x 16 / 4 x 8 for narrow triangles); a row ends when a lane of that row sees a
// closing edge fail (everything right of it fails as well) or at the reference's maxx.
SLOTH_DEV void walk_body(const FrameParams& p, const Scene& sc, unsigned long long* __restrict__ keys, const Queues& q,
                         uint32_t n_blocks, const unsigned long long* __restrict__ tile_info)
{
    const unsigned long long packed = q.aux->walk_counter;
    const unsigned long long n_items = packed & ITEM_MASK;
    const uint32_t n_slots = (uint32_t)(packed >> ITEM_BITS);
    if (n_items == 0) return;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long n_warps = (unsigned long long)n_blocks * (blockDim.x >> 5);
    uint32_t nfrag = 0;
    for (unsigned long long item = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
         item < n_items; item += n_warps) {
        // last slot whose base <= item
        uint32_t lo = 0, hi = n_slots;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (q.walk_base[mid] <= item) lo = mid; else hi = mid;
        }
        if (tile_info && tile_info[lo] != ~0ull) continue;   // rasterised by the tile path (tile_kernels.cuh)
        const uint32_t t = q.walk_tri[lo];
        const uint32_t band = (uint32_t)(item - q.walk_base[lo]);
        float v[9];
        load_tri(sc, t, v);
        Setup s;
        setup_tri(p, v, s);
        Shade sh;
        shade_setup(s, sh);
        const uint32_t tw = tight_width(s);
        const uint32_t rpi = walk_rows_per_item(tw);
        const uint32_t y0 = s.miny + band * rpi;
        const uint32_t y1 = min(s.maxy, y0 + rpi);
        // lanes as (32/cw) rows x cw columns: narrow triangles take several rows per step
        const uint32_t cw = tw <= 6u ? 8u : (tw <= 14u ? 16u : 32u);
        const uint32_t lr = lane / cw, lx = lane % cw, rows_per_step = 32u / cw;
        const uint32_t group = (cw == 32u ? 0xFFFFFFFFu : ((1u << cw) - 1u)) << (lr * cw);
        for (uint32_t yb = y0; yb < y1; yb += rows_per_step) {
            const uint32_t y = yb + lr;
            bool done = !(y < y1 && y + 1u >= p.krow0 && y < p.row1);
            const RowC rc = row_setup(s, y);
            for (uint32_t xb = s.minx; xb < s.maxx; xb += cw) {
                const uint32_t x = xb + lx;
                bool closed = false;
                if (!done && x < s.maxx) {
                    float w0, w1, w2;
                    edge_eval(s, rc, x, w0, w1, w2);
                    if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) {
                        emit_fragment(p, s, sh, t, x, y, w0, w1, w2, keys);
                        ++nfrag;
                    } else {
                        closed = row_closed(s, w0, w1, w2);
                    }
                }
                if (__ballot_sync(0xFFFFFFFFu, closed) & group) done = true;   // this lane's row is finished
                if (__all_sync(0xFFFFFFFFu, done)) break;
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
SLOTH_DEV void irregular_body(const FrameParams& p, const Scene& sc, unsigned long long* __restrict__ keys,
                              const Queues& q, uint32_t block, uint32_t n_blocks)
{
    const uint32_t n = q.aux->irr_count;
    uint32_t nfrag = 0;
    for (uint32_t i = block; i < n; i += n_blocks) {
        const uint32_t t = q.irr_tri[i];
        float v[9];
        load_tri(sc, t, v);
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

// One launch for both follow-up passes: blocks [0, walk_blocks) walk the queued row bands,
// the remaining blocks take the irregular triangles.  Both queues are usually empty or tiny.
__global__ void __launch_bounds__(128) k_tail(const __grid_constant__ FrameParams p, const Scene sc,
                                              unsigned long long* __restrict__ keys, const Queues q,
                                              const uint32_t walk_blocks, const unsigned long long* __restrict__ tile_info)
{
    if (blockIdx.x < walk_blocks) walk_body(p, sc, keys, q, walk_blocks, tile_info);
    else irregular_body(p, sc, keys, q, blockIdx.x - walk_blocks, gridDim.x - walk_blocks);
}

// ---------------------------------------------------------------------------------
// Resolve: key plane -> cells.
// ---------------------------------------------------------------------------------
SLOTH_DEV uint32_t key_tri(unsigned long long k) { return ((uint32_t)k >> 5) & MAX_TRIS; }

SLOTH_DEV uint32_t cell_of(const FrameParams& p, const Scene& sc, unsigned long long key)
{
    const uint32_t g = (uint32_t)key & 15u;
    const uint32_t rgb = __ldg(sc.rgb + key_tri(key));
    return (uint32_t)(uint8_t)p.glyph[g] | (rgb << 8);
}

// Column-1 cell of a stamped row on which a (wrapped) fragment also landed: the stamp of
// triangle T is written after all fragments of triangles <= T (the fragment comes from source
// row y-1, the stamp follows row y of the same triangle), so the cell shows '\n' unless the
// fragment's triangle is later than EVERY triangle that stamps the row.  rowmax gives the last
// stamping chunk; only when the fragment's triangle sits in that very chunk are its 32
// triangles looked at (one per lane).  All 32 lanes must call this; returns true where the
// newline wins.
SLOTH_DEV bool stamp_beats_fragment(const FrameParams& p, const Scene& sc, const Queues& q, bool contested,
                                    uint32_t row, uint32_t tri)
{
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t last = 0u;
    if (contested) last = q.rowmax[row] - 1u;             // contested implies the row is stamped
    bool newline = contested && last > (tri >> 5);
    unsigned exact = __ballot_sync(0xFFFFFFFFu, contested && last == (tri >> 5));
    while (exact) {
        const int src = __ffs(exact) - 1;
        exact &= exact - 1u;
        const uint32_t r = __shfl_sync(0xFFFFFFFFu, row, src), tw = __shfl_sync(0xFFFFFFFFu, tri, src);
        const uint32_t t = (tw & ~31u) + lane;            // the 32 triangles of that chunk
        bool hit = false;
        if (t >= tw && t < p.n_tri) {
            float v[9];
            load_tri(sc, t, v);
            Setup s;
            setup_tri(p, v, s);
            hit = r >= s.miny && r < s.maxy;
        }
        const bool any = __any_sync(0xFFFFFFFFu, hit);
        if ((int)lane == src) {
            newline = any;
            atomicAdd(&q.aux->stamp_exact, 1u);
        }
    }
    return newline;
}

// W even: ids are even, one key slot owns cells (2k, 2k+1) of its row.  The grid is (segments of a row, rows): a block
// takes RESOLVE_SEG consecutive slots of ONE row, a warp 32 * RESOLVE_SLOTS of them, a thread RESOLVE_SLOTS slots 32
// apart (every access of a warp is one contiguous 256-byte run).  Row and column come from the block index: no
// division, and the newline stamp of the row (its cell (row, 1), owned by the slot at column 0) concerns exactly one
// warp of the row's first block -- everybody else runs the bare key -> cell path.  The key -> colour gather is a
// dependent load chain, so each thread keeps RESOLVE_SLOTS of them in flight: the kernel usually runs beside the next
// frame's triangle kernel, where only instruction-level parallelism hides the latency, and every instruction it
// issues there is one the triangle kernel does not.
#ifndef RESOLVE_SLOTS_PER_THREAD
#define RESOLVE_SLOTS_PER_THREAD 4
#endif
static constexpr uint32_t RESOLVE_SLOTS = RESOLVE_SLOTS_PER_THREAD;   // per thread
// threads per block of the kernels that run beside the triangle kernel of a neighbouring frame (k_xform, k_resolve_even):
// what is left of an SM next to three k_tri blocks is 16 K registers
#ifndef CO_THREADS
#define CO_THREADS 128
#endif
static constexpr uint32_t RESOLVE_SEG = CO_THREADS * RESOLVE_SLOTS;   // slots of a row per block

__global__ void __launch_bounds__(CO_THREADS) k_resolve_even(const __grid_constant__ FrameParams p, const Scene sc,
                                                             unsigned long long* __restrict__ keys, const Queues q,
                                                             uint32_t* __restrict__ cells, uint32_t n_slots,
                                                             uint32_t n_tail)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t row_local = blockIdx.y;                                            // row of the key plane / cell buffer
    const uint32_t col0 = blockIdx.x * RESOLVE_SEG + warp * (32u * RESOLVE_SLOTS) + lane;   // this thread's columns: col0 + 32 m
    const uint32_t i0 = row_local * p.KW + col0;
    const uint32_t blank = (uint32_t)' ';
    unsigned long long key[RESOLVE_SLOTS];
#pragma unroll
    for (uint32_t m = 0; m < RESOLVE_SLOTS; ++m) {   // all loads first: nothing between them that waits for one of them
        key[m] = KEY_EMPTY;
        if (col0 + 32u * m < p.KW) key[m] = __ldcs(keys + i0 + 32u * m);
    }
#pragma unroll
    for (uint32_t m = 0; m < RESOLVE_SLOTS; ++m)     // the key plane is clean again for the next frame (sectors without a
        if (key[m] != KEY_EMPTY) keys[i0 + 32u * m] = KEY_EMPTY;   // fragment are not written at all)
    uint32_t c0[RESOLVE_SLOTS], c1[RESOLVE_SLOTS];
#pragma unroll
    for (uint32_t m = 0; m < RESOLVE_SLOTS; ++m) {
        c0[m] = blank;
        if (key[m] != KEY_EMPTY) c0[m] = cell_of(p, sc, key[m]);
        c1[m] = c0[m];
    }
    if (blockIdx.x == 0u && warp == 0u) {   // the warp that holds the slot at column 0 (lane 0, m = 0)
        if (p.image && p.KW) {
            const uint32_t row = row_local + p.row0;
            const bool stamped = lane == 0u && q.rowmax[row] != 0u;
            const bool contested = stamped && key[0] != KEY_EMPTY;
            bool newline = stamped;
            if (__any_sync(0xFFFFFFFFu, contested)) {   // a wrapped fragment also landed on the stamped cell: who was later?
                const bool nl = stamp_beats_fragment(p, sc, q, contested, row, key_tri(key[0]));   // all 32 lanes call
                if (contested) newline = nl;
            }
            if (newline) c1[0] = (uint32_t)'\n';
        }
        // image-mode tail, context.rs:38-39: one cell per row behind the frame
        if (lane == 0u && row_local < n_tail) cells[2u * n_slots + row_local] = blank;
    }
#pragma unroll
    for (uint32_t m = 0; m < RESOLVE_SLOTS; ++m)
        if (col0 + 32u * m < p.KW) reinterpret_cast<uint2*>(cells)[i0 + 32u * m] = make_uint2(c0[m], c1[m]);
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
    unsigned long long kw = KEY_EMPTY;
    uint32_t c = blank, row = 0;
    bool stamped = false;
    if (i < n_cells) {
        const uint32_t id = p.row0 * p.W + i;  // global cell index
        const unsigned long long ka = keys[halo + i];
        const unsigned long long kb = (halo + i) > 0 ? keys[halo + i - 1] : KEY_EMPTY;
        if (ka != KEY_EMPTY && kb != KEY_EMPTY) kw = frag_later(p, ka, id, kb, id - 1u) ? ka : kb;
        else if (ka != KEY_EMPTY) kw = ka;
        else if (kb != KEY_EMPTY) kw = kb;
        if (kw != KEY_EMPTY) c = cell_of(p, sc, kw);
        // the stamp of row y is cell y*W + 1 (rasterizer.rs:90): column 1 of row y, or -- in a frame that is one
        // column wide -- column 0 of row y+1
        row = id / p.W;
        bool stamp_cell = id % p.W == 1u;
        if (p.W == 1u) {
            stamp_cell = id >= 1u;
            row = id - (stamp_cell ? 1u : 0u);
        }
        stamped = p.image && stamp_cell && q.rowmax[row] != 0u;
    }
    if (p.image) {
        const bool contested = stamped && kw != KEY_EMPTY;
        const bool nl = stamp_beats_fragment(p, sc, q, contested, row, key_tri(kw));
        if (stamped && (!contested || nl)) c = (uint32_t)'\n';
    }
    if (i < n_cells) {
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

// Scene upload for the TMA feed: the geometry streams regrouped into 1152-byte chunks of 32 triangles.
__global__ void __launch_bounds__(256) k_pack_chunks(const float4* __restrict__ a, const float4* __restrict__ b,
                                                     const float* __restrict__ z3, uint32_t n, uint32_t n_padded,
                                                     float* __restrict__ chunks)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_padded) return;
    float* base = chunks + (size_t)(t >> 5) * CHUNK_FLOATS;
    const uint32_t l = t & 31u;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(base)[l] = t < n ? a[t] : z4;
    reinterpret_cast<float4*>(base + 128)[l] = t < n ? b[t] : z4;
    base[256 + l] = t < n ? z3[t] : 0.f;
}

// Scene upload: raw soup (9 floats + 3 bytes per triangle) -> the four resident streams.
// Bounding sphere of every chunk of 32 triangles (object space) for the band cull of k_geom3, and the largest
// |coordinate| of the scene (bits of a non-negative float in *absmax).  One thread per chunk, once per scene.
__global__ void __launch_bounds__(128) k_chunk_bounds(const float4* __restrict__ a, const float4* __restrict__ b,
                                                      const float* __restrict__ z3, uint32_t n_tri, float4* __restrict__ bounds,
                                                      uint32_t* __restrict__ absmax)
{
    const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_chunks = (n_tri + 31u) >> 5;
    float big = 0.0f;
    if (chunk < n_chunks) {
        const uint32_t t0 = chunk * 32u, t1 = min(n_tri, t0 + 32u);
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        bool finite = true;
        for (uint32_t t = t0; t < t1; ++t) {
            const float4 A = a[t], B = b[t];
            const float v[9] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, z3[t]};
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                lo[k % 3] = fminf(lo[k % 3], v[k]);
                hi[k % 3] = fmaxf(hi[k % 3], v[k]);
                finite = finite && fabsf(v[k]) <= 3.0e38f;
                big = fmaxf(big, fabsf(v[k]));
            }
        }
        const float cx = 0.5f * lo[0] + 0.5f * hi[0], cy = 0.5f * lo[1] + 0.5f * hi[1], cz = 0.5f * lo[2] + 0.5f * hi[2];
        float r2 = 0.0f;
        for (uint32_t t = t0; t < t1; ++t) {
            const float4 A = a[t], B = b[t];
            const float v[9] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, z3[t]};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float dx = v[3 * k] - cx, dy = v[3 * k + 1] - cy, dz = v[3 * k + 2] - cz;
                r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
            }
        }
        // rounded up generously: the radius only has to be an upper bound
        const float r = finite ? sqrtf(r2) * 1.0001f + 1.0e-30f : __int_as_float(0x7FC00000);
        bounds[chunk] = make_float4(cx, cy, cz, r);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) big = fmaxf(big, __shfl_xor_sync(0xFFFFFFFFu, big, d));
    if ((threadIdx.x & 31u) == 0u && big > 0.0f) atomicMax(absmax, __float_as_uint(big));
}

__global__ void __launch_bounds__(256) k_pack_scene(const float* __restrict__ xyz, const uint8_t* __restrict__ rgb,
                                                    uint32_t n, float4* __restrict__ a, float4* __restrict__ b,
                                                    float* __restrict__ z3, uint32_t* __restrict__ col)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* v = xyz + (size_t)t * 9;
    a[t] = make_float4(v[0], v[1], v[2], v[3]);
    b[t] = make_float4(v[4], v[5], v[6], v[7]);
    z3[t] = v[8];
    col[t] = (uint32_t)rgb[(size_t)t * 3] | ((uint32_t)rgb[(size_t)t * 3 + 1] << 8) | ((uint32_t)rgb[(size_t)t * 3 + 2] << 16);
}

}  // namespace sloth
