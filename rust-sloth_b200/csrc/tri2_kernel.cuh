// tri2_kernel.cuh -- k_tri2: k_tri with TWO chunks of 32 triangles per warp turn (sm_100a).
//
// Same job and the same arithmetic as k_tri (tri_kernel.cuh; draw_triangle's prologue and candidate loop,
// rasterizer.rs:56-91), for whole-frame contexts.  k_tri spends about 120 of its ~300 warp instructions per chunk
// on work that does not depend on the triangle count of the turn: the cp.async bookkeeping, the loop, the votes
// and the branch skeleton around phase A.  Here a lane carries triangle t of chunk 2i and triangle t of chunk
// 2i + 1 through one turn: that overhead is paid once per 64 triangles, and the two independent dependency chains
// of phase A fill each other's issue gaps.  Everything from the footprint on (phase B, the uncommon tiers, the
// parking of covered fragments, the emit pass) runs once per half over the same code, the second half's state
// copied into the working registers in between; chunk indices, row stamps and keys are per 32-triangle chunk
// exactly as in k_tri, so k_tail and the resolve kernels do not know the difference.
//
// Pipeline per warp, pair i -> warp i mod n_warps, all copies through cp.async into slots only the copying lane
// reads:    turn k:  wait for everything the previous turn started (coordinates k, records k + 1 were there before)
//                    start the copy of the records of pair k + 3, read the records of pair k + 1 from their slot and
//                    start the six (x', y') gathers of pair k + 1
//                    compute pair k
// A turn lasts several thousand cycles of the SM sub-partition's time (all resident warps take turns), so one turn
// of lead covers the L2 latency of the gathers and two turns the HBM latency of the records.
#pragma once
#include "tri_kernel.cuh"

namespace sloth {

struct TPipe2 {
    uint4 rec[4][2][32];          // 4 slots x 2 halves: 1 KB per slot
    float2 xy[2][2][3][32];       // 2 slots x 2 halves x 3 corners: 1.5 KB per slot
};

struct TWarpSmem2 {
    TPipe2 pipe;
    TRing ring;
};

template <bool CHECK_REGULAR, bool ROWMAX_SHARED>
__global__ void __launch_bounds__(T_WARPS * 32, T_REG_BLOCKS)
k_tri2(const __grid_constant__ FrameParams p, const Scene sc, unsigned long long* __restrict__ keys, const Queues q)
{
    extern __shared__ __align__(16) unsigned char t_smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n_rowmax = (uint32_t)t_rowmax_words(p.H);
    uint32_t* const s_rowmax = reinterpret_cast<uint32_t*>(t_smem);
    TWarpSmem2& ws = reinterpret_cast<TWarpSmem2*>(t_smem + (ROWMAX_SHARED ? n_rowmax * 4u : 0u))[warp];
    TRing& wq = ws.ring;
    const uint32_t n_pairs = (p.n_tri + 63u) >> 6;
    const uint32_t n_warps = gridDim.x * T_WARPS;
    const uint32_t gw = blockIdx.x * T_WARPS + warp;
    uint32_t q_head = 0, q_count = 0, nfrag_count = 0, chunks_done = 0;   // warp-uniform
    const bool do_stamps = p.image && !(p.debug & 2u);
    if (ROWMAX_SHARED && do_stamps)
        for (uint32_t i = threadIdx.x; i < n_rowmax; i += blockDim.x) s_rowmax[i] = 0u;
    __syncthreads();
    uint32_t rowmax_a = smem_u32(s_rowmax);
    asm volatile("" : "+r"(rowmax_a));
    auto stamp = [&](uint32_t row, uint32_t value) {
        if (ROWMAX_SHARED) asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(rowmax_a + row * 4u), "r"(value) : "memory");
        else atomicMax(q.rowmax + row, value);
    };

    const uint32_t n_iter = gw < n_pairs ? (n_pairs - gw + n_warps - 1u) / n_warps : 0u;
    uint32_t rec_a = smem_u32(&ws.pipe.rec[0][0][lane]), xy_a = smem_u32(&ws.pipe.xy[0][0][0][lane]);
    asm volatile("" : "+r"(rec_a), "+r"(xy_a));
    const uint4* const rec_g = sc.rec + (size_t)gw * 64u + lane;
    const uint32_t rec_step = n_warps * 64u;   // records between consecutive pairs of this warp
    auto fetch_rec = [&](uint32_t k, uint32_t slot) {   // both records of pair k -> ring slot
        const uint4* g = rec_g + (size_t)min(k, n_iter - 1u) * rec_step;
        cp_async16(rec_a + (slot << 10), g);
        cp_async16(rec_a + (slot << 10) + 512u, g + 32);
    };
    auto gather_xy = [&](uint32_t rslot, uint32_t xslot) {   // (x', y') of the six corners of the pair in record slot rslot
        const uint4 r0 = lds128(rec_a + (rslot << 10)), r1 = lds128(rec_a + (rslot << 10) + 512u);
        const uint32_t d = xy_a + xslot * 1536u;
        cp_async8(d, sc.vxy + r0.x);
        cp_async8(d + 256u, sc.vxy + r0.y);
        cp_async8(d + 512u, sc.vxy + r0.z);
        cp_async8(d + 768u, sc.vxy + r1.x);
        cp_async8(d + 1024u, sc.vxy + r1.y);
        cp_async8(d + 1280u, sc.vxy + r1.z);
    };
    if (n_iter) {
        fetch_rec(0u, 0u);
        fetch_rec(1u, 1u);
        fetch_rec(2u, 2u);
    }
    cp_async_commit();
    cp_async_wait<0>();
    if (n_iter) gather_xy(0u, 0u);
    cp_async_commit();

    for (uint32_t k = 0; k < n_iter; ++k) {
        const uint32_t c0 = (gw + k * n_warps) * 2u;   // chunk of the first half; the second half is chunk c0 + 1
        const uint32_t rs = k & 3u, xs = k & 1u;
        cp_async_wait<0>();
        fetch_rec(k + 3u, (k + 3u) & 3u);
        gather_xy((k + 1u) & 3u, xs ^ 1u);
        cp_async_commit();
        if (p.count_frags) chunks_done += c0 + 1u < ((p.n_tri + 31u) >> 5) ? 2u : 1u;

        const uint32_t xa = xy_a + xs * 1536u;
        // working set of the half being rasterised (phase B on); phase A fills it for half 0 and parks half 1
        float x1, y1, x2, y2, x3, y3;
        uint32_t miny, maxy;
        bool has_rows, back, regular = true, tall = false;
        float bx1, by1, bx2, by2, bx3, by3;
        uint32_t bminy, bmaxy;
        bool bhas_rows, bback, bregular = true, btall = false;
        {
            const float2 P1 = lds64f(xa), P2 = lds64f(xa + 256u), P3 = lds64f(xa + 512u);
            const float2 Q1 = lds64f(xa + 768u), Q2 = lds64f(xa + 1024u), Q3 = lds64f(xa + 1280u);
            x1 = P1.x; y1 = P1.y; x2 = P2.x; y2 = P2.y; x3 = P3.x; y3 = P3.y;
            bx1 = Q1.x; by1 = Q1.y; bx2 = Q2.x; by2 = Q2.y; bx3 = Q3.x; by3 = Q3.y;
        }
        // ---- phase A, both halves: bounds (Triangle::aabb, rasterizer.rs:58-66), row stamps, back-face proof ----
        const float mn1 = fminf(y1, fminf(y2, y3)), mx1 = fmaxf(y1, fmaxf(y2, y3));
        const float bmn1 = fminf(by1, fminf(by2, by3)), bmx1 = fmaxf(by1, fmaxf(by2, by3));
        miny = __float2uint_rz(ceilf(fmaxf(mn1, 1.0f)));
        maxy = __float2uint_rz(ceilf(fminf(mx1, p.hm1)));
        bminy = __float2uint_rz(ceilf(fmaxf(bmn1, 1.0f)));
        bmaxy = __float2uint_rz(ceilf(fminf(bmx1, p.hm1)));
        has_rows = miny < maxy;      // padding triangles sit on the sentinel vertex (-1e30): maxy = 0, no rows
        bhas_rows = bminy < bmaxy;

        // row stamps (rasterizer.rs:89-91), per chunk as in k_tri: rowmax[y] = max(c + 1) over chunks c stamping y
        if (do_stamps) {
            const uint32_t flags = lds32(rec_a + (rs << 10) + 12u) & lds32(rec_a + (rs << 10) + 512u + 12u);   // warp-uniform
            if (!CHECK_REGULAR && (flags & 1u)) {
                // both chunks hang together through shared vertices: each stamps exactly [min miny, max maxy)
                const uint32_t lo0 = __reduce_min_sync(0xFFFFFFFFu, miny), hi0 = __reduce_max_sync(0xFFFFFFFFu, maxy);
                const uint32_t lo1 = __reduce_min_sync(0xFFFFFFFFu, bminy), hi1 = __reduce_max_sync(0xFFFFFFFFu, bmaxy);
                if (lo0 + lane < hi0) stamp(lo0 + lane, c0 + 1u);
                if (lo1 + lane < hi1) stamp(lo1 + lane, c0 + 2u);
                if (lo0 + 32u < hi0 || lo1 + 32u < hi1) {   // taller than a warp (rare): the remaining rows, strided
                    for (uint32_t y = lo0 + 32u + lane; y < hi0; y += 32u) stamp(y, c0 + 1u);
                    for (uint32_t y = lo1 + 32u + lane; y < hi1; y += 32u) stamp(y, c0 + 2u);
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t sy0 = h ? bminy : miny, sy1 = h ? bmaxy : maxy;
                    const bool st = h ? bhas_rows : has_rows;
                    const uint32_t first = __reduce_min_sync(0xFFFFFFFFu, st ? sy0 : 0xFFFFFFFFu);
                    const uint32_t lo = sy0 - first, n = sy1 - sy0;   // meaningful when st (then n >= 1)
                    const bool fits = st && lo + n <= 32u;
                    if (h) btall = st && !fits; else tall = st && !fits;
                    const uint32_t m = fits ? (0xFFFFFFFFu >> (32u - n)) << lo : 0u;
                    const uint32_t need = __reduce_or_sync(0xFFFFFFFFu, m);
                    if ((need >> lane) & 1u) stamp(first + lane, c0 + 1u + (uint32_t)h);
                }
            }
        }

        // back-face proof (backface_proven, kernels.cuh); bounded scenes use the per-frame distance bound p.bf_k
        {
            const float mn0 = fminf(x1, fminf(x2, x3)), mx0 = fmaxf(x1, fmaxf(x2, x3));
            const float bmn0 = fminf(bx1, fminf(bx2, bx3)), bmx0 = fmaxf(bx1, fmaxf(bx2, bx3));
            const float dx1 = sub(x1, x3), dy1 = sub(y1, y3), dx2 = sub(x2, x1), dy2 = sub(y2, y1);
            const float bdx1 = sub(bx1, bx3), bdy1 = sub(by1, by3), bdx2 = sub(bx2, bx1), bdy2 = sub(by2, by1);
            if (CHECK_REGULAR) {
                regular = in_limit(x1) && in_limit(y1) && in_limit(x2) && in_limit(y2) && in_limit(x3) && in_limit(y3);
                bregular = in_limit(bx1) && in_limit(by1) && in_limit(bx2) && in_limit(by2) && in_limit(bx3) && in_limit(by3);
                back = backface_proven(p, dx1, dy1, dx2, dy2, mn0, mx0, mn1, mx1);
                bback = backface_proven(p, bdx1, bdy1, bdx2, bdy2, bmn0, bmx0, bmn1, bmx1);
            } else {
                const float area = sub(mul(dy2, dx1), mul(dx2, dy1));
                const float barea = sub(mul(bdy2, bdx1), mul(bdx2, bdy1));
                const float T = mul(fmaxf(sub(mx0, mn0), sub(mx1, mn1)), p.bf_k);
                const float bT = mul(fmaxf(sub(bmx0, bmn0), sub(bmx1, bmn1)), p.bf_k);
                back = T > 1e-30f && area < -T;
                bback = bT > 1e-30f && barea < -bT;
            }
        }
        // pairs on the far side of a closed mesh end here (after the stamps); so do empty ones
        const bool maybe = (has_rows && (CHECK_REGULAR ? (!regular || !back) : !back)) || tall;
        const bool bmaybe = (bhas_rows && (CHECK_REGULAR ? (!bregular || !bback) : !bback)) || btall;
        const unsigned any0 = (p.debug & 16u) ? 0u : __ballot_sync(0xFFFFFFFFu, maybe);
        const unsigned any1 = (p.debug & 16u) ? 0u : __ballot_sync(0xFFFFFFFFu, bmaybe);
        if ((any0 | any1) == 0u) continue;

#pragma unroll 1
        for (uint32_t h = any0 ? 0u : 1u; h < 2u; ++h) {
            if (h) {   // second half: its state moves into the working registers
                if (!any1) break;
                x1 = bx1; y1 = by1; x2 = bx2; y2 = by2; x3 = bx3; y3 = by3;
                miny = bminy; maxy = bmaxy;
                has_rows = bhas_rows; back = bback; regular = bregular; tall = btall;
            }
            const uint32_t c = c0 + h;
            const uint32_t t = c * 32u + lane;
            const uint32_t ra = rec_a + (rs << 10) + (h << 9);   // this half's record slot
            uint32_t mask = 0;
            const float mn0 = fminf(x1, fminf(x2, x3)), mx0 = fmaxf(x1, fmaxf(x2, x3));
            const float dx1 = sub(x1, x3), dy1 = sub(y1, y3);
            const float dx2 = sub(x2, x1), dy2 = sub(y2, y1);
            const uint32_t minx = __float2uint_rz(ceilf(fmaxf(mn0, 1.0f)));
            const uint32_t maxx = __float2uint_rz(ceilf(fminf(mul(mx0, 2.0f), p.wm1)));
            const bool live = has_rows && minx < maxx;
            const float dx0 = sub(x3, x2), dy0 = sub(y3, y2);
            const bool cand = live && regular && !back;
            const uint32_t rows = maxy - miny, span = maxx - minx;
            // tight width <= 2  <=>  span <= 2 or floor(max_x) <= minx + 1   (see tight_width)
            const bool foot = cand && rows <= 2u && (span <= 2u || __float2uint_rz(floorf(mx0)) <= minx + 1u);   // tier 1
            bool beyond = cand && !foot;   // tier 2 / 3: handled in the rare block

            // ---- tier 1: 2 x 3 footprint in registers, lockstep (same evaluation as k_tri / k_geom3) ----------
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
                for (int kk = 0; kk < 3; ++kk) {
                    const float px = (float)(minx + kk);
                    gc[kk][0] = mul(dy0, sub(px, x2));
                    gc[kk][1] = mul(dy1, sub(px, x3));
                    gc[kk][2] = mul(dy2, sub(px, x1));
                }
                uint32_t cov = 0;
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) {
                        // regular triangle: no NaN, so "all >= 0" == "none < 0"
                        const float w0 = sub(cr[r][0], gc[kk][0]), w1 = sub(cr[r][1], gc[kk][1]), w2 = sub(cr[r][2], gc[kk][2]);
                        if (!(w0 < 0.0f || w1 < 0.0f || w2 < 0.0f)) cov |= 1u << (r * 3 + kk);
                    }
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
                    for (uint32_t y = miny; y < maxy; ++y) stamp(y, c + 1u);
                uint32_t tw;
                {
                    const uint32_t f = __float2uint_rz(floorf(mx0));
                    const uint32_t te = f >= maxx ? maxx : f + 1u;
                    tw = te > minx ? te - minx : 0u;
                }
                bool walk = beyond;
                const bool mid = beyond && !foot && rows <= 8u && tw <= 6u;   // tier 2: up to 8 x 8, one lane each
                unsigned long long m64 = 0ull;
                if (mid) {
                    Setup s;
                    s.x1 = x1; s.y1 = y1; s.x2 = x2; s.y2 = y2; s.x3 = x3; s.y3 = y3;
                    s.dx0 = dx0; s.dy0 = dy0; s.dx1 = dx1; s.dy1 = dy1; s.dx2 = dx2; s.dy2 = dy2;
                    unsigned long long m = 0ull;
                    bool open = false;
                    for (uint32_t r = 0; r < rows && !open; ++r) {
                        const RowC rc = row_setup(s, miny + r);
                        bool closed = false;
                        for (uint32_t kk = 0; kk < 8u && kk < span; ++kk) {
                            float w0, w1, w2;
                            edge_eval(s, rc, minx + kk, w0, w1, w2);
                            if (!(w0 < 0.0f || w1 < 0.0f || w2 < 0.0f)) m |= 1ull << (r * 8u + kk);
                            else if (row_closed(s, w0, w1, w2)) { closed = true; break; }
                        }
                        open = !closed && span > 8u;   // candidates remain right of the window
                    }
                    if (!open) { m64 = m; walk = false; }
                }
                while (__any_sync(0xFFFFFFFFu, m64 != 0ull)) {   // tier-2 fragments, one per lane and turn
                    const bool has = m64 != 0ull;
                    const uint32_t bit = has ? (uint32_t)__ffsll((long long)m64) - 1u : 0u;
                    m64 &= m64 - 1ull;
                    const unsigned who = __ballot_sync(0xFFFFFFFFu, has);
                    if (has) {
                        const uint32_t slot = (q_head + q_count + __popc(who & ((1u << lane) - 1u))) & (T_RING - 1u);
                        wq.rec[slot] = lds128(ra);   // the record names its triangle (index.cuh)
                        wq.xy[slot] = (minx + (bit & 7u)) | ((miny + (bit >> 3)) << 16);
                    }
                    q_count += __popc(who);
                    if (p.count_frags) nfrag_count += __popc(who);
                    if (q_count >= 32u) {
                        __syncwarp();
                        t_emit(p, sc, wq, q_head, 32u, lane, keys);
                        __syncwarp();
                        q_head = (q_head + 32u) & (T_RING - 1u);
                        q_count -= 32u;
                    }
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

            // ---- phase C: park the covered fragments of the 2 x 3 footprint, one per lane and turn; every 32
            // parked fragments are emitted with all lanes busy ---------------------------------------------
            const uint32_t xy0 = minx | (miny << 16);
            while (__any_sync(0xFFFFFFFFu, mask != 0u)) {
                const bool has = mask != 0u;
                const uint32_t bit = (uint32_t)__ffs((int)mask) - 1u;   // garbage when !has, unused
                mask &= mask - 1u;
                const unsigned who = __ballot_sync(0xFFFFFFFFu, has);
                if (has) {
                    const uint32_t slot = (q_head + q_count + __popc(who & ((1u << lane) - 1u))) & (T_RING - 1u);
                    wq.rec[slot] = lds128(ra);
                    wq.xy[slot] = xy0 + bit + (bit >= 3u ? 65536u - 3u : 0u);   // bit = row * 3 + column
                }
                q_count += __popc(who);
                if (p.count_frags) nfrag_count += __popc(who);
                if (q_count >= 32u) {
                    __syncwarp();
                    t_emit(p, sc, wq, q_head, 32u, lane, keys);
                    __syncwarp();
                    q_head = (q_head + 32u) & (T_RING - 1u);
                    q_count -= 32u;
                }
            }
        }
    }
    cp_async_wait<0>();
    if (q_count) {
        __syncwarp();
        t_emit(p, sc, wq, q_head, q_count, lane, keys);
    }
    if (ROWMAX_SHARED && do_stamps) {   // publish this block's stamps (the probe skips most atomics)
        __syncthreads();
        for (uint32_t i0 = threadIdx.x; i0 < n_rowmax - 64u; i0 += 4u * blockDim.x) {
            uint32_t m[4], g[4];   // four probes in flight per thread: the loop is latency, not bandwidth
#pragma unroll
            for (uint32_t kk = 0; kk < 4u; ++kk) {
                const uint32_t i = i0 + kk * blockDim.x;
                m[kk] = i < n_rowmax - 64u ? s_rowmax[i] : 0u;
                g[kk] = m[kk] ? __ldcg(q.rowmax + i) : 0xFFFFFFFFu;
            }
#pragma unroll
            for (uint32_t kk = 0; kk < 4u; ++kk)
                if (g[kk] < m[kk]) atomicMax(q.rowmax + i0 + kk * blockDim.x, m[kk]);
        }
    }
    if (p.count_frags && lane == 0) {
        if (nfrag_count) atomicAdd(&q.aux->frag_counter, (unsigned long long)nfrag_count);
        if (chunks_done) atomicAdd(&q.aux->chunks_done, chunks_done);
    }
}

}  // namespace sloth
