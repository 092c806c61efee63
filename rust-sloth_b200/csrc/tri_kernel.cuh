// tri_kernel.cuh -- k_tri: the geometry kernel of the indexed path (sm_100a).
//
// Same job as k_geom3 (draw_triangle's prologue and candidate loop, rasterizer.rs:56-91, for every triangle of the
// scene), restructured around the per-vertex stage: k_xform has already applied Triangle::mul to every unique
// vertex, so a triangle costs one 16-byte record plus three 8-byte gathers of (x', y') instead of nine coordinate
// loads and 36 un-fusable multiplies/adds.
//
// Persistent warps, lane = triangle, chunk i -> warp i mod n_warps.  The memory pipeline runs through shared memory
// with cp.async (LDGSTS), every lane copying into slots only it reads, so no registers are held by loads in
// flight, nothing hangs on the six scoreboards, and the depth is what the latency needs (the register-pipelined
// predecessors of this kernel measured ~0.64 IPC with a third of the warp time in long-scoreboard waits):
//   iteration k:  wait until the copies of iteration k-2 have landed
//                 read record k+2 from its ring slot, start the three (x', y') gathers of chunk k+2
//                 compute chunk k from its gathered coordinates
//                 start the copy of record k+4
//   A  bounds (aabb, rasterizer.rs:58-66), image-mode row stamps (chunks that hang together: one REDUX.MIN + one
//      REDUX.MAX + one shared-memory ATOMS.MAX per lane and row; others: a 32-row window and one REDUX.OR), back-face
//      proof with a per-frame distance bound; chunks that are entirely back-facing (the far side of a closed mesh)
//      end here
//   B  2 x 3 lockstep footprint for triangles of at most 2 rows x 2 tight columns (separable edge terms); coverage
//      compares the row term with the column term (cov_test), the edge values themselves are not needed
//   C  covered FRAGMENTS (not triangles) are parked in a per-warp shared-memory ring (the triangle's record, which
//      names the triangle, + x | y << 16); every 32 of them are emitted with all lanes busy -- gather (x', y', z') of
//      the three vertices, normal / 1/area / depth / glyph, one 64-bit atomicMin into the key plane
// Larger triangles: up to 8 x 8 candidates one lane each, beyond that row-band items for k_tail; non-finite or
// huge triangles go to k_tail's brute-force part.  Bit-exactness: every emitted value is produced by the same
// round-to-nearest operations in the same order as raster_core.cuh's soup path; the footprint's coverage decisions use
// the sign identity fl(a - b) < 0 <=> a < b (cov_test).
#pragma once
#include "kernels.cuh"
#include "index.cuh"

namespace sloth {

#ifndef T_WARPS_PER_BLOCK
#define T_WARPS_PER_BLOCK 8
#endif
static constexpr uint32_t T_WARPS = T_WARPS_PER_BLOCK;     // warps per block
#ifndef T_BLOCKS_PER_SM
#define T_BLOCKS_PER_SM 3      // persistent blocks per SM (shared memory: 3 x 60 KB)
#endif
#ifndef T_REG_BLOCKS
#define T_REG_BLOCKS 4         // register budget = 65536 / (256 * T_REG_BLOCKS): 64 registers, so that three resident blocks leave 16 K
                               // registers to the neighbouring frames' kernels (k_xform, k_tail, k_resolve): 142 -> 133 us per frame in batches
#endif
// SLOTH_DEBUG bits 1, 3 and 4 (no stamps / park without emitting / stop after the back-face proof) are timing experiments
// that cost k_tri's inner loop a constant load and a test each per chunk: compiled in only with -DSLOTH_TRI_KNOBS=1
#ifndef SLOTH_TRI_KNOBS
#define SLOTH_TRI_KNOBS 0
#endif
static constexpr uint32_t T_RING = 64;     // per-warp ring of covered fragments (power of two, >= 2 * 32)
static constexpr uint32_t T_STAGES = 4;     // ring depth of the cp.async pipeline (records and coordinates)

// a parked fragment = its triangle's record as it came from memory (i0, i1, i2, triangle << 1 | flag) + the candidate
struct TRing {
    uint4 rec[T_RING];
    uint32_t xy[T_RING];      // x | y << 16
};

// per-warp landing zone of the cp.async pipeline; lane l only ever touches [..][l]
struct TPipe {
    uint4 rec[T_STAGES][32];        // 512 B per stage
    float2 xy[3][T_STAGES][32];     // corner-major: stage s at s * 256 in each corner's KB, so that the slots of k and k + 2
                                    // are one XOR apart like the records'
};

struct TWarpSmem {
    TPipe pipe;
    TRing ring;
};

// smem_dst: 32-bit shared-window address
SLOTH_DEV void cp_async8(uint32_t smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc));
}
SLOTH_DEV void cp_async16(uint32_t smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
SLOTH_DEV uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
SLOTH_DEV uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
SLOTH_DEV void sts128(uint32_t a, const uint4& v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
SLOTH_DEV void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// if (on) { *(uint4*)a16 = v; *(uint32_t*)a4 = w; } in shared memory, as two predicated stores
SLOTH_DEV void park_if(bool on, uint32_t a16, const uint4& v, uint32_t a4, uint32_t w)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.u32 p, %0, 0;\n\t"
                 "@p st.shared.v4.u32 [%1], {%2,%3,%4,%5};\n\t"
                 "@p st.shared.u32 [%6], %7;\n\t}"
                 ::"r"((uint32_t)on), "r"(a16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(a4), "r"(w)
                 : "memory");
}
SLOTH_DEV float2 lds64f(uint32_t a)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
SLOTH_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SLOTH_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// cov |= BIT when none of the three edge values fl(a_i - b_i) is negative, i.e. when no a_i < b_i (finite operands):
// three chained compares and one predicated OR (written out because the compiler's own version needs selects)
template <int BIT>
SLOTH_DEV void cov_test(uint32_t& cov, float a0, float b0, float a1, float b1, float a2, float b2)
{
    asm("{\n\t.reg .pred p;\n\t"
        "setp.geu.f32 p, %1, %2;\n\t"
        "setp.geu.and.f32 p, %3, %4, p;\n\t"
        "setp.geu.and.f32 p, %5, %6, p;\n\t"
        "@p or.b32 %0, %0, %7;\n\t}"
        : "+r"(cov)
        : "f"(a0), "f"(b0), "f"(a1), "f"(b1), "f"(a2), "f"(b2), "n"(BIT));
}

// Emit `count` parked fragments starting at ring position `head`, lane = fragment.  LEAN: whole-frame context.
template <bool LEAN = false>
SLOTH_DEV void t_emit(const FrameParams& p, const Scene& sc, const TRing& wq, uint32_t head, uint32_t count, uint32_t lane,
                      unsigned long long* __restrict__ keys)
{
    if (lane >= count || (SLOTH_TRI_KNOBS && (p.debug & 8u))) return;
    const uint32_t slot = (head + lane) & (T_RING - 1u);
    const uint4 r = wq.rec[slot];
    const uint32_t i0 = r.x, i1 = r.y, i2 = r.z;
    const float2 P1 = __ldg(sc.vxy + i0), P2 = __ldg(sc.vxy + i1), P3 = __ldg(sc.vxy + i2);
    const float z1 = __ldg(sc.vz + i0), z2 = __ldg(sc.vz + i1), z3 = __ldg(sc.vz + i2);
    const uint32_t tri = r.w >> 1, xy = wq.xy[slot];
    Setup s;
    s.x1 = P1.x; s.y1 = P1.y; s.z1 = z1;
    s.x2 = P2.x; s.y2 = P2.y; s.z2 = z2;
    s.x3 = P3.x; s.y3 = P3.y; s.z3 = z3;
    s.dx0 = sub(s.x3, s.x2); s.dy0 = sub(s.y3, s.y2);
    s.dx1 = sub(s.x1, s.x3); s.dy1 = sub(s.y1, s.y3);
    s.dx2 = sub(s.x2, s.x1); s.dy2 = sub(s.y2, s.y1);
    Shade sh;
    shade_setup(s, sh);
    const uint32_t x = xy & 0xFFFFu, y = xy >> 16;
    const RowC rc = row_setup(s, y);
    float w0, w1, w2;
    edge_eval(s, rc, x, w0, w1, w2);
    emit_fragment<LEAN>(p, s, sh, tri, x, y, w0, w1, w2, keys);
}

// ---------------------------------------------------------------------------------
// k_super_cert: once per frame, before k_tri (bounded whole-frame scenes of the indexed path), over the
// super-chunks of SC_TRIS (128) triangles (index.cuh).  It completes the back-face certificate for this frame's matrix;
// a certified super-chunk never reaches k_tri -- k_super_stamp, which runs between k_tri and the resolve of the frame,
// stamps the rows its triangles would have stamped; the others are appended to the list k_tri works through.
//
// Certificate.  The doubled screen area k_tri computes for a triangle is, in exact arithmetic on the ideal
// coordinates, A = c . n_t with c = (row 0 of M) x (row 1 of M) and n_t = (V1-V3) x (V2-V1).  All unit normals of the
// super-chunk lie within theta of the axis a, so with beta = angle(c, a):  A <= |n_t| |c| cos(beta - theta) <=
// n_min (c.a cos(theta) + |c| sin(theta))  as soon as that bound is negative.  k_tri proves a triangle back-facing
// when its COMPUTED area is below -T, T = (larger bbox side) * bf_k (backface_proven).  With e = bound on the
// rounding error of a computed coordinate, L = emax * cone_s + 2e >= every bbox side, eta = 2e + 2^-23 L >= the error
// of a computed edge difference:  |computed area - A| <= 4 L eta + 2 eta^2 + 2^-21 (L + eta)^2 =: E  and  T <= L bf_k.
// The super-chunk is skipped when  n_min (c.a cos(theta) + |c| sin(theta)) <= -2 (L bf_k + E)  (twice what is needed,
// plus 1e-5 |c| for the float evaluation of this very inequality): then every one of its triangles would have been
// proven back-facing by k_tri, none has a candidate, and all it contributes are row stamps.
//
// Stamps.  The triangles hang together through shared vertices, so together they stamp exactly the rows
// [ceil(max(min y', 1)), ceil(min(max y', H-1))) over the super-chunk's vertices (ceil commutes with min / max); y'
// comes from k_xform's output.  The value stamped is the index of the super-chunk's LAST chunk (+1): rowmax is only
// ever compared with the chunk of a fragment's triangle (stamp_beats_fragment), fragments never come from a skipped
// super-chunk, and every index inside the super-chunk compares alike with every index outside it.
// ---------------------------------------------------------------------------------
// k_super_cert: lane = super-chunk.  Certified ones go to the skip list (k_super_stamp), the others to the list k_tri
// works through; one warp-aggregated reservation per list and warp (the order of the lists carries no meaning).
__global__ void __launch_bounds__(256) k_super_cert(const __grid_constant__ FrameParams p, const ix::SuperChunk* __restrict__ super,
                                                    uint32_t n_sc, const Queues q)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t sc = blockIdx.x * 256u + threadIdx.x;
    bool skip = false;
    if (sc < n_sc) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(super + sc));        // ax ay az sin_theta
        const float4 b = __ldg(reinterpret_cast<const float4*>(super + sc) + 1);    // inv_nmin emax n_ids pad
        const float e = p.cone_ev;
        const float L = b.y * p.cone_s + 2.0f * e;
        const float eta = 2.0f * e + 1.1920929e-07f * L;
        const float E = 4.0f * L * eta + 2.0f * eta * eta + 4.7683716e-07f * (L + eta) * (L + eta);
        const float Tm = L * p.bf_k;
        const float K = 2.002f * (Tm + E);
        const float lhs = p.cone_c[0] * a.x + p.cone_c[1] * a.y + p.cone_c[2] * a.z + 1.00001f * p.cone_cnorm * a.w +
                          1.00001f * K * b.x + 1.0e-5f * p.cone_cnorm;
        skip = a.w < 1.5f && Tm > 1.0e-20f && lhs <= 0.0f;   // NaN / inf anywhere: not skipped
    }
    const unsigned m_skip = __ballot_sync(0xFFFFFFFFu, skip);
    const unsigned m_live = __ballot_sync(0xFFFFFFFFu, sc < n_sc && !skip);
    uint32_t base_live = 0, base_skip = 0;
    if (lane == 0) {
        if (m_live) base_live = atomicAdd(&q.cone_cnt->n_live, (uint32_t)__popc(m_live));
        if (m_skip) base_skip = atomicAdd(&q.cone_cnt->n_skip, (uint32_t)__popc(m_skip));
    }
    base_live = __shfl_sync(0xFFFFFFFFu, base_live, 0);
    base_skip = __shfl_sync(0xFFFFFFFFu, base_skip, 0);
    const unsigned below = (1u << lane) - 1u;
    if (skip) {
        q.skip_sc[base_skip + __popc(m_skip & below)] = sc;
    } else if (sc < n_sc) {
        // k_tri's work list is flat: one entry per chunk.  It starts with the chunks behind the last full super-chunk
        // (written once, when the scene is set), the chunks of the live super-chunks follow.
        const uint32_t n_tail = ((p.n_tri + 31u) >> 5) - p.cone_n_super * ix::SC_CHUNKS;
        uint32_t* out = q.live_sc + n_tail + (base_live + __popc(m_live & below)) * ix::SC_CHUNKS;
#pragma unroll
        for (uint32_t j = 0; j < ix::SC_CHUNKS; ++j) out[j] = sc * ix::SC_CHUNKS + j;
    }
}

// k_super_stamp (image mode, after k_tri, before the resolve of the same frame): the rows [lo, hi) a certified
// super-chunk stamps are the min / max of y' over its vertex list.  One warp takes E entries of the skip list per turn
// (E x SC_IDS = 384 ids, 12 per lane) and issues each level of the dependent chain -- list entry, vertex ids, y' -- for
// all of them before it waits: the kernel is nothing but memory latency.  The list is walked from the end: it is
// roughly ascending, so the highest indices reach a row first and most later entries find it stamped already (one load
// instead of an atomic on a contended word).
__global__ void __launch_bounds__(128) k_super_stamp(const __grid_constant__ FrameParams p, const uint32_t* __restrict__ ids,
                                                     const float2* __restrict__ vxy, const Queues q)
{
    constexpr uint32_t E = 8u / ix::SC_CHUNKS, J = ix::SC_IDS / 32u;
    const uint32_t lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    const uint32_t n_skip = q.cone_cnt->n_skip, n_warps = gridDim.x * wpb;
    for (uint32_t i = (blockIdx.x * wpb + (threadIdx.x >> 5)) * E; i < n_skip; i += n_warps * E) {
        uint32_t s_c[E], id[E][J];
        float y[E][J];
#pragma unroll
        for (uint32_t e = 0; e < E; ++e) s_c[e] = q.skip_sc[n_skip - 1u - min(i + e, n_skip - 1u)];   // past the end: the last entry again
#pragma unroll
        for (uint32_t e = 0; e < E; ++e)
#pragma unroll
            for (uint32_t j = 0; j < J; ++j) id[e][j] = __ldg(ids + (size_t)s_c[e] * ix::SC_IDS + j * 32u + lane);
#pragma unroll
        for (uint32_t e = 0; e < E; ++e)
#pragma unroll
            for (uint32_t j = 0; j < J; ++j) y[e][j] = __ldg(&vxy[id[e][j]].y);
#pragma unroll
        for (uint32_t e = 0; e < E; ++e) {
            if (e && i + e >= n_skip) break;   // warp-uniform
            float mn = y[e][0], mx = y[e][0];
#pragma unroll
            for (uint32_t j = 1; j < J; ++j) { mn = fminf(mn, y[e][j]); mx = fmaxf(mx, y[e][j]); }
            // ceil commutes with min / max: the same rows as the union of the triangles' own [miny, maxy)
            const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, __float2uint_rz(ceilf(fmaxf(mn, 1.0f))));
            const uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, __float2uint_rz(ceilf(fminf(mx, p.hm1))));
            const uint32_t value = s_c[e] * ix::SC_CHUNKS + ix::SC_CHUNKS;   // 1 + index of its last chunk
            for (uint32_t r = lo + lane; r < hi; r += 32u)
                if (__ldcg(q.rowmax + r) < value) atomicMax(q.rowmax + r, value);
        }
    }
}

// Dynamic shared memory of one block: [rowmax copy (ROWMAX_SHARED)] [TWarpSmem x T_WARPS]
SLOTH_DEV size_t t_rowmax_words(uint32_t H) { return ((H + 31u) & ~31u) + 64u; }

// ROWMAX_SHARED: the per-block copy of rowmax lives in dynamic shared memory (frames up to ~8 K rows), so the row
// stamps are shared-memory atomics (ATOMS) instead of generic ones.
// CONE: the chunks come from the flat list k_super_cert left (whole-frame, bounded scenes): the static entries for
// the chunks behind the last full super-chunk, then the chunks of every super-chunk it could not certify.
#ifdef T_MAXNREG   // register cap given directly (A/B builds): what three resident blocks leave goes to the neighbouring frames' kernels
#define T_BOUNDS __maxnreg__(T_MAXNREG)   // (cannot be combined with __launch_bounds__)
#else
#define T_BOUNDS __launch_bounds__(T_WARPS * 32, T_REG_BLOCKS)
#endif
template <bool CHECK_REGULAR, bool BAND, bool ROWMAX_SHARED, bool CONE = false>
__global__ void T_BOUNDS
k_tri(const __grid_constant__ FrameParams p, const Scene sc, unsigned long long* __restrict__ keys, const Queues q)
{
    extern __shared__ __align__(16) unsigned char t_smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n_rowmax = (uint32_t)t_rowmax_words(p.H);
    uint32_t* const s_rowmax = reinterpret_cast<uint32_t*>(t_smem);
    TWarpSmem& ws = reinterpret_cast<TWarpSmem*>(t_smem + (ROWMAX_SHARED ? n_rowmax * 4u : 0u))[warp];
    TRing& wq = ws.ring;
    const uint32_t n_chunks = (p.n_tri + 31u) >> 5;
    const uint32_t n_warps = gridDim.x * T_WARPS;
    const uint32_t gw = blockIdx.x * T_WARPS + warp;
    // fragment ring, absolute counters (warp-uniform): q_tail fragments parked so far, q_lim - 32 of them emitted; entry i
    // lives in slot i mod T_RING.  Fewer than 32 wait at any time, at most 32 arrive per turn.
    uint32_t q_tail = 0, q_lim = 32u, chunks_done = 0;
    const bool do_stamps = p.image && !(SLOTH_TRI_KNOBS && (p.debug & 2u));
    if (ROWMAX_SHARED && do_stamps)
        for (uint32_t i = threadIdx.x; i < n_rowmax; i += blockDim.x) s_rowmax[i] = 0u;
    __syncthreads();
    uint32_t stamp_bit = do_stamps ? 1u : 0u;   // ANDed with the record's connected flag: one test per chunk for both
    asm volatile("" : "+r"(stamp_bit));
    uint32_t rowmax_a = smem_u32(s_rowmax);
    asm volatile("" : "+r"(rowmax_a));   // keep the shared-window address in a register (the compiler would re-derive it
                                         // from the CTA id with an S2UR at every use)
    auto stamp = [&](uint32_t row, uint32_t value) {
        if (ROWMAX_SHARED) asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(rowmax_a + row * 4u), "r"(value) : "memory");
        else atomicMax(q.rowmax + row, value);
    };

    // Band contexts skip a chunk whose bounding sphere cannot reach the band's rows (same test as k_geom3).
    auto culled = [&](uint32_t idx) -> bool {
        if (!BAND || !p.cull_on) return false;
        const float4 s = __ldg(sc.bounds + idx);
        const float yc = xform_row(p.m + 4, s.x, s.y, s.z);
        const float R = add(mul(s.w, p.cull_scale), p.cull_pad);
        return sub(yc, R) >= (float)p.row1 || add(add(yc, R), 2.0f) <= (float)p.krow0;
    };

    // This warp's chunks: gw, gw + n_warps, ...  (n_iter of them).  Ring slot of iteration k: k & 3, so the slots
    // of k + 2 and k + 4 are one XOR away.  Band contexts keep `live`, a shift register over the next five
    // iterations (bit j <-> iteration k + j: not culled): only live chunks are copied, gathered and computed.
    // Past the end the record copies are clamped to the warp's last chunk (valid memory, results unused).
    // CONE: work item i = entry i of the flat chunk list k_super_cert left (the chunks behind the last full super-chunk,
    // then the chunks of the super-chunks it could not certify)
    const uint32_t n_items = CONE ? (n_chunks - p.cone_n_super * ix::SC_CHUNKS) + q.cone_cnt->n_live * ix::SC_CHUNKS : n_chunks;
    const uint32_t n_iter = gw < n_items ? (n_items - gw + n_warps - 1u) / n_warps : 0u;
    uint32_t rec_a = smem_u32(&ws.pipe.rec[0][lane]), xy_a = smem_u32(&ws.pipe.xy[0][0][lane]);
    uint32_t ring_a = smem_u32(&wq.rec[0]);   // wq.xy follows at + 16 * T_RING
    asm volatile("" : "+r"(rec_a), "+r"(xy_a), "+r"(ring_a));
    const uint4* const rec_g = sc.rec + (size_t)gw * 32u + lane;
    const uint32_t rec_step = n_warps * 32u;   // records between consecutive chunks of this warp
    auto is_live = [&](uint32_t k) -> uint32_t {
        if (!BAND) return 1u;
        return (k < n_iter && !culled(gw + k * n_warps)) ? 1u : 0u;
    };
    // CONE: the chunk index is only needed to fetch the record (the record names its own triangle): one broadcast load
    // from the list, issued at the top of the iteration that ends with the fetch.  The list is allocated with room for
    // the look-ahead past its end and only ever holds valid chunk indices (zeroed at allocation), so the walk needs no
    // clamp: what it fetches past the end is never used.
    const uint32_t* live_p = q.live_sc + gw;   // list entry of the next chunk index to load (a running pointer: no constant
                                               // load and no index arithmetic per chunk)
    const float2* const vxy_p = sc.vxy;
    auto next_chunk = [&]() -> uint32_t {
        const uint32_t c = __ldg(live_p);
        live_p += n_warps;
        return c;
    };
    auto fetch_rec = [&](uint32_t k, uint32_t slot) {   // record of iteration k -> ring slot (CONE: k ascends by one per call)
        if (CONE) cp_async16(rec_a + (slot << 9), sc.rec + (size_t)next_chunk() * 32u + lane);
        else cp_async16(rec_a + (slot << 9), rec_g + (size_t)min(k, n_iter - 1u) * rec_step);
    };
    auto gather_xy = [&](uint32_t slot) {   // (x', y') of the three corners of the record in ring slot `slot`
        const uint4 r = lds128(rec_a + (slot << 9));
        cp_async8(xy_a + (slot << 8), vxy_p + r.x);
        cp_async8(xy_a + (slot << 8) + 1024u, vxy_p + r.y);
        cp_async8(xy_a + (slot << 8) + 2048u, vxy_p + r.z);
    };
    uint32_t live = 0;
    if (n_iter) {
        for (uint32_t j = 0; j < 4u; ++j) {
            const uint32_t l = is_live(j);
            live |= l << j;
            if (l) fetch_rec(j, j);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    if (n_iter && (live & 1u)) gather_xy(0u);
    cp_async_commit();
    if (n_iter && (live & 2u)) gather_xy(1u);
    cp_async_commit();

    for (uint32_t k = 0; k < n_iter; ++k) {
        const uint32_t ps = k & 3u;   // ring slot of this iteration's record and coordinates
        uint32_t c_next = 0;
        if (CONE) c_next = next_chunk();   // chunk of iteration k + 4: in flight during this one, used by the fetch at its end
        cp_async_wait<1>();   // everything committed two iterations ago has landed: coordinates k, record k + 2
        if (!BAND || (live & 4u)) gather_xy(ps ^ 2u);
        uint32_t mask = 0, minx = 0, miny = 0;
        if (!BAND || (live & 1u)) {
        if (BAND) ++chunks_done;
        const uint32_t rec_w = lds32(rec_a + (ps << 9) + 12u);   // triangle << 1 | chunk-connected flag (index.cuh)
        const uint32_t t = rec_w >> 1, c = t >> 5;
        const float2 P1 = lds64f(xy_a + (ps << 8)), P2 = lds64f(xy_a + (ps << 8) + 1024u), P3 = lds64f(xy_a + (ps << 8) + 2048u);
        const float x1 = P1.x, y1 = P1.y, x2 = P2.x, y2 = P2.y, x3 = P3.x, y3 = P3.y;

        // ---- phase A: bounds (Triangle::aabb, rasterizer.rs:58-66) ----------------------------
        const float mn1 = fminf(y1, fminf(y2, y3)), mx1 = fmaxf(y1, fmaxf(y2, y3));
        miny = __float2uint_rz(ceilf(fmaxf(mn1, 1.0f)));
        const uint32_t maxy = __float2uint_rz(ceilf(fminf(mx1, p.hm1)));
        // padding triangles sit on the sentinel vertex (-1e30): maxy = 0, no rows
        const bool has_rows = miny < maxy && (!BAND || (miny < p.row1 && maxy + 1u > p.krow0));

        // ---- row stamps (rasterizer.rs:89-91): rowmax[y] = max(c + 1) over chunks c that stamp row y.  The
        // lanes of a chunk are neighbours on screen: one REDUX.MIN gives the first row, each lane's rows
        // become bits of a 32-row window, one REDUX.OR, then lane i owns row first + i.
        const uint32_t sy0 = BAND ? max(miny, p.srow0) : miny, sy1 = BAND ? min(maxy, p.srow1) : maxy;
        bool tall = false;   // rows outside the window: stamped in the rare block
        // A chunk whose 32 triangles hang together through shared vertices (flag in the record, set at scene-set)
        // stamps exactly the rows [min miny, max maxy): the y-ranges of triangles that share a vertex overlap or
        // touch, so the union of the chunk's ranges has no gap, and ceil() commutes with min / max.
        const bool connected = !BAND && !CHECK_REGULAR && (rec_w & stamp_bit) != 0u;   // warp-uniform; implies do_stamps
        if (connected) {
            const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, miny);
            const uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, maxy);
            if (lo + lane < hi) stamp(lo + lane, c + 1u);
            if (lo + 32u < hi)   // taller than a warp (rare): the remaining rows, strided
                for (uint32_t y = lo + 32u + lane; y < hi; y += 32u) stamp(y, c + 1u);
        } else if (do_stamps) {
            const bool st = has_rows && (!BAND || sy0 < sy1);
            const uint32_t first = __reduce_min_sync(0xFFFFFFFFu, st ? sy0 : 0xFFFFFFFFu);
            const uint32_t lo = sy0 - first, n = sy1 - sy0;   // meaningful when st (then n >= 1)
            const bool fits = st && lo + n <= 32u;
            tall = st && !fits;
            const uint32_t m = fits ? (0xFFFFFFFFu >> (32u - n)) << lo : 0u;
            const uint32_t need = __reduce_or_sync(0xFFFFFFFFu, m);
            if ((need >> lane) & 1u) stamp(first + lane, c + 1u);
        }

        // ---- back-face proof (backface_proven, kernels.cuh).  Bounded scenes use one distance bound for the
        // whole frame (p.bf_k = 2^-18 * D_frame, D_frame >= every triangle's D, rounded up on the host): a
        // larger D only proves fewer triangles.
        const float mn0 = fminf(x1, fminf(x2, x3)), mx0 = fmaxf(x1, fmaxf(x2, x3));
        const float dx1 = sub(x1, x3), dy1 = sub(y1, y3);
        const float dx2 = sub(x2, x1), dy2 = sub(y2, y1);
        bool regular = true;
        if (CHECK_REGULAR)
            regular = in_limit(x1) && in_limit(y1) && in_limit(x2) && in_limit(y2) && in_limit(x3) && in_limit(y3);
        bool back;
        if (CHECK_REGULAR) {
            back = backface_proven(p, dx1, dy1, dx2, dy2, mn0, mx0, mn1, mx1);
        } else {
            const float area = sub(mul(dy2, dx1), mul(dx2, dy1));
            const float T = mul(fmaxf(sub(mx0, mn0), sub(mx1, mn1)), p.bf_k);
            back = T > 1e-30f && area < -T;
        }
        // chunks on the far side of a closed mesh end here (after the stamps); so do empty ones
        const bool maybe = has_rows && (CHECK_REGULAR ? (!regular || !back) : !back) && !(SLOTH_TRI_KNOBS && (p.debug & 16u));
        if (__any_sync(0xFFFFFFFFu, maybe || tall)) {
            minx = __float2uint_rz(ceilf(fmaxf(mn0, 1.0f)));
            const uint32_t maxx = __float2uint_rz(ceilf(fminf(mul(mx0, 2.0f), p.wm1)));
            const bool live = has_rows && minx < maxx;
            const float dx0 = sub(x3, x2), dy0 = sub(y3, y2);
            const bool cand = live && regular && !back;
            const uint32_t rows = maxy - miny, span = maxx - minx;
            // tight width <= 2  <=>  span <= 2 or floor(max_x) <= minx + 1   (see tight_width)
            // (span <= 3: a triangle at the left edge of the screen; tier 2 takes those, so that the footprint's three columns
            // are always inside the scan domain and column 2 always decides whether the row is finished)
            const bool foot = cand && rows <= 2u && span > 3u && __float2uint_rz(floorf(mx0)) <= minx + 1u;   // tier 1
            bool beyond = cand && !foot;   // tier 2 / 3: handled in the rare block

            // ---- tier 1: 2 x 3 footprint in registers, lockstep (same evaluation as k_geom3) ------------
            if (__any_sync(0xFFFFFFFFu, foot)) {
                float cr[2][3], gc[3][3];
                const float fx0 = (float)minx, fy0 = (float)miny;   // < 2^16 where it matters: fx0 + k == (float)(minx + k) exactly
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float py = r ? add(fy0, 1.0f) : fy0;
                    cr[r][0] = mul(dx0, sub(py, y2));
                    cr[r][1] = mul(dx1, sub(py, y3));
                    cr[r][2] = mul(dx2, sub(py, y1));
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float px = k ? add(fx0, (float)k) : fx0;
                    gc[k][0] = mul(dy0, sub(px, x2));
                    gc[k][1] = mul(dy1, sub(px, x3));
                    gc[k][2] = mul(dy2, sub(px, x1));
                }
                uint32_t cov = 0;
                // regular triangle: no NaN, so "all >= 0" == "none < 0"; and the edge value w = fl(cr - gc) is negative
                // exactly when cr < gc: both are finite, a non-zero difference of two floats is at least 2^-149 in
                // magnitude and denormals are kept, so rounding never changes its sign or makes it zero
                cov_test<1>(cov, cr[0][0], gc[0][0], cr[0][1], gc[0][1], cr[0][2], gc[0][2]);
                cov_test<2>(cov, cr[0][0], gc[1][0], cr[0][1], gc[1][1], cr[0][2], gc[1][2]);
                cov_test<4>(cov, cr[0][0], gc[2][0], cr[0][1], gc[2][1], cr[0][2], gc[2][2]);
                cov_test<8>(cov, cr[1][0], gc[0][0], cr[1][1], gc[0][1], cr[1][2], gc[0][2]);
                cov_test<16>(cov, cr[1][0], gc[1][0], cr[1][1], gc[1][1], cr[1][2], gc[1][2]);
                cov_test<32>(cov, cr[1][0], gc[2][0], cr[1][1], gc[2][1], cr[1][2], gc[2][2]);
                const uint32_t valid = rows > 1u ? 63u : 7u;
                // a row is finished after column 2 if a closing edge (dy >= 0) fails there
                bool open = false;
                {
                    const bool nd0 = !(dy0 < 0.0f), nd1 = !(dy1 < 0.0f), nd2 = !(dy2 < 0.0f);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        // (bitwise on purpose: predicate logic instead of a chain of branches)
                        const bool closed = (nd0 & (cr[r][0] < gc[2][0])) | (nd1 & (cr[r][1] < gc[2][1])) | (nd2 & (cr[r][2] < gc[2][2]));
                        open |= ((uint32_t)r < rows) & !closed;
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
                    for (uint32_t y = sy0; y < sy1; ++y) stamp(y, c + 1u);
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
                        for (uint32_t k = 0; k < 8u && k < span; ++k) {
                            float w0, w1, w2;
                            edge_eval(s, rc, minx + k, w0, w1, w2);
                            if (!(w0 < 0.0f || w1 < 0.0f || w2 < 0.0f)) m |= 1ull << (r * 8u + k);
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
                        const uint32_t slot = (q_tail + __popc(who & ((1u << lane) - 1u))) & (T_RING - 1u);
                        wq.rec[slot] = lds128(rec_a + (ps << 9));
                        wq.xy[slot] = (minx + (bit & 7u)) | ((miny + (bit >> 3)) << 16);
                    }
                    q_tail += __popc(who);
                    if (q_tail >= q_lim) {
                        __syncwarp();
                        t_emit<!BAND && !SLOTH_TRI_KNOBS>(p, sc, wq, q_lim - 32u, 32u, lane, keys);
                        __syncwarp();
                        q_lim += 32u;
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
        }

        }   // live chunk

        // ---- phase C: park the covered fragments of the 2 x 3 footprint, one per lane and turn (most chunks need one
        // turn, few more than two): the record as it stands in the pipeline ring (one 16-byte store) and the candidate.
        // Every 32 parked fragments are emitted with all lanes busy.
        unsigned who = __ballot_sync(0xFFFFFFFFu, mask != 0u);
        if (who) {
            const uint32_t xy0 = minx | (miny << 16);
            const uint4 r = lds128(rec_a + (ps << 9));
            unsigned below = (1u << lane) - 1u;
            asm volatile("" : "+r"(below));   // otherwise recomputed in every turn
            do {
                const bool has = mask != 0u;
                const uint32_t bit = (uint32_t)__ffs((int)mask) - 1u;   // row * 3 + column; garbage when !has, unused
                mask &= mask - 1u;
                uint32_t o16;   // byte offset of the ring entry (opaque to the compiler, which otherwise goes back to a slot
                                // index and shifts it twice)
                asm("and.b32 %0, %1, %2;" : "=r"(o16) : "r"((q_tail + __popc(who & below)) << 4), "n"(T_RING * 16u - 16u));
                // both stores under one predicate instead of a branch around them (the address arithmetic above is harmless
                // for lanes that have nothing to park)
                park_if(has, ring_a + o16, r, ring_a + T_RING * 16u + (o16 >> 2), xy0 + bit + (bit >= 3u ? 65536u - 3u : 0u));
                q_tail += __popc(who);
                if (q_tail >= q_lim) {
                    __syncwarp();
                    t_emit<!BAND && !SLOTH_TRI_KNOBS>(p, sc, wq, q_lim - 32u, 32u, lane, keys);
                    __syncwarp();
                    q_lim += 32u;
                }
                who = __ballot_sync(0xFFFFFFFFu, mask != 0u);
            } while (who);
        }

        // ---- record of iteration k + 4 (same ring slot as record k, which the parking above was the last to
        // read), then one commit for everything this iteration started
        if (BAND) {
            const uint32_t l = is_live(k + 4u);
            live = (live >> 1) | (l << 3);
            if (l) fetch_rec(k + 4u, ps);
        } else if (CONE) {
            cp_async16(rec_a + (ps << 9), sc.rec + (size_t)c_next * 32u + lane);
        } else {
            fetch_rec(k + 4u, ps);
        }
        cp_async_commit();
    }
    cp_async_wait<0>();
    if (!BAND) chunks_done = n_iter;
    if (q_tail != q_lim - 32u) {
        __syncwarp();
        t_emit<!BAND && !SLOTH_TRI_KNOBS>(p, sc, wq, q_lim - 32u, q_tail - (q_lim - 32u), lane, keys);
    }
    if (ROWMAX_SHARED && do_stamps) {   // publish this block's stamps (the probe skips most atomics)
        __syncthreads();
        for (uint32_t i0 = threadIdx.x; i0 < n_rowmax - 64u; i0 += 4u * blockDim.x) {
            uint32_t m[4], g[4];   // four probes in flight per thread: the loop is latency, not bandwidth
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) {
                const uint32_t i = i0 + k * blockDim.x;
                m[k] = i < n_rowmax - 64u ? s_rowmax[i] : 0u;
                g[k] = m[k] ? __ldcg(q.rowmax + i) : 0xFFFFFFFFu;
            }
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k)
                if (g[k] < m[k]) atomicMax(q.rowmax + i0 + k * blockDim.x, m[k]);
        }
    }
    if (p.count_frags && lane == 0) {
        if (q_tail) atomicAdd(&q.aux->frag_counter, (unsigned long long)q_tail);
        if (chunks_done) atomicAdd(&q.aux->chunks_done, chunks_done);
    }
}

}  // namespace sloth
