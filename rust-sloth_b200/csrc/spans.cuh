// spans.cuh -- run-length "span" wire format of the cell buffer (sm_100a), SURVEY.md 8(e)/(f): what crosses PCIe when
// the frame's destination is host memory (Context.frame_buffer, context.rs:16).
//
// A frame is mostly runs of equal cells: the blank background, and inside a model every fragment writes the same
// (glyph, colour) pair to two neighbouring cells (rasterizer.rs:83-85) with the glyph changing only where the shade
// crosses one of nine thresholds.  Instead of 4 bytes per cell the device sends one (start index, cell) pair per
// run; the host side of the C ABI rebuilds the 4-byte cells (host/wire.cpp) so callers see the same buffer.
//   k_span_count   run starts per block of 1024 cells (cell i starts a run iff i == 0 or cell[i] != cell[i-1])
//   k_flush_scan   (flush.cuh) exclusive scan of the block sums, total
//   k_span_write   (start, cell) of every run at its rank -- skipped when the total exceeds the buffer, the host then
//                  copies the plain cells instead (noise frames have no runs to exploit)
#pragma once
#include "flush.cuh"

namespace sloth {

__device__ __forceinline__ uint32_t span_flags(const uint32_t* __restrict__ cells, uint32_t n_cells, uint32_t base, uint32_t cell[4])
{
    // base is a multiple of 4 and the buffer is 16-byte aligned: one 16-byte load, the predecessor from the word before
    uint32_t prev = 0, flags = 0;
    if (base + 4u <= n_cells) {
        const uint4 v = *reinterpret_cast<const uint4*>(cells + base);
        cell[0] = v.x; cell[1] = v.y; cell[2] = v.z; cell[3] = v.w;
    } else {
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) cell[k] = base + k < n_cells ? cells[base + k] : 0u;
    }
    if (base) prev = cells[base - 1u];
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k) {
        if (base + k < n_cells && (base + k == 0u || cell[k] != prev)) flags |= 1u << k;
        prev = cell[k];
    }
    return flags;
}

__global__ void __launch_bounds__(FLUSH_THREADS) k_span_count(const uint32_t* __restrict__ cells, uint32_t n_cells,
                                                              uint32_t* __restrict__ block_sum)
{
    __shared__ uint32_t s_warp[FLUSH_THREADS / 32];
    const uint32_t base = blockIdx.x * FLUSH_CELLS_PER_BLOCK + threadIdx.x * FLUSH_CELLS_PER_THREAD;
    uint32_t cell[4];
    uint32_t n = base < n_cells ? __popc(span_flags(cells, n_cells, base, cell)) : 0u;
    n = __reduce_add_sync(0xFFFFFFFFu, n);
    if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (uint32_t w = 0; w < FLUSH_THREADS / 32; ++w) t += s_warp[w];
        block_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(FLUSH_THREADS) k_span_write(const uint32_t* __restrict__ cells, uint32_t n_cells,
                                                              const unsigned long long* __restrict__ block_off,
                                                              const unsigned long long* __restrict__ total, unsigned long long cap,
                                                              uint2* __restrict__ runs)
{
    __shared__ uint32_t s_warp[FLUSH_THREADS / 32];
    if (*total > cap) return;   // the host falls back to the plain cells
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * FLUSH_CELLS_PER_BLOCK + threadIdx.x * FLUSH_CELLS_PER_THREAD;
    uint32_t cell[4];
    const uint32_t flags = base < n_cells ? span_flags(cells, n_cells, base, cell) : 0u;
    const uint32_t mine = __popc(flags);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t off = incl - mine;
#pragma unroll
    for (uint32_t w = 0; w < FLUSH_THREADS / 32; ++w)
        if (w < warp) off += s_warp[w];
    uint2* out = runs + block_off[blockIdx.x] + off;
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k)
        if ((flags >> k) & 1u) *out++ = make_uint2(base + k, cell[k]);
}

}  // namespace sloth
