// flush.cuh -- Context::flush on the device (src/context.rs:50-92), SURVEY.md 8(f) "next-1".
//
// The reference prints the frame buffer cell by cell; for the -j export that is ~31 bytes of
// text per cell (64 MB per 1920x1080 frame), and formatting it on the host costs far more than
// rendering.  Here the exact byte stream is produced on the GPU: per-cell length, block sums,
// one small scan, then every block formats its 1024 cells into shared memory and copies them
// out with coalesced 16-byte stores.  Modes:
//   0  plain:   glyph                                         (flush(color = false))
//   1  ANSI:    ESC[48;2;25;25;25m ESC[38;2;R;G;Bm glyph ESC[0m   (crossterm 0.18 StyledContent)
//   2  webify:  <span style="color:rgb(R,G,B)">glyph          (flush(color, webify))
// Frame prefixes/suffixes (cursor move, "`\n", "`,\n", println's newline) stay with the host.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sloth {

static constexpr uint32_t FLUSH_CELLS_PER_THREAD = 4;
static constexpr uint32_t FLUSH_THREADS = 256;
static constexpr uint32_t FLUSH_CELLS_PER_BLOCK = FLUSH_CELLS_PER_THREAD * FLUSH_THREADS;   // 1024
static constexpr uint32_t FLUSH_MAX_CELL_BYTES = 44;                                        // ANSI with 3-digit channels

__device__ __forceinline__ uint32_t dec_digits(uint32_t v) { return 1u + (v >= 10u) + (v >= 100u); }

__device__ __forceinline__ uint32_t cell_text_len(uint32_t cell, int mode)
{
    if (mode == 0) return 1u;
    const uint32_t d = dec_digits((cell >> 8) & 0xFFu) + dec_digits((cell >> 16) & 0xFFu) + dec_digits(cell >> 24);
    return (mode == 1 ? 31u : 29u) + d;
}

__device__ __forceinline__ char* put_dec(char* o, uint32_t v)
{
    if (v >= 100u) { *o++ = (char)('0' + v / 100u); v %= 100u; *o++ = (char)('0' + v / 10u); *o++ = (char)('0' + v % 10u); }
    else if (v >= 10u) { *o++ = (char)('0' + v / 10u); *o++ = (char)('0' + v % 10u); }
    else *o++ = (char)('0' + v);
    return o;
}

__device__ __forceinline__ char* put_str(char* o, const char* s, int n)
{
    for (int i = 0; i < n; ++i) o[i] = s[i];
    return o + n;
}

__device__ __forceinline__ void cell_text_write(char* o, uint32_t cell, int mode)
{
    const char glyph = (char)(cell & 0xFFu);
    const uint32_t r = (cell >> 8) & 0xFFu, g = (cell >> 16) & 0xFFu, b = cell >> 24;
    if (mode == 0) { *o = glyph; return; }
    if (mode == 1) {
        o = put_str(o, "\x1b[48;2;25;25;25m\x1b[38;2;", 23);
        o = put_dec(o, r); *o++ = ';';
        o = put_dec(o, g); *o++ = ';';
        o = put_dec(o, b); *o++ = 'm';
        *o++ = glyph;
        put_str(o, "\x1b[0m", 4);
    } else {
        o = put_str(o, "<span style=\"color:rgb(", 23);
        o = put_dec(o, r); *o++ = ',';
        o = put_dec(o, g); *o++ = ',';
        o = put_dec(o, b);
        o = put_str(o, ")\">", 3);
        *o = glyph;
    }
}

// pass 1: text bytes of each block of 1024 cells
__global__ void __launch_bounds__(FLUSH_THREADS) k_flush_sizes(const uint32_t* __restrict__ cells, uint32_t n_cells, int mode,
                                                               uint32_t* __restrict__ block_sum)
{
    __shared__ uint32_t s_warp[FLUSH_THREADS / 32];
    const uint32_t base = blockIdx.x * FLUSH_CELLS_PER_BLOCK + threadIdx.x * FLUSH_CELLS_PER_THREAD;
    uint32_t len = 0;
#pragma unroll
    for (uint32_t k = 0; k < FLUSH_CELLS_PER_THREAD; ++k)
        if (base + k < n_cells) len += cell_text_len(cells[base + k], mode);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) len += __shfl_xor_sync(0xFFFFFFFFu, len, d);
    if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = len;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (uint32_t w = 0; w < FLUSH_THREADS / 32; ++w) t += s_warp[w];
        block_sum[blockIdx.x] = t;
    }
}

// pass 2: exclusive scan of the block sums (one block; n_blocks <= a few 10^4), total to *total_out
__global__ void __launch_bounds__(1024) k_flush_scan(const uint32_t* __restrict__ block_sum, uint32_t n_blocks,
                                                     unsigned long long* __restrict__ block_off,
                                                     unsigned long long* __restrict__ total_out)
{
    __shared__ unsigned long long s_part[1024];
    const uint32_t per = (n_blocks + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per, hi = min(n_blocks, lo + per);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += block_sum[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024u; d <<= 1) {   // Hillis-Steele inclusive scan
        const unsigned long long v = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0ull;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = s_part[threadIdx.x] - sum;   // exclusive prefix of this thread's range
    for (uint32_t i = lo; i < hi; ++i) {
        block_off[i] = run;
        run += block_sum[i];
    }
    if (threadIdx.x == 1023) *total_out = s_part[1023];
}

// pass 3: format 1024 cells into shared memory, then copy the block's bytes out coalesced
__global__ void __launch_bounds__(FLUSH_THREADS) k_flush_write(const uint32_t* __restrict__ cells, uint32_t n_cells, int mode,
                                                               const unsigned long long* __restrict__ block_off,
                                                               char* __restrict__ text)
{
    extern __shared__ __align__(16) char s_text[];   // FLUSH_CELLS_PER_BLOCK * max cell bytes (mode dependent)
    __shared__ uint32_t s_warp[FLUSH_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * FLUSH_CELLS_PER_BLOCK + threadIdx.x * FLUSH_CELLS_PER_THREAD;
    uint32_t cell[FLUSH_CELLS_PER_THREAD], len[FLUSH_CELLS_PER_THREAD], mine = 0;
#pragma unroll
    for (uint32_t k = 0; k < FLUSH_CELLS_PER_THREAD; ++k) {
        cell[k] = base + k < n_cells ? cells[base + k] : 0u;
        len[k] = base + k < n_cells ? cell_text_len(cell[k], mode) : 0u;
        mine += len[k];
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t off = incl - mine, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < FLUSH_THREADS / 32; ++w) {
        const uint32_t c = s_warp[w];
        if (w < warp) off += c;
        total += c;
    }
    // lay the text out in shared memory with the same 16-byte phase as its destination, so the copy
    // below moves aligned 16-byte words on both sides
    char* dst = text + block_off[blockIdx.x];
    const uint32_t shift = (uint32_t)((uintptr_t)dst & 15u);
    off += shift;
#pragma unroll
    for (uint32_t k = 0; k < FLUSH_CELLS_PER_THREAD; ++k) {
        if (len[k]) cell_text_write(s_text + off, cell[k], mode);
        off += len[k];
    }
    __syncthreads();
    const uint32_t head = min((16u - shift) & 15u, total);           // bytes before the first aligned word
    if (threadIdx.x < head) dst[threadIdx.x] = s_text[shift + threadIdx.x];
    const uint32_t body = (total - head) / 16u;
    const uint4* src16 = reinterpret_cast<const uint4*>(s_text + shift + head);   // 16-byte aligned by construction
    uint4* dst16 = reinterpret_cast<uint4*>(dst + head);
    for (uint32_t i = threadIdx.x; i < body; i += FLUSH_THREADS) dst16[i] = src16[i];
    const uint32_t done = head + body * 16u;
    if (threadIdx.x < total - done) dst[done + threadIdx.x] = s_text[shift + done + threadIdx.x];
}

}  // namespace sloth
