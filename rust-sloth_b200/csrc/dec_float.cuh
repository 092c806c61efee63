// dec_float.cuh -- decimal text -> f32, correctly rounded (round to nearest, ties to even), as Rust's
// f32::from_str (tobj 3.2.2, src/inputs.rs:108) and glibc strtof (host/mesh_io.cpp) do it.  Used by the device
// loader (loader.cuh); free of CUDA headers so that the CPU test suite can run the very same code on the host
// (host/host_capi.cpp: sloth_host_parse_f32, tests/test_host_loaders.py).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define LD_HD __host__ __device__
#else
#define LD_HD
#endif

namespace sloth {
namespace ld {

enum : uint32_t { TOK_OK = 0, TOK_BAD = 1, TOK_UNSUPPORTED = 2 };

LD_HD inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }

#define SLOTH_POW5_TABLE                                                                                             \
    {1ull, 5ull, 25ull, 125ull, 625ull, 3125ull, 15625ull, 78125ull, 390625ull, 1953125ull, 9765625ull, 48828125ull, \
     244140625ull, 1220703125ull, 6103515625ull, 30517578125ull, 152587890625ull, 762939453125ull,                  \
     3814697265625ull, 19073486328125ull, 95367431640625ull, 476837158203125ull, 2384185791015625ull,               \
     11920928955078125ull, 59604644775390625ull, 298023223876953125ull, 1490116119384765625ull,                     \
     7450580596923828125ull}

#ifdef __CUDACC__
__constant__ unsigned long long POW5_DEV[28] = SLOTH_POW5_TABLE;
#endif

LD_HD inline unsigned long long pow5(int k)   // 5^k, k in [0, 27] (5^27 < 2^63)
{
#ifdef __CUDA_ARCH__
    return POW5_DEV[k];
#else
    static const unsigned long long table[28] = SLOTH_POW5_TABLE;
    return table[k];
#endif
}

LD_HD inline int clz64(unsigned long long x)
{
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

// w * 10^e10 (w != 0, |e10| <= 27) rounded to nearest-even f32, exactly: the value is held as a 128-bit integer
// times a power of two plus a sticky bit (the remainder of the division by 5^k), so no double rounding occurs.
// Rust's f32::from_str (tobj) and glibc strtof (host loader) are both correctly rounded.
LD_HD inline float dec_to_f32(unsigned long long w, int e10)
{
    unsigned __int128 N;
    int bexp;
    bool sticky = false;
    if (e10 >= 0) {
        N = (unsigned __int128)w * pow5(e10);   // < 2^64 * 2^63
        bexp = e10;
    } else {
        const int lz = clz64(w);
        const unsigned __int128 num = (unsigned __int128)(w << lz) << 64;
        const unsigned long long d = pow5(-e10);
        N = num / d;                             // >= 2^64: plenty of quotient bits
        sticky = (num % d) != 0;
        bexp = -64 - lz + e10;
    }
    const unsigned long long hi = (unsigned long long)(N >> 64), lo = (unsigned long long)N;
    const int p = hi ? 127 - clz64(hi) : 63 - clz64(lo);
    uint32_t mant;
    int shift = 0;
    if (p <= 23) {
        mant = (uint32_t)lo;
    } else {
        shift = p - 23;
        mant = (uint32_t)(N >> shift);
        const unsigned __int128 rest = N & (((unsigned __int128)1 << shift) - 1);
        const unsigned __int128 half = (unsigned __int128)1 << (shift - 1);
        if (rest > half || (rest == half && (sticky || (mant & 1u)))) ++mant;
    }
    // mant <= 2^24 is exact in f32; the scale is a power of two and the supported range never reaches the
    // denormals, so ldexpf is exact (and overflows to +inf like both reference parsers)
    return ldexpf((float)mant, shift + bexp);
}

// [+-]digits[.digits][(e|E)[+-]digits] over the whole token
LD_HD inline uint32_t parse_float(const unsigned char* p, const unsigned char* end, float& out)
{
    bool neg = false;
    if (p < end && (*p == '+' || *p == '-')) { neg = *p == '-'; ++p; }
    if (p < end && ((*p | 0x20) == 'i' || (*p | 0x20) == 'n')) return TOK_UNSUPPORTED;   // inf / infinity / nan
    unsigned long long w = 0;
    int nd = 0, e10 = 0;
    bool any = false, inexact = false;
    for (; p < end && is_digit(*p); ++p) {
        const uint32_t d = *p - '0';
        any = true;
        if (nd < 19) { w = w * 10ull + d; if (w) ++nd; }
        else { ++e10; inexact |= d != 0; if (e10 > 4096) return TOK_UNSUPPORTED; }
    }
    if (p < end && *p == '.') {
        ++p;
        for (; p < end && is_digit(*p); ++p) {
            const uint32_t d = *p - '0';
            any = true;
            if (nd < 19) { w = w * 10ull + d; if (w) ++nd; --e10; if (e10 < -4096) return TOK_UNSUPPORTED; }
            else inexact |= d != 0;
        }
    }
    if (!any) return TOK_BAD;
    if (p < end && (*p == 'e' || *p == 'E')) {
        ++p;
        bool eneg = false;
        if (p < end && (*p == '+' || *p == '-')) { eneg = *p == '-'; ++p; }
        if (!(p < end && is_digit(*p))) return TOK_BAD;
        int ex = 0;
        for (; p < end && is_digit(*p); ++p) ex = ex < 100000 ? ex * 10 + (int)(*p - '0') : ex;
        e10 += eneg ? -ex : ex;
    }
    if (p != end) return TOK_BAD;
    if (w == 0) { out = neg ? -0.0f : 0.0f; return TOK_OK; }
    if (e10 < -27 || e10 > 27) return TOK_UNSUPPORTED;
    float v = dec_to_f32(w, e10);
    // digits beyond the 19th were dropped: the value lies in (w, w+1) * 10^e10; decided only when both ends agree
    if (inexact && dec_to_f32(w + 1ull, e10) != v) return TOK_UNSUPPORTED;
    out = neg ? -v : v;
    return TOK_OK;
}

}  // namespace ld
}  // namespace sloth
