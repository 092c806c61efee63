// Microbenchmark: scalar FMUL/FADD vs packed __fmul2_rn/__fadd2_rn throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, float a, float b, int iters)
{
    float2 x0 = make_float2(threadIdx.x * 1e-3f, 1.0f), x1 = make_float2(2.0f, 3.0f);
    float2 x2 = make_float2(0.5f, 0.25f), x3 = make_float2(1.5f, 2.5f);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {   // scalar: 8 mul + 8 add per iteration
            x0.x = __fadd_rn(__fmul_rn(x0.x, a), b); x0.y = __fadd_rn(__fmul_rn(x0.y, a), b);
            x1.x = __fadd_rn(__fmul_rn(x1.x, a), b); x1.y = __fadd_rn(__fmul_rn(x1.y, a), b);
            x2.x = __fadd_rn(__fmul_rn(x2.x, a), b); x2.y = __fadd_rn(__fmul_rn(x2.y, a), b);
            x3.x = __fadd_rn(__fmul_rn(x3.x, a), b); x3.y = __fadd_rn(__fmul_rn(x3.y, a), b);
        } else if (MODE == 2) {   // packed, fusion-proof: every mul / add is its own FFMA2 with an identity operand
            const float2 NZ = make_float2(-0.0f, -0.0f), ONE = make_float2(1.0f, 1.0f);
            x0 = __ffma2_rn(__ffma2_rn(x0, A, NZ), ONE, B);
            x1 = __ffma2_rn(__ffma2_rn(x1, A, NZ), ONE, B);
            x2 = __ffma2_rn(__ffma2_rn(x2, A, NZ), ONE, B);
            x3 = __ffma2_rn(__ffma2_rn(x3, A, NZ), ONE, B);
        } else {           // packed: 4 mul2 + 4 add2 per iteration (same flops) -- ptxas FUSES these into 4 FFMA2
            x0 = __fadd2_rn(__fmul2_rn(x0, A), B);
            x1 = __fadd2_rn(__fmul2_rn(x1, A), B);
            x2 = __fadd2_rn(__fmul2_rn(x2, A), B);
            x3 = __fadd2_rn(__fmul2_rn(x3, A), B);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y;
}

int main()
{
    float* d;
    cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, 0.999f, 0.001f, iters);
            else if (mode == 1) k<1><<<148 * 8, 256>>>(d, 0.999f, 0.001f, iters);
            else k<2><<<148 * 8, 256>>>(d, 0.999f, 0.001f, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 16.0 * iters * 148 * 8 * 256;
            printf("mode %d (%s): %.3f ms, %.2f Tflop/s (non-fused ops)\n", mode, mode == 0 ? "scalar" : mode == 1 ? "packed mul2+add2 (ptxas fuses to FFMA2!)" : "packed, unfused (FFMA2 with identity operands)", ms, flops / ms / 1e9);
        }
    return 0;
}
