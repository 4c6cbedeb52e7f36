// Micro-benchmark (round 2): is a packed two-wide exponential worth it for the attention softmax?
//   ex2.approx.ftz.bf16x2 / ex2.approx.f16x2 (two exponentials per MUFU instruction?) against ex2.approx.ftz.f32,
//   the 3-input max (max.f32 d, a, b, c), and the accuracy of the packed exponentials over the softmax range [-16, 0].
#include <cstdio>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2_bf16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2_f16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float y; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
template <int MODE>
__global__ void k(float* out, int iters) {
    float a[8];
    unsigned u[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i * 0.1f - 3.0f; u[i] = 0xbf80bf80u + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ex2(a[i]) - 1.5f;
            if (MODE == 1) u[i] = ex2_bf16x2(u[i]) ^ 0x80008000u;   // 2^x in (0.5, 1] -> negated again: stays in range
            if (MODE == 2) u[i] = ex2_f16x2(u[i]) ^ 0x80008000u;
            if (MODE == 3) a[i] = max3(a[i], a[(i + 1) & 7], a[(i + 2) & 7]) + 0.25f;
            if (MODE == 4) a[i] = fmaxf(fmaxf(a[i], a[(i + 1) & 7]), a[(i + 2) & 7]) + 0.25f;
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void acc(float* err) {  // max relative error of the packed exponentials against exp2f over [-16, 0], inputs exactly representable
    float eb = 0.f, eh = 0.f;
    for (int i = threadIdx.x; i < 16 * 256; i += blockDim.x) {
        const float x = -(float)i / 256.0f;
        __nv_bfloat162 xb = __floats2bfloat162_rn(x, x);
        __half2 xh = __floats2half2_rn(x, x);
        unsigned rb = ex2_bf16x2(*reinterpret_cast<unsigned*>(&xb)), rh = ex2_f16x2(*reinterpret_cast<unsigned*>(&xh));
        const float fb = __bfloat162float(reinterpret_cast<__nv_bfloat162*>(&rb)->x), fh = __half2float(reinterpret_cast<__half2*>(&rh)->x);
        const float xbr = __bfloat162float(xb.x), xhr = __half2float(xh.x);
        eb = fmaxf(eb, fabsf(fb - exp2f(xbr)) / exp2f(xbr));
        eh = fmaxf(eh, fabsf(fh - exp2f(xhr)) / exp2f(xhr));
    }
    atomicMax(reinterpret_cast<int*>(err), __float_as_int(eb));
    atomicMax(reinterpret_cast<int*>(err + 1), __float_as_int(eh));
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    const char* names[5] = {"ex2.approx.ftz.f32 (+FADD)", "ex2.approx.ftz.bf16x2 (+LOP)", "ex2.approx.f16x2 (+LOP)", "max3 (+FADD)", "2 x max (+FADD)"};
    for (int mode = 0; mode < 5; ++mode) {
        int iters = 4096;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 2, 1024>>>(out, iters);
            if (mode == 1) k<1><<<148 * 2, 1024>>>(out, iters);
            if (mode == 2) k<2><<<148 * 2, 1024>>>(out, iters);
            if (mode == 3) k<3><<<148 * 2, 1024>>>(out, iters);
            if (mode == 4) k<4><<<148 * 2, 1024>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = 148.0 * 2 * 1024 * iters * 8;
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("%-32s %.3f ms  = %.2f warp-lane instr/clk/SM at %.0f MHz (nominal)\n", names[mode], ms, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1e3);
    }
    float* err; cudaMalloc(&err, 8); cudaMemset(err, 0, 8);
    acc<<<1, 256>>>(err);
    float h[2]; cudaMemcpy(h, err, 8, cudaMemcpyDeviceToHost);
    printf("max relative error over [-16, 0]: bf16x2 %.3e (bf16 ulp/2 = 1.95e-3)   f16x2 %.3e (f16 ulp/2 = 4.9e-4)\n", h[0], h[1]);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
