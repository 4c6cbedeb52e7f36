// Micro-benchmark: per-SM throughput of MUFU.EX2, F2FP (f32x2->bf16x2), FFMA, and the polynomial exp2, with many warps.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float poly(float x) {
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float pz = fmaf(0.05500893f, f, 0.24221096f);
    pz = fmaf(pz, f, 0.69328293f);
    pz = fmaf(pz, f, 1.0f);
    return __int_as_float(__float_as_int(pz) + (__float_as_int(t) << 23));
}
template <int MODE>
__global__ void k(float* out, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.1f - 3.0f;
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ex2(a[i]) - 1.5f;
            if (MODE == 1) { __nv_bfloat162 t = __floats2bfloat162_rn(a[i], a[(i + 1) & 7]); acc += *reinterpret_cast<unsigned*>(&t); a[i] += 1.0f; }
            if (MODE == 2) a[i] = fmaf(a[i], 1.0001f, 0.5f);
            if (MODE == 3) a[i] = poly(a[i]) - 1.5f;
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    const char* names[4] = {"MUFU.EX2 (+FADD)", "F2FP.BF16 pack (+IADD,FADD)", "FFMA", "poly exp2 (+FADD)"};
    for (int mode = 0; mode < 4; ++mode) {
        int iters = 4096;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 2, 1024>>>(out, iters);
            if (mode == 1) k<1><<<148 * 2, 1024>>>(out, iters);
            if (mode == 2) k<2><<<148 * 2, 1024>>>(out, iters);
            if (mode == 3) k<3><<<148 * 2, 1024>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = 148.0 * 2 * 1024 * iters * 8;
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("%-30s %.3f ms  %.1f Gop/s  = %.2f lane-ops/clk/SM at %.0f MHz (nominal)\n", names[mode], ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1e3);
    }
    return 0;
}
