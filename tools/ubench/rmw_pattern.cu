// Read-modify-write of an fp32 [M, N] matrix in the access pattern of the GEMM's fp32 epilogue (CTA = 128 x 256 tile, warp = 32 rows,
// quarter-warp = one 128-byte line of a 32-column chunk) against a linear sweep.  nvcc -O3 -arch=sm_100a -o rmw_pattern rmw_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int DEPTH>  // chunks whose loads are issued before the first store
__global__ void __launch_bounds__(256) tile_rmw(float* x, int M, int N, int tiles_n, int num_tiles) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, grp = warp >> 2;
    const int rr0 = lane >> 3, cc = (lane & 7) * 4;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * 128, n0 = (tile % tiles_n) * 256;
        for (int c0 = grp; c0 < 8; c0 += 2 * DEPTH) {
            float4 v[DEPTH][8];
#pragma unroll
            for (int d = 0; d < DEPTH; ++d)
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = m0 + q * 32 + it * 4 + rr0, col = n0 + (c0 + 2 * d) * 32 + cc;
                    v[d][it] = row < M ? *reinterpret_cast<const float4*>(x + (size_t)row * N + col) : make_float4(0, 0, 0, 0);
                }
#pragma unroll
            for (int d = 0; d < DEPTH; ++d)
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = m0 + q * 32 + it * 4 + rr0, col = n0 + (c0 + 2 * d) * 32 + cc;
                    float4 o = v[d][it]; o.x += 1.f; o.y += 1.f; o.z += 1.f; o.w += 1.f;
                    if (row < M) *reinterpret_cast<float4*>(x + (size_t)row * N + col) = o;
                }
        }
    }
}
__global__ void __launch_bounds__(256) linear_rmw(float4* x, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        float4 o = x[i]; o.x += 1.f; o.y += 1.f; o.z += 1.f; o.w += 1.f; x[i] = o;
    }
}
int main() {
    const int M = 65600, N = 1024;
    float* x; cudaMalloc(&x, (size_t)M * N * 4); cudaMemset(x, 0, (size_t)M * N * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int tiles_n = N / 256, num_tiles = ((M + 127) / 128) * tiles_n;
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0); for (int i = 0; i < 20; ++i) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
        printf("%-44s %.4f ms  %6.0f GB/s\n", name, ms, (double)M * N * 8 / ms / 1e6);
    };
    run("linear RMW, 148x8 CTAs", [&] { linear_rmw<<<148 * 8, 256>>>((float4*)x, (size_t)M * N / 4); });
    run("tile pattern, 148 CTAs, 1 chunk in flight", [&] { tile_rmw<1><<<148, 256>>>(x, M, N, tiles_n, num_tiles); });
    run("tile pattern, 148 CTAs, 2 chunks in flight", [&] { tile_rmw<2><<<148, 256>>>(x, M, N, tiles_n, num_tiles); });
    run("tile pattern, 148 CTAs, 4 chunks in flight", [&] { tile_rmw<4><<<148, 256>>>(x, M, N, tiles_n, num_tiles); });
    run("tile pattern, 296 CTAs, 4 chunks in flight", [&] { tile_rmw<4><<<296, 256>>>(x, M, N, tiles_n, num_tiles); });
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
