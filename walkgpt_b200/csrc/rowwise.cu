// Bandwidth-bound row-wise kernels of the ViT tower: LayerNorm (fp32/bf16 -> bf16), patch im2col,
// CLS/position-embedding + pre-LayerNorm, CLS-dropping cast of the selected hidden state.
// One warp per row, 16-byte vector accesses, warp-shuffle reductions, rows kept in registers.
#include "host.h"
#include "ptx.cuh"

namespace wg {
namespace {

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x));
    float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4_bf16(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
}

// Each lane owns CH groups of 4 consecutive elements: element index = (c*32 + lane)*4 + e.
template <int CH>
__device__ __forceinline__ void warp_ln_stats(const float (&x)[CH][4], int D, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += (x[c][0] + x[c][1]) + (x[c][2] + x[c][3]);
    mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float d = x[c][e] - mean;
            q = fmaf(d, d, q);
        }
    }
    rstd = rsqrtf(warp_sum(q) / D + eps);
}

template <typename TIn, int CH>
__global__ void __launch_bounds__(256) layernorm_kernel(const TIn* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y,
                                                        long long ldy, long long rows) {
    constexpr int D = CH * 128;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const TIn* xr = x + row * ldx;
    float v[CH][4];
#pragma unroll
    for (int c = 0; c < CH; ++c) load4<TIn>(xr + (c * 32 + lane) * 4, v[c]);
    float mean, rstd;
    warp_ln_stats<CH>(v, D, eps, mean, rstd);
    __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int col = (c * 32 + lane) * 4;
        float o[4];
        float g[4] = {1.f, 1.f, 1.f, 1.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (gamma) load4<float>(gamma + col, g);
        if (beta) load4<float>(beta + col, b);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = (v[c][e] - mean) * rstd * g[e] + b[e];
        store4_bf16(yr + col, o);
    }
}

// Generic-width fallback of the same kernel (D % 128 == 0, any size): three streaming passes.
template <typename TIn>
__global__ void __launch_bounds__(256) layernorm_generic_kernel(const TIn* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps,
                                                                __nv_bfloat16* __restrict__ y, long long ldy, long long rows, int D) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const TIn* xr = x + row * ldx;
    float s = 0.f;
    for (int col = lane * 4; col < D; col += 128) {
        float v[4];
        load4<TIn>(xr + col, v);
        s += (v[0] + v[1]) + (v[2] + v[3]);
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
    for (int col = lane * 4; col < D; col += 128) {
        float v[4];
        load4<TIn>(xr + col, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float d = v[e] - mean;
            q = fmaf(d, d, q);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
    __nv_bfloat16* yr = y + row * ldy;
    for (int col = lane * 4; col < D; col += 128) {
        float v[4], o[4];
        float g[4] = {1.f, 1.f, 1.f, 1.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        load4<TIn>(xr + col, v);
        if (gamma) load4<float>(gamma + col, g);
        if (beta) load4<float>(beta + col, b);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = (v[e] - mean) * rstd * g[e] + b[e];
        store4_bf16(yr + col, o);
    }
}

// pixels [B,3,IMG,IMG] (fp32 or bf16) -> patches bf16 [B*G*G, KPAD]; column k = (c*P + dy)*P + dx, zero for k >= 3*P*P
template <typename TIn>
__global__ void __launch_bounds__(256) im2col_patch_kernel(const TIn* __restrict__ px, __nv_bfloat16* __restrict__ out, int B, int IMG, int P,
                                                           int KPAD) {
    const int G = IMG / P;
    const long long row = blockIdx.x;  // b*G*G + py*G + px
    const int b = (int)(row / (G * G));
    const int pi = (int)(row % (G * G));
    const int py = pi / G, pxi = pi % G;
    const int KV = 3 * P * P;
    for (int k = threadIdx.x; k < KPAD; k += blockDim.x) {
        float v = 0.f;
        if (k < KV) {
            int c = k / (P * P);
            int rem = k - c * P * P;
            int dy = rem / P, dx = rem - dy * P;
            v = static_cast<float>(px[(((size_t)b * 3 + c) * IMG + (py * P + dy)) * IMG + pxi * P + dx]);
        }
        out[row * KPAD + k] = __float2bfloat16(v);
    }
}

// Same mapping, one CTA per (image, patch row): the 3 x P image rows the patch row needs are read as coalesced 32-bit words (two
// pixels; P is even) and scattered into a shared-memory copy of the G output rows, which then goes out with 16-byte stores.
// Needs P % 2 == 0, IMG % 2 == 0 and KPAD % 8 == 0 (CLIP-L/14@448: P 14, IMG 448, KPAD 640); other shapes use the kernel above.
template <typename TIn>
__global__ void __launch_bounds__(256) im2col_patch_rows_kernel(const TIn* __restrict__ px, __nv_bfloat16* __restrict__ out, int IMG, int P, int KPAD) {
    extern __shared__ __align__(16) uint32_t rowbuf[];  // [G][pitch] words; pitch = KPAD / 2 + 4 keeps rows 16-byte aligned and
    const int G = IMG / P;                              // spreads the scattered word stores over the banks
    const int py = blockIdx.x, b = blockIdx.y;
    const int pitch = KPAD / 2 + 4;
    const int HP = P / 2, words_per_row = IMG / 2, kv_words = 3 * P * P / 2;
    for (int i = threadIdx.x; i < 3 * P * words_per_row; i += 256) {
        const int r = i / words_per_row, j = i - r * words_per_row;  // r = c * P + dy, j = word within the image row
        const int c = r / P, dy = r - c * P;
        const int pxi = j / HP, dxp = j - pxi * HP;
        const TIn* src = px + (((size_t)b * 3 + c) * IMG + (py * P + dy)) * IMG + 2 * j;
        uint32_t w;
        if constexpr (sizeof(TIn) == 2) {
            w = __ldg(reinterpret_cast<const uint32_t*>(src));
        } else {
            const float2 a = __ldg(reinterpret_cast<const float2*>(src));
            w = pack_bf16x2(a.x, a.y);
        }
        rowbuf[pxi * pitch + (c * P * P + dy * P) / 2 + dxp] = w;
    }
    const int pad_words = KPAD / 2 - kv_words;
    for (int i = threadIdx.x; i < G * pad_words; i += 256) rowbuf[(i / pad_words) * pitch + kv_words + i % pad_words] = 0u;
    __syncthreads();
    const int vec_per_patch = KPAD / 8;
    __nv_bfloat16* orow = out + ((size_t)b * G * G + (size_t)py * G) * KPAD;
    for (int i = threadIdx.x; i < G * vec_per_patch; i += 256) {
        const int pxi = i / vec_per_patch, k8 = i - pxi * vec_per_patch;
        *reinterpret_cast<uint4*>(orow + (size_t)pxi * KPAD + k8 * 8) = *reinterpret_cast<const uint4*>(rowbuf + pxi * pitch + k8 * 4);
    }
}

// x[b, 0] = LN(cls + pos[0]);  x[b, 1+i] = LN(patch_emb[b*L + i] + pos[1+i])   (D = 1024)
template <int CH>
__global__ void __launch_bounds__(256) embed_ln_kernel(const float* __restrict__ patch_emb, const float* __restrict__ cls,
                                                       const float* __restrict__ pos, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps, float* __restrict__ x, int B, int T) {
    constexpr int D = CH * 128;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * T) return;
    const int b = (int)(row / T), t = (int)(row % T);
    const float* src = t == 0 ? cls : patch_emb + ((size_t)b * (T - 1) + (t - 1)) * D;
    const float* pr = pos + (size_t)t * D;
    float v[CH][4];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        float a[4], pp[4];
        load4<float>(src + (c * 32 + lane) * 4, a);
        load4<float>(pr + (c * 32 + lane) * 4, pp);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[c][e] = a[e] + pp[e];
    }
    float mean, rstd;
    warp_ln_stats<CH>(v, D, eps, mean, rstd);
    float* xr = x + row * D;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int col = (c * 32 + lane) * 4;
        float g[4], bb[4];
        load4<float>(gamma + col, g);
        load4<float>(beta + col, bb);
        float4 o;
        o.x = (v[c][0] - mean) * rstd * g[0] + bb[0];
        o.y = (v[c][1] - mean) * rstd * g[1] + bb[1];
        o.z = (v[c][2] - mean) * rstd * g[2] + bb[2];
        o.w = (v[c][3] - mean) * rstd * g[3] + bb[3];
        *reinterpret_cast<float4*>(xr + col) = o;
    }
}

// out[b, i, :] = x[b, skip + i, :]   (skip = 1 drops CLS: select_feature "patch"; skip = 0 keeps it: "cls_patch"), fp32 -> fp32 or bf16
template <typename TOut>
__global__ void __launch_bounds__(256) drop_cls_cast_kernel(const float* __restrict__ x, TOut* __restrict__ out, int B, int T, int D, int skip) {
    const int To = T - skip;
    const long long n4 = (long long)B * To * (D / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int d4 = (int)(i % (D / 4));
        const long long tok = i / (D / 4);
        const int b = (int)(tok / To), t = (int)(tok % To);
        float4 v = *reinterpret_cast<const float4*>(x + (((size_t)b * T + t + skip) * D) + d4 * 4);
        if constexpr (sizeof(TOut) == 4) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + tok * D + d4 * 4) = v;
        } else {
            uint2 u;
            u.x = pack_bf16x2(v.x, v.y);
            u.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + tok * D + d4 * 4) = u;
        }
    }
}

}  // namespace

int launch_im2col_patch(const void* pixels, int is_bf16, void* out_bf16, int B, int IMG, int P, int KPAD, cudaStream_t s) {
    const int G = IMG / P;
    const long long rows = (long long)B * G * G;
    Prof prof("im2col_patch", s, 0.0, (double)B * 3 * IMG * IMG * (is_bf16 ? 2 : 4) + (double)rows * KPAD * 2);
    const size_t slab_bytes = (size_t)G * (KPAD / 2 + 4) * sizeof(uint32_t);
    if (P % 2 == 0 && IMG % 2 == 0 && KPAD % 8 == 0 && slab_bytes <= 200 * 1024 && B <= 65535) {
        const dim3 grid(G, B);
        if (is_bf16) {
            WG_SMEM_OPT_IN((im2col_patch_rows_kernel<__nv_bfloat16>), 200 * 1024);
            im2col_patch_rows_kernel<__nv_bfloat16><<<grid, 256, slab_bytes, s>>>(static_cast<const __nv_bfloat16*>(pixels), static_cast<__nv_bfloat16*>(out_bf16), IMG, P, KPAD);
        } else {
            WG_SMEM_OPT_IN((im2col_patch_rows_kernel<float>), 200 * 1024);
            im2col_patch_rows_kernel<float><<<grid, 256, slab_bytes, s>>>(static_cast<const float*>(pixels), static_cast<__nv_bfloat16*>(out_bf16), IMG, P, KPAD);
        }
    } else if (is_bf16)
        im2col_patch_kernel<__nv_bfloat16><<<(unsigned)rows, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(pixels),
                                                                         static_cast<__nv_bfloat16*>(out_bf16), B, IMG, P, KPAD);
    else
        im2col_patch_kernel<float><<<(unsigned)rows, 256, 0, s>>>(static_cast<const float*>(pixels), static_cast<__nv_bfloat16*>(out_bf16), B,
                                                                 IMG, P, KPAD);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

int launch_embed_ln(const float* patch_emb, const float* cls, const float* pos, const float* gamma, const float* beta, float eps, float* x,
                    int B, int T, int D, cudaStream_t s) {
    WG_REQUIRE(D == 1024, "embed_ln: hidden size %d not supported (1024 only)", D);
    const long long rows = (long long)B * T;
    Prof prof("embed_ln", s, 0.0, (double)rows * D * 12.0);
    embed_ln_kernel<8><<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(patch_emb, cls, pos, gamma, beta, eps, x, B, T);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

int launch_drop_cls_cast(const float* x, void* out, int out_is_bf16, int B, int T, int D, cudaStream_t s, int keep_cls) {
    const int skip = keep_cls ? 0 : 1;
    const long long n4 = (long long)B * (T - skip) * (D / 4);
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    Prof prof("drop_cls_cast", s, 0.0, (double)n4 * 4 * (out_is_bf16 ? 6.0 : 8.0));
    if (out_is_bf16)
        drop_cls_cast_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>(x, static_cast<__nv_bfloat16*>(out), B, T, D, skip);
    else
        drop_cls_cast_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(x, static_cast<float*>(out), B, T, D, skip);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

}  // namespace wg

extern "C" int wg_layernorm(const void* x, int x_is_bf16, int64_t ldx, const float* gamma, const float* beta, float eps, void* y_bf16,
                            int64_t ldy, int64_t rows, int D, void* stream_) {
    using namespace wg;
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(x && y_bf16, "wg_layernorm: null pointer");
    WG_REQUIRE(D > 0 && D % 128 == 0 && D <= 8192, "wg_layernorm: D=%d must be a multiple of 128 and <= 8192", D);
    WG_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "wg_layernorm: leading dimensions must be multiples of 4");
    if (rows <= 0) return WG_OK;
    if (!device_is_sm100()) {
        set_error("wg_layernorm: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const unsigned grid = (unsigned)((rows + 7) / 8);
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(y_bf16);
    Prof prof("layernorm", s, 0.0, (double)rows * D * ((x_is_bf16 ? 2.0 : 4.0) + 2.0));
#define WG_LN_CASE(CHN)                                                                                                         \
    case CHN:                                                                                                                   \
        if (x_is_bf16)                                                                                                          \
            layernorm_kernel<__nv_bfloat16, CHN><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, eps, y, ldy, rows); \
        else                                                                                                                    \
            layernorm_kernel<float, CHN><<<grid, 256, 0, s>>>(static_cast<const float*>(x), ldx, gamma, beta, eps, y, ldy, rows);  \
        break;
    switch (D / 128) {
        WG_LN_CASE(2)
        WG_LN_CASE(8)
        default:
            if (x_is_bf16)
                layernorm_generic_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, eps, y, ldy, rows, D);
            else
                layernorm_generic_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), ldx, gamma, beta, eps, y, ldy, rows, D);
    }
#undef WG_LN_CASE
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}
