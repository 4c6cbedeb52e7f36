// A2 Multi-Scale Query Projector (utils/utils_walkgpt.py:220-300).
// GEMM-shaped work (input projection, gate MLP, per-scale K|V projections over all kv tokens, query-side projections and
// FFNs, to_llama) runs on the tcgen05 GEMM; the small pieces live here: grid average pooling, the seg-aware gate
// tail + kv normalisation, the Nq<=12-query cross attention and the token assembly.
#include "internal.h"
#include "ptx.cuh"

namespace wg {
namespace {

constexpr int D = 1024;

// x1 [B, g, g, D] bf16 -> pooled [B, g/s, g/s, D] bf16 (avg_pool2d kernel=stride=s, floor)
__global__ void __launch_bounds__(256) pool_grid_kernel(const __nv_bfloat16* __restrict__ x1, __nv_bfloat16* __restrict__ out, int B, int g, int s) {
    const int gp = g / s;
    const long long total = (long long)B * gp * gp * (D / 8);
    const float inv = 1.0f / (s * s);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % (D / 8));
        const long long tok = i / (D / 8);
        const int b = (int)(tok / (gp * gp));
        const int pos = (int)(tok % (gp * gp));
        const int py = pos / gp, px = pos % gp;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int dy = 0; dy < s; ++dy)
            for (int dx = 0; dx < s; ++dx) {
                uint4 u = *reinterpret_cast<const uint4*>(x1 + (((size_t)b * g + py * s + dy) * g + px * s + dx) * D + c8 * 8);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float2 t = __bfloat1622float2(h[e]);
                    acc[e * 2] += t.x;
                    acc[e * 2 + 1] += t.y;
                }
            }
        uint4 o;
        o.x = pack_bf16x2(acc[0] * inv, acc[1] * inv);
        o.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
        o.z = pack_bf16x2(acc[4] * inv, acc[5] * inv);
        o.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
        *reinterpret_cast<uint4*>(out + tok * D + c8 * 8) = o;
    }
}

// global token: mean over the L tokens of an image.  grid (D/256, B), 256 threads = 8 token-groups x 32 lanes x 8 channels
__global__ void __launch_bounds__(256) global_mean_kernel(const __nv_bfloat16* __restrict__ x1, __nv_bfloat16* __restrict__ out, int L) {
    __shared__ float part[8][256];
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * 256 + (threadIdx.x & 31) * 8;
    const int grp = threadIdx.x >> 5;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = grp; t < L; t += 8) {
        uint4 u = *reinterpret_cast<const uint4*>(x1 + ((size_t)b * L + t) * D + c0);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float2 v = __bfloat1622float2(h[e]);
            acc[e * 2] += v.x;
            acc[e * 2 + 1] += v.y;
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[grp][(threadIdx.x & 31) * 8 + e] = acc[e];
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][threadIdx.x];
    out[(size_t)b * D + blockIdx.x * 256 + threadIdx.x] = __float2bfloat16(s / L);
}

// seg-aware gate tail + kv normalisation, one warp per kv token:
//   gate = sigmoid(h . w2 + b2);  y = gate * x;  xhat = (y - mean) * rsqrt(var + 1e-5)   (kv_norm affine is folded into W_kv)
__global__ void __launch_bounds__(256) gate_scale_norm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ h,
                                                              const float* __restrict__ w2, const float* __restrict__ b2,
                                                              __nv_bfloat16* __restrict__ xhat, long long rows) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float dot = 0.f;
    {
        uint2 u = *reinterpret_cast<const uint2*>(h + row * 128 + lane * 4);
        float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x));
        float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
        float4 w = *reinterpret_cast<const float4*>(w2 + lane * 4);
        dot = a.x * w.x + a.y * w.y + b.x * w.z + b.y * w.w;
    }
    const float gate = 1.0f / (1.0f + expf(-(warp_sum(dot) + b2[0])));
    float v[4][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 u = *reinterpret_cast<const uint4*>(x + row * D + (c * 32 + lane) * 8);
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float2 t = __bfloat1622float2(hh[e]);
            v[c][e * 2] = t.x * gate;
            v[c][e * 2 + 1] = t.y * gate;
            s += v[c][e * 2] + v[c][e * 2 + 1];
        }
    }
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < 8; ++e) q = fmaf(v[c][e] - mean, v[c][e] - mean, q);
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 o;
        o.x = pack_bf16x2((v[c][0] - mean) * rstd, (v[c][1] - mean) * rstd);
        o.y = pack_bf16x2((v[c][2] - mean) * rstd, (v[c][3] - mean) * rstd);
        o.z = pack_bf16x2((v[c][4] - mean) * rstd, (v[c][5] - mean) * rstd);
        o.w = pack_bf16x2((v[c][6] - mean) * rstd, (v[c][7] - mean) * rstd);
        *reinterpret_cast<uint4*>(xhat + row * D + (c * 32 + lane) * 8) = o;
    }
}

// dst[b*nq + i, :] = src[i, :]   (learned queries expanded over the batch), fp32
__global__ void __launch_bounds__(256) broadcast_queries_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int nq) {
    const long long total = (long long)B * nq * (D / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % (D / 4));
        const int qi = (int)((i / (D / 4)) % nq);
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[(size_t)qi * (D / 4) + c4];
    }
}

// Cross attention of nq (<=12) projected queries against Nkv keys, head_dim 128.  grid (B, heads); 256 threads.
//   qp  bf16 [B*nq, D]; kv bf16 [B*Nkv, ldkv] with K at column k_off + h*128 and V at v_off + h*128; out bf16 [B*nq, D]
// fp32 CUDA-core arithmetic (3 GFLOP per launch).  The first version was bound by shared-memory instructions (one scalar
// LDS per FMA for the broadcast query values, one per probability): 0.50 ms per launch at Nkv = 1024.  Here the queries are
// read as float4 and shared by two keys per thread, scores / probabilities live as [key][12] rows (three float4 per key),
// and the softmax statistics are reduced across the block instead of being re-read per query row.
constexpr int MAXQ = 12;
constexpr int QPAD = 12;  // score row pitch (floats): three float4 per key, 48-byte pitch (conflict-free for 16-byte accesses)
__global__ void __launch_bounds__(256) msqp_attention_kernel(const __nv_bfloat16* __restrict__ qp, const __nv_bfloat16* __restrict__ kv, int ldkv,
                                                             int k_off, int v_off, __nv_bfloat16* __restrict__ out, int nq, int Nkv) {
    extern __shared__ __align__(16) float sm[];
    float* qs = sm;                            // [MAXQ][128], rows >= nq zero
    float* sc = sm + MAXQ * 128;               // [Nkv][QPAD] scores, then exp(score - max)
    float* red = sc + (size_t)Nkv * QPAD;      // [4][MAXQ][128] partial outputs; first used as [8 warps][QPAD] reduction scratch
    const int b = blockIdx.x, h = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < MAXQ * 128; i += 256)
        qs[i] = (i / 128 < nq) ? __bfloat162float(qp[((size_t)b * nq + i / 128) * D + h * 128 + (i & 127)]) : 0.f;
    __syncthreads();
    const float scale = 0.08838834764831845f;  // 1/sqrt(128)
    const __nv_bfloat16* kbase = kv + (size_t)b * Nkv * ldkv + k_off + h * 128;
    const __nv_bfloat16* vbase = kv + (size_t)b * Nkv * ldkv + v_off + h * 128;
    // ---- scores: two keys per thread and pass, so every float4 of a query feeds 8 FMAs
    float mx[MAXQ];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) mx[i] = -INFINITY;
    for (int key0 = tid; key0 < Nkv; key0 += 512) {
        const int key1 = key0 + 256;
        const bool has1 = key1 < Nkv;
        float acc0[MAXQ], acc1[MAXQ];
#pragma unroll
        for (int i = 0; i < MAXQ; ++i) acc0[i] = acc1[i] = 0.f;
        const uint4* kr0 = reinterpret_cast<const uint4*>(kbase + (size_t)key0 * ldkv);
        const uint4* kr1 = reinterpret_cast<const uint4*>(kbase + (size_t)(has1 ? key1 : key0) * ldkv);
#pragma unroll 2
        for (int c = 0; c < 16; ++c) {
            const uint4 u0 = __ldg(kr0 + c), u1 = __ldg(kr1 + c);
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&u0);
            const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&u1);
            float k0[8], k1[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 t0 = __bfloat1622float2(h0[e]), t1 = __bfloat1622float2(h1[e]);
                k0[e * 2] = t0.x; k0[e * 2 + 1] = t0.y;
                k1[e * 2] = t1.x; k1[e * 2 + 1] = t1.y;
            }
#pragma unroll
            for (int i = 0; i < MAXQ; ++i) {
                const float4 qa = *reinterpret_cast<const float4*>(qs + i * 128 + c * 8);
                const float4 qb = *reinterpret_cast<const float4*>(qs + i * 128 + c * 8 + 4);
                acc0[i] = fmaf(k0[0], qa.x, acc0[i]); acc0[i] = fmaf(k0[1], qa.y, acc0[i]);
                acc0[i] = fmaf(k0[2], qa.z, acc0[i]); acc0[i] = fmaf(k0[3], qa.w, acc0[i]);
                acc0[i] = fmaf(k0[4], qb.x, acc0[i]); acc0[i] = fmaf(k0[5], qb.y, acc0[i]);
                acc0[i] = fmaf(k0[6], qb.z, acc0[i]); acc0[i] = fmaf(k0[7], qb.w, acc0[i]);
                acc1[i] = fmaf(k1[0], qa.x, acc1[i]); acc1[i] = fmaf(k1[1], qa.y, acc1[i]);
                acc1[i] = fmaf(k1[2], qa.z, acc1[i]); acc1[i] = fmaf(k1[3], qa.w, acc1[i]);
                acc1[i] = fmaf(k1[4], qb.x, acc1[i]); acc1[i] = fmaf(k1[5], qb.y, acc1[i]);
                acc1[i] = fmaf(k1[6], qb.z, acc1[i]); acc1[i] = fmaf(k1[7], qb.w, acc1[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < MAXQ; ++i) {
            acc0[i] *= scale;
            acc1[i] *= scale;
            mx[i] = fmaxf(mx[i], acc0[i]);
            if (has1) mx[i] = fmaxf(mx[i], acc1[i]);
        }
#pragma unroll
        for (int g = 0; g < MAXQ / 4; ++g) {
            *reinterpret_cast<float4*>(sc + (size_t)key0 * QPAD + g * 4) = make_float4(acc0[g * 4], acc0[g * 4 + 1], acc0[g * 4 + 2], acc0[g * 4 + 3]);
            if (has1) *reinterpret_cast<float4*>(sc + (size_t)key1 * QPAD + g * 4) = make_float4(acc1[g * 4], acc1[g * 4 + 1], acc1[g * 4 + 2], acc1[g * 4 + 3]);
        }
    }
    // ---- block-wide max per query
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) mx[i] = warp_max(mx[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < MAXQ; ++i) red[warp * QPAD + i] = mx[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) {
        float m = red[i];
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w * QPAD + i]);
        mx[i] = m;
    }
    __syncthreads();
    // ---- exponentials (each thread on the keys it wrote) + block-wide sums; the 1/sum is applied to the output
    float sum[MAXQ];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) sum[i] = 0.f;
    for (int key = tid; key < Nkv; key += 256) {
#pragma unroll
        for (int g = 0; g < MAXQ / 4; ++g) {
            float4 v = *reinterpret_cast<const float4*>(sc + (size_t)key * QPAD + g * 4);
            v.x = __expf(v.x - mx[g * 4]); v.y = __expf(v.y - mx[g * 4 + 1]); v.z = __expf(v.z - mx[g * 4 + 2]); v.w = __expf(v.w - mx[g * 4 + 3]);
            sum[g * 4] += v.x; sum[g * 4 + 1] += v.y; sum[g * 4 + 2] += v.z; sum[g * 4 + 3] += v.w;
            *reinterpret_cast<float4*>(sc + (size_t)key * QPAD + g * 4) = v;
        }
    }
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) sum[i] = warp_sum(sum[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < MAXQ; ++i) red[warp * QPAD + i] = sum[i];
    }
    __syncthreads();
    float inv[MAXQ];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) {
        float t = red[i];
#pragma unroll
        for (int w = 1; w < 8; ++w) t += red[w * QPAD + i];
        inv[i] = 1.0f / t;
    }
    __syncthreads();
    // ---- out[i][dim] = sum_k p[i][k] V[k][dim]:  thread = (key group 0..3, dim pair 0..63)
    const int dp = tid & 63, grp = tid >> 6;
    float o0[MAXQ], o1[MAXQ];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) o0[i] = o1[i] = 0.f;
#pragma unroll 4
    for (int key = grp; key < Nkv; key += 4) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vbase + (size_t)key * ldkv + dp * 2));
#pragma unroll
        for (int g = 0; g < MAXQ / 4; ++g) {
            const float4 pr = *reinterpret_cast<const float4*>(sc + (size_t)key * QPAD + g * 4);
            o0[g * 4] = fmaf(pr.x, v.x, o0[g * 4]);         o1[g * 4] = fmaf(pr.x, v.y, o1[g * 4]);
            o0[g * 4 + 1] = fmaf(pr.y, v.x, o0[g * 4 + 1]); o1[g * 4 + 1] = fmaf(pr.y, v.y, o1[g * 4 + 1]);
            o0[g * 4 + 2] = fmaf(pr.z, v.x, o0[g * 4 + 2]); o1[g * 4 + 2] = fmaf(pr.z, v.y, o1[g * 4 + 2]);
            o0[g * 4 + 3] = fmaf(pr.w, v.x, o0[g * 4 + 3]); o1[g * 4 + 3] = fmaf(pr.w, v.y, o1[g * 4 + 3]);
        }
    }
#pragma unroll
    for (int i = 0; i < MAXQ; ++i)
        if (i < nq) {
            red[(grp * MAXQ + i) * 128 + dp * 2] = o0[i] * inv[i];
            red[(grp * MAXQ + i) * 128 + dp * 2 + 1] = o1[i] * inv[i];
        }
    __syncthreads();
    for (int i = tid; i < nq * 128; i += 256) {
        const int qi = i / 128, d = i & 127;
        float s = red[(0 * MAXQ + qi) * 128 + d] + red[(1 * MAXQ + qi) * 128 + d] + red[(2 * MAXQ + qi) * 128 + d] + red[(3 * MAXQ + qi) * 128 + d];
        out[((size_t)b * nq + qi) * D + h * 128 + d] = __float2bfloat16(s);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// The same cross attention on the tensor cores (tcgen05 / TMEM / TMA).  One CTA per (image, head, split of <= 1024 keys), two per SM;
// with more than one split (the 4096-token scale of Path B) every CTA leaves an unnormalised partial (m, l, O) and
// msqp_merge_splits_kernel combines them.
// With <= 12 queries the natural orientation is the TRANSPOSED score tile: keys on the MMA's M dimension.
//   warp 0   : TMA producer: Q [16 x 128] once, then the K tiles and the V tiles ([128 keys x 128 dims] = two 64-wide 128B-swizzled
//              boxes, 32 KB) of this head through ONE 3-stage ring -- each K / V byte crosses L2 -> SM exactly once
//   warp 1   : TMEM allocation + MMA issue.  Phase 1: S^T_kb [128 keys x 16 queries] = K_kb Q^T (M = 128, N = 16, 8 k-steps) for every
//              key block into its own 16 tensor-memory columns -- the whole score matrix (<= 8 x 16 columns) stays in TMEM, so the
//              softmax is an exact two-pass one with no running rescale.  Phase 2: O [16 (of 128) x 128] += P_kb V_kb, P as a K-major
//              shared-memory operand, V as the MN-major operand straight from its TMA tile (two N = 64 halves)
//   warps 2-5: softmax, ONE THREAD PER KEY (TMEM lane == key): pass A row maxima per query (16 registers, then shuffles + one shared
//              exchange), pass B exponentials, per-thread partial sums, P^T written as bf16 into the 16 real rows of the P operand.
//              The MMA's M is 128 but only rows < nq mean anything: rows 16..127 of the P operand are never initialised -- the
//              descriptor walks into whatever follows (the K/V ring) and the garbage lands in O lanes that nobody reads.
// Keys past Nkv in the last block (Nkv = 256 / 64 / 1 for the pooled and global scales) are masked to -inf; their V rows are other
// images' rows or TMA zero fill and meet P = 0.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int TC_MAX_NKV = 1024;
constexpr int TC_PART_PITCH = 132;                // floats per partial row: 128 O values, m, l, 2 pad (16-byte rows)
constexpr int TC_OFF_Q = 0;                       // 2 boxes [16 x 64] bf16, 2 KB each
constexpr int TC_OFF_P = 4096;                    // 2 buffers x 2 boxes [16 real rows x 64 keys], 2 KB each
constexpr int TC_OFF_KV = 12288;                  // 3 stages x 32 KB
constexpr int TC_STAGES = 3;
constexpr int TC_STAGE_BYTES = 32768;
constexpr int TC_OFF_BAR = TC_OFF_KV + TC_STAGES * TC_STAGE_BYTES;
constexpr int TC_NUM_BARS = 14;
constexpr int TC_OFF_RED = TC_OFF_BAR + TC_NUM_BARS * 8 + 16;   // [4 warps][16] floats
constexpr int TC_SMEM_BYTES = TC_OFF_RED + 4 * 16 * 4;

__global__ void __launch_bounds__(192, 2)
msqp_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, int k_off, int v_off,
                         __nv_bfloat16* __restrict__ out, int nq, int Nkv_all, int nsplit, float* __restrict__ part) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;    // [3]
    uint64_t* kv_empty = bars + 4;   // [3]
    uint64_t* s_done = bars + 7;
    uint64_t* o_done = bars + 8;
    uint64_t* p_ready = bars + 9;    // [2]
    uint64_t* p_free = bars + 11;    // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + TC_NUM_BARS);
    float* red = reinterpret_cast<float*>(smem + TC_OFF_RED);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sp = blockIdx.x % nsplit, bh = blockIdx.x / nsplit;
    const int b = bh >> 3, h = bh & 7;
    const int key0 = sp * TC_MAX_NKV;                                     // this split's first key
    const int Nkv = min(TC_MAX_NKV, Nkv_all - key0);                       // and its key count
    const int NB = (Nkv + 127) >> 7;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        mbar_init(q_full, 1);
        for (int i = 0; i < TC_STAGES; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        mbar_init(s_done, 1);
        mbar_init(o_done, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&p_ready[i], 128);
            mbar_init(&p_free[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_S = *tmem_ptr_smem;   // 16 columns per key block
    const uint32_t tmem_O = tmem_S + 128;     // 128 columns

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, 4096);
            tma_load_2d(smem + TC_OFF_Q, &tmQ, q_full, h * 128, b * nq);
            tma_load_2d(smem + TC_OFF_Q + 2048, &tmQ, q_full, h * 128 + 64, b * nq);
            for (int i = 0; i < 2 * NB; ++i) {
                const int st = i % TC_STAGES, use = i / TC_STAGES;
                mbar_wait_relaxed(&kv_empty[st], (use & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[st], TC_STAGE_BYTES);
                const int col = (i < NB ? k_off : v_off) + h * 128;
                const int row = b * Nkv_all + key0 + (i < NB ? i : i - NB) * 128;
                uint8_t* dst = smem + TC_OFF_KV + st * TC_STAGE_BYTES;
                tma_load_2d(dst, &tmKV, &kv_full[st], col, row);
                tma_load_2d(dst + 16384, &tmKV, &kv_full[st], col + 64, row);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 16, false, false);
        constexpr uint32_t IDESC_O = umma_idesc_bf16(128, 64, false, true);
        const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + TC_OFF_Q));
        const uint64_t p_desc = umma_desc_sw128(smem_u32(smem + TC_OFF_P));
        const uint64_t kv_desc = umma_desc_sw128(smem_u32(smem + TC_OFF_KV));
        mbar_wait(q_full, 0);
        for (int kb = 0; kb < NB; ++kb) {
            const int st = kb % TC_STAGES;
            mbar_wait(&kv_full[st], (kb / TC_STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t kd = kv_desc + st * (TC_STAGE_BYTES >> 4);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)  // 128 dims = 2 boxes x 4 k-steps of 16
                    umma_f16_ss(tmem_S + kb * 16, kd + (ks >> 2) * (16384 >> 4) + (ks & 3) * 2, q_desc + (ks >> 2) * (2048 >> 4) + (ks & 3) * 2, IDESC_S,
                                ks != 0);
                umma_commit(&kv_empty[st]);
                if (kb + 1 == NB) umma_commit(s_done);
            }
            __syncwarp();
        }
        for (int kb = 0; kb < NB; ++kb) {
            const int i = NB + kb, st = i % TC_STAGES, buf = kb & 1;
            mbar_wait(&kv_full[st], (i / TC_STAGES) & 1);
            mbar_wait(&p_ready[buf], (kb >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t vd = kv_desc + st * (TC_STAGE_BYTES >> 4);
                const uint64_t pd = p_desc + buf * (4096 >> 4);
#pragma unroll
                for (int nh = 0; nh < 2; ++nh)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)  // 128 keys = 8 k-steps; P: 2 boxes x 4 k-steps, V: 2048 B per k-step inside its box
                        umma_f16_ss(tmem_O + nh * 64, pd + (ks >> 2) * (2048 >> 4) + (ks & 3) * 2, vd + nh * (16384 >> 4) + ks * (2048 >> 4), IDESC_O,
                                    (kb | ks) != 0);
                umma_commit(&kv_empty[st]);
                umma_commit(&p_free[buf]);
                if (kb + 1 == NB) umma_commit(o_done);
            }
            __syncwarp();
        }
    } else {
        const int quarter = warp & 3;
        const int t = quarter * 32 + lane;  // key within a block == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const float c = 0.08838834764831845f * 1.4426950408889634f;  // 1/sqrt(128) * log2(e)
        mbar_wait(s_done, 0);
        tc_fence_after();
        // ---- pass A: maxima
        float mx[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) mx[q] = -INFINITY;
        for (int kb = 0; kb < NB; ++kb) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_S + lane_off + kb * 16, v);
            tmem_ld_wait();
            if (kb * 128 + t < Nkv) {
#pragma unroll
                for (int q = 0; q < 16; ++q) mx[q] = fmaxf(mx[q], __uint_as_float(v[q]));
            }
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) mx[q] = warp_max(mx[q]);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) red[quarter * 16 + q] = mx[q];
        }
        named_bar_sync(1, 128);
#pragma unroll
        for (int q = 0; q < 16; ++q) mx[q] = fmaxf(fmaxf(red[q], red[16 + q]), fmaxf(red[32 + q], red[48 + q])) * c;  // key 0 is always valid: finite
        named_bar_sync(1, 128);  // red is reused for the sums
        // ---- pass B: exponentials, partial sums, P^T -> the operand tile (row q, key t)
        float sum[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) sum[q] = 0.f;
        const uint32_t p_elem = static_cast<uint32_t>((t >> 6) * 2048 + (t & 7) * 2);
        const uint32_t p_chunk = static_cast<uint32_t>((t & 63) >> 3);
        for (int kb = 0; kb < NB; ++kb) {
            const int buf = kb & 1;
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_S + lane_off + kb * 16, v);
            tmem_ld_wait();
            const bool valid = kb * 128 + t < Nkv;
            float pr[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                pr[q] = valid ? exp2f(fmaf(__uint_as_float(v[q]), c, -mx[q])) : 0.f;
                sum[q] += pr[q];
            }
            if (kb >= 2) mbar_wait(&p_free[buf], ((kb >> 1) - 1) & 1);
            uint8_t* pb = smem + TC_OFF_P + buf * 4096 + p_elem;
#pragma unroll
            for (int q = 0; q < 16; ++q)
                if (q < nq) *reinterpret_cast<__nv_bfloat16*>(pb + q * 128 + ((p_chunk ^ (q & 7)) << 4)) = __float2bfloat16(pr[q]);
            fence_proxy_async_smem();
            mbar_arrive(&p_ready[buf]);
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) sum[q] = warp_sum(sum[q]);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) red[quarter * 16 + q] = sum[q];
        }
        named_bar_sync(1, 128);
        // ---- epilogue: the warp that owns TMEM lanes 0..31 holds the query rows
        if (quarter == 0) {
            mbar_wait(o_done, 0);
            tc_fence_after();
            const float l_row = red[lane & 15] + red[16 + (lane & 15)] + red[32 + (lane & 15)] + red[48 + (lane & 15)];
            const float inv = 1.0f / l_row;
            __nv_bfloat16* orow = out + ((size_t)b * nq + lane) * D + h * 128;
            // several splits: partial record [16 rows][128 O values | m | l] per (image, head, split), unnormalised, m in log2 units
            float* prow = part != nullptr ? part + (((size_t)bh * nsplit + sp) * 16 + lane) * TC_PART_PITCH : nullptr;
            if (prow != nullptr && lane < nq) {
                float m_row = mx[0];
#pragma unroll
                for (int q = 1; q < 16; ++q) m_row = (lane == q) ? mx[q] : m_row;
                prow[128] = m_row;
                prow[129] = l_row;
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t ov[32];
                tmem_ld_32x32b_x32(tmem_O + cc * 32, ov);  // warp-collective: every lane loads, only real query rows store
                tmem_ld_wait();
                if (prow != nullptr) {
                    if (lane < nq) {
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4)
                            *reinterpret_cast<uint4*>(prow + cc * 32 + g4 * 4) = make_uint4(ov[g4 * 4], ov[g4 * 4 + 1], ov[g4 * 4 + 2], ov[g4 * 4 + 3]);
                    }
                } else if (lane < nq) {
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 0]) * inv, __uint_as_float(ov[g4 * 8 + 1]) * inv);
                        u.y = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 2]) * inv, __uint_as_float(ov[g4 * 8 + 3]) * inv);
                        u.z = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 4]) * inv, __uint_as_float(ov[g4 * 8 + 5]) * inv);
                        u.w = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 6]) * inv, __uint_as_float(ov[g4 * 8 + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + cc * 32 + g4 * 8) = u;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_S);
    }
}

// out[b, q, h*128 + d] = sum_s O_s 2^(m_s - M) / sum_s l_s 2^(m_s - M) over the key splits.  grid B*8, 128 threads (d)
__global__ void __launch_bounds__(128) msqp_merge_splits_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ out, int nq, int nsplit) {
    const int bh = blockIdx.x, b = bh >> 3, h = bh & 7, d = threadIdx.x;
    for (int q = 0; q < nq; ++q) {
        float M = -INFINITY;
        for (int sp = 0; sp < nsplit; ++sp) M = fmaxf(M, part[(((size_t)bh * nsplit + sp) * 16 + q) * TC_PART_PITCH + 128]);
        float o = 0.f, l = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) {
            const float* r = part + (((size_t)bh * nsplit + sp) * 16 + q) * TC_PART_PITCH;
            const float w = exp2f(r[128] - M);
            o = fmaf(r[d], w, o);
            l = fmaf(r[129], w, l);
        }
        out[((size_t)b * nq + q) * D + h * 128 + d] = __float2bfloat16(o / l);
    }
}

// tokens[b, 0..31] = concat of the four query groups, tokens[b, 32..n_tok) = pad_token   (bf16)
struct AssembleArgs {
    const float* q[4];
    int nq[4];
};
__global__ void __launch_bounds__(256) assemble_tokens_kernel(AssembleArgs a, const float* __restrict__ pad, __nv_bfloat16* __restrict__ tokens, int B,
                                                              int n_tok) {
    const long long total = (long long)B * n_tok * (D / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % (D / 4));
        const long long tk = i / (D / 4);
        const int b = (int)(tk / n_tok);
        int t = (int)(tk % n_tok);
        const float* src = pad + c4 * 4;
        for (int s = 0; s < 4; ++s) {
            if (t < a.nq[s]) {
                src = a.q[s] + ((size_t)b * a.nq[s] + t) * D + c4 * 4;
                break;
            }
            t -= a.nq[s];
            if (s == 3) t = -1;
        }
        float4 v = *reinterpret_cast<const float4*>(src);
        uint2 u;
        u.x = pack_bf16x2(v.x, v.y);
        u.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(tokens + tk * D + c4 * 4) = u;
    }
}

struct MsqpBuffers {
    __nv_bfloat16 *xs, *t, *hg, *kv[4], *qn, *qp, *ao, *f, *f1, *tok;
    float* q[4];
    float* part;  // key-split partials of the tensor-core cross attention (scales with more than 1024 keys)
    long long rows[4], row_off[4], R;
};

bool carve(Workspace& ws, const wg_msqp_weights* w, int B, int L, MsqpBuffers& m) {
    const int g = (int)(sqrtf((float)L) + 0.5f);
    m.rows[0] = (long long)B * L;
    m.rows[1] = (long long)B * (g / 2) * (g / 2);
    m.rows[2] = (long long)B * (g / 4) * (g / 4);
    m.rows[3] = B;
    m.R = 0;
    for (int s = 0; s < 4; ++s) {
        m.row_off[s] = m.R;
        m.R += m.rows[s];
    }
    bool ok = true;
    auto take = [&](size_t bytes) {
        void* p = ws.take(bytes);
        ok = ok && p != nullptr;
        return p;
    };
    m.xs = (__nv_bfloat16*)take((size_t)m.R * D * 2);
    m.t = (__nv_bfloat16*)take((size_t)m.R * D * 2);
    m.hg = (__nv_bfloat16*)take((size_t)m.R * 128 * 2);
    for (int s = 0; s < 4; ++s) m.kv[s] = (__nv_bfloat16*)take((size_t)m.rows[s] * 4 * D * 2);
    const size_t qmax = (size_t)B * MAXQ;
    for (int s = 0; s < 4; ++s) m.q[s] = (float*)take((size_t)B * w->scales[s].nq * D * 4);
    m.qn = (__nv_bfloat16*)take(qmax * D * 2);
    m.qp = (__nv_bfloat16*)take(qmax * D * 2);
    m.ao = (__nv_bfloat16*)take(qmax * D * 2);
    m.f = (__nv_bfloat16*)take(qmax * D * 2);
    m.f1 = (__nv_bfloat16*)take(qmax * 4 * D * 2);
    m.tok = (__nv_bfloat16*)take((size_t)B * w->n_tokens * D * 2);
    m.part = (float*)take((size_t)B * 8 * ((L + TC_MAX_NKV - 1) / TC_MAX_NKV) * 16 * TC_PART_PITCH * 4);
    return ok;
}

inline unsigned grid_for(long long work_items) {
    long long blocks = (work_items + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace
}  // namespace wg

using namespace wg;

extern "C" size_t wg_msqp_workspace_bytes(const wg_msqp_weights* w, int B, int L) {
    if (!w || B <= 0 || L <= 0) return 0;
    Workspace ws(nullptr, 0);
    MsqpBuffers m;
    carve(ws, w, B, L, m);
    return ws.used();
}

extern "C" int wg_msqp_forward(const wg_msqp_weights* w, const void* feats_bf16, int B, int L, void* out, int out_is_bf16, void* workspace,
                               size_t workspace_bytes, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(w && feats_bf16 && out && workspace, "wg_msqp_forward: null pointer");
    WG_REQUIRE(B > 0 && L > 0, "wg_msqp_forward: bad sizes");
    const int g = (int)(sqrtf((float)L) + 0.5f);
    WG_REQUIRE(g * g == L, "Token length %d is not a perfect square.", L);  // utils_walkgpt.py:188-192
    WG_REQUIRE(w->d == D && w->heads == 8, "wg_msqp_forward: built for d_proj=1024, 8 heads");
    int nq_total = 0;
    for (int i = 0; i < 4; ++i) {
        WG_REQUIRE(w->scales[i].nq > 0 && w->scales[i].nq <= MAXQ, "wg_msqp_forward: query group size %d unsupported", w->scales[i].nq);
        nq_total += w->scales[i].nq;
    }
    WG_REQUIRE(w->n_tokens >= nq_total, "target_square_side too small");  // utils_walkgpt.py:294
    WG_REQUIRE(g >= 4, "wg_msqp_forward: token grid %dx%d too small for the x4 scale", g, g);
    if (!device_is_sm100()) {
        set_error("wg_msqp_forward: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    Workspace ws(workspace, workspace_bytes);
    MsqpBuffers m;
    WG_REQUIRE(carve(ws, w, B, L, m), "wg_msqp_forward: workspace too small (%zu given)", workspace_bytes);

    // 1. input projection -> x1 (first rows of XS)
    WG_TRY(gemm_bf16_out(feats_bf16, w->sam_dim, w->w_in, (int)m.rows[0], D, w->sam_dim, w->b_in, WG_ACT_NONE, m.xs, D, s));
    // 2. pooled scales + global token
    {
        Prof prof("msqp_pool", s, 0.0, (double)(m.rows[0] * 3 + m.rows[1] + m.rows[2] + m.rows[3]) * D * 2.0, 3);
    pool_grid_kernel<<<grid_for(m.rows[1] * (D / 8)), 256, 0, s>>>(m.xs, m.xs + m.row_off[1] * D, B, g, 2);
    pool_grid_kernel<<<grid_for(m.rows[2] * (D / 8)), 256, 0, s>>>(m.xs, m.xs + m.row_off[2] * D, B, g, 4);
    global_mean_kernel<<<dim3(D / 256, B), 256, 0, s>>>(m.xs, m.xs + m.row_off[3] * D, L);
    }
    WG_CHECK_CUDA(cudaGetLastError());
    // 3-5. seg-aware gate over all kv tokens of all scales, then kv normalisation
    WG_TRY(wg_layernorm(m.xs, 1, D, w->gate_ln_g, w->gate_ln_b, 1e-5f, m.t, D, m.R, D, s));
    WG_TRY(gemm_bf16_out(m.t, D, w->w_g1, (int)m.R, 128, D, w->b_g1, WG_ACT_GELU_ERF, m.hg, 128, s));
    {
        Prof prof("msqp_gate_norm", s, 0.0, (double)m.R * (D * 4.0 + 256.0));
    gate_scale_norm_kernel<<<(unsigned)((m.R + 7) / 8), 256, 0, s>>>(m.xs, m.hg, w->w_g2, w->b_g2, m.t, m.R);
    }
    WG_CHECK_CUDA(cudaGetLastError());
    // 6-7. per scale: K|V projections for both layers in one GEMM, then the two cross-attention blocks on the queries
    for (int sc = 0; sc < 4; ++sc) {
        const wg_msqp_scale& S = w->scales[sc];
        const int Nkv = (int)(m.rows[sc] / B);
        const int Mq = B * S.nq;
        WG_TRY(gemm_bf16_out(m.t + m.row_off[sc] * D, D, S.w_kv, (int)m.rows[sc], 4 * D, D, S.b_kv, WG_ACT_NONE, m.kv[sc], 4 * D, s));
        {
            Prof prof("msqp_broadcast_q", s, 0.0, (double)Mq * D * 4.0);
        broadcast_queries_kernel<<<grid_for((long long)Mq * (D / 4)), 256, 0, s>>>(S.queries, m.q[sc], B, S.nq);
        }
        WG_CHECK_CUDA(cudaGetLastError());
        const size_t smem = (size_t)(MAXQ * 128 + (size_t)QPAD * Nkv + 4 * MAXQ * 128) * sizeof(float);  // CUDA-core kernel (WG_MSQP_TC=0)
        WG_SMEM_OPT_IN(msqp_attention_kernel, 227 * 1024);
        WG_SMEM_OPT_IN(msqp_attention_tc_kernel, TC_SMEM_BYTES);
        // tensor-core kernel (Nkv <= 1024): queries as [16 x 64] boxes of qp, K / V as [128 x 64] boxes of this scale's kv matrix
        static const bool use_tc = [] { const char* e = getenv("WG_MSQP_TC"); return e == nullptr || atoi(e) != 0; }();
        CUtensorMap tmQ, tmKV;
        if (use_tc) {
            WG_TRY(make_tmap_2d_bf16(&tmQ, m.qp, (uint64_t)Mq, D, D, 16, 64));
            WG_TRY(make_tmap_2d_bf16(&tmKV, m.kv[sc], (uint64_t)m.rows[sc], 4 * D, 4 * D, 128, 64));
        }
        for (int l = 0; l < 2; ++l) {
            const wg_msqp_block& Bk = S.blocks[l];
            WG_TRY(wg_layernorm(m.q[sc], 0, D, Bk.qn_g, Bk.qn_b, 1e-5f, m.qn, D, Mq, D, s));
            WG_TRY(gemm_bf16_out(m.qn, D, Bk.w_q, Mq, D, D, Bk.b_q, WG_ACT_NONE, m.qp, D, s));
            {
                Prof prof("msqp_attention", s, 4.0 * B * 8 * (double)S.nq * Nkv * 128, (double)B * Nkv * 2 * D * 2.0);
            if (use_tc) {
                const int nsplit = (Nkv + TC_MAX_NKV - 1) / TC_MAX_NKV;
                msqp_attention_tc_kernel<<<B * 8 * nsplit, 192, TC_SMEM_BYTES, s>>>(tmQ, tmKV, l * 2 * D, l * 2 * D + D, m.ao, S.nq, Nkv, nsplit,
                                                                                   nsplit > 1 ? m.part : nullptr);
                if (nsplit > 1) msqp_merge_splits_kernel<<<B * 8, 128, 0, s>>>(m.part, m.ao, S.nq, nsplit);
            } else {
                WG_REQUIRE(smem <= 227 * 1024, "wg_msqp_forward: %d kv tokens x %d queries exceed shared memory", Nkv, S.nq);
                msqp_attention_kernel<<<dim3(B, 8), 256, smem, s>>>(m.qp, m.kv[sc], 4 * D, l * 2 * D, l * 2 * D + D, m.ao, S.nq, Nkv);
            }
            }
            WG_CHECK_CUDA(cudaGetLastError());
            WG_TRY(gemm_f32_out(m.ao, D, Bk.w_o, Mq, D, D, Bk.b_o, WG_ACT_NONE, m.q[sc], D, m.q[sc], s));
            WG_TRY(wg_layernorm(m.q[sc], 0, D, Bk.ffn_ln_g, Bk.ffn_ln_b, 1e-5f, m.f, D, Mq, D, s));
            WG_TRY(gemm_bf16_out(m.f, D, Bk.w_f1, Mq, 4 * D, D, Bk.b_f1, WG_ACT_GELU_ERF, m.f1, 4 * D, s));
            WG_TRY(gemm_f32_out(m.f1, 4 * D, Bk.w_f2, Mq, D, 4 * D, Bk.b_f2, WG_ACT_NONE, m.q[sc], D, m.q[sc], s));
        }
    }
    // 8. concat + pad tokens, project to the LLM width
    AssembleArgs aa;
    for (int i = 0; i < 4; ++i) {
        aa.q[i] = m.q[i];
        aa.nq[i] = w->scales[i].nq;
    }
    {
        Prof prof("msqp_assemble", s, 0.0, (double)B * w->n_tokens * D * 6.0);
    assemble_tokens_kernel<<<grid_for((long long)B * w->n_tokens * (D / 4)), 256, 0, s>>>(aa, w->pad_token, m.tok, B, w->n_tokens);
    }
    WG_CHECK_CUDA(cudaGetLastError());
    const int Mt = B * w->n_tokens;
    if (out_is_bf16)
        WG_TRY(gemm_bf16_out(m.tok, D, w->w_out, Mt, w->llama_dim, D, w->b_out, WG_ACT_NONE, out, w->llama_dim, s));
    else
        WG_TRY(gemm_f32_out(m.tok, D, w->w_out, Mt, w->llama_dim, D, w->b_out, WG_ACT_NONE, static_cast<float*>(out), w->llama_dim, nullptr, s));
    return WG_OK;
}
