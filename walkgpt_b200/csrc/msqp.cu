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
        const size_t smem = (size_t)(MAXQ * 128 + (size_t)QPAD * Nkv + 4 * MAXQ * 128) * sizeof(float);
        WG_REQUIRE(smem <= 227 * 1024, "wg_msqp_forward: %d kv tokens x %d queries exceed shared memory", Nkv, S.nq);
        WG_SMEM_OPT_IN(msqp_attention_kernel, 227 * 1024);
        for (int l = 0; l < 2; ++l) {
            const wg_msqp_block& Bk = S.blocks[l];
            WG_TRY(wg_layernorm(m.q[sc], 0, D, Bk.qn_g, Bk.qn_b, 1e-5f, m.qn, D, Mq, D, s));
            WG_TRY(gemm_bf16_out(m.qn, D, Bk.w_q, Mq, D, D, Bk.b_q, WG_ACT_NONE, m.qp, D, s));
            {
                Prof prof("msqp_attention", s, 4.0 * B * 8 * (double)S.nq * Nkv * 128, (double)B * Nkv * 2 * D * 2.0);
            msqp_attention_kernel<<<dim3(B, 8), 256, smem, s>>>(m.qp, m.kv[sc], 4 * D, l * 2 * D, l * 2 * D + D, m.ao, S.nq, Nkv);
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
