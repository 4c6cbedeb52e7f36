// A6 + A7: prompt encoder (text-embedding path) and the SAM-style two-way mask decoder, batched over ALL prompts of a
// batch of images (the reference loops per image and per prompt: model/walkgpt.py:511-535, mask_decoder*.py).
//
//   image side (P*hw rows x 256): every projection over the hw image tokens is one tcgen05 GEMM over all prompts:
//       [K_t2i | V_t2i | Q_i2t] = keys W^T + (pe W^T + b)   -- the positional term is a per-position bias table
//       keys' = LayerNorm4(keys + attn_i2t W_o^T + b)        -- residual + LayerNorm fused in the GEMM epilogue
//       upscaled = keys W_up^T + b                           -- ConvTranspose2d(k=2,s=2) as a GEMM with N = 4 sub-pixels x 32
//   token side (6 tokens per prompt): one CTA per prompt runs self-attention, token->image attention, MLP, the
//   LayerNorms, the hypernetwork MLPs and the IoU head out of shared memory (transposed bf16 weights, fp32 math).
//
// Reference: segment_anything/modeling/transformer.py:62-106,151-182,220-242; mask_decoder_multi_scale.py:137-213;
// prompt_encoder.py:140-186 (text_embeds passthrough + no_mask_embed dense embedding), :67-76 (dense PE, folded into
// the bias tables at weight-pack time).
#include <stdlib.h>

#include "internal.h"
#include "ptx.cuh"

namespace wg {
namespace {

constexpr int C = 256;   // transformer dim
constexpr int NT = 6;    // tokens per prompt: iou, 4 mask tokens, 1 text embedding
constexpr int CI = 128;  // cross-attention internal dim (downsample 2)
constexpr int TK_THREADS = 256;

__device__ __forceinline__ float2 ldbf2(const __nv_bfloat16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

// out[t][n] = act( base + sum_k in[t][k] * Wt[k][n] ), t < NT.   in/out in shared memory, Wt bf16 [K][N] in global.
// base = bias[n] (+ out[t][n] when `accumulate`).  Split-K over thread groups with shared-memory atomics.
__device__ void tok_linear(const float* in, int ldi, int K, const __nv_bfloat16* __restrict__ Wt, const float* __restrict__ bias, float* out,
                           int ldo, int N, bool accumulate, bool relu) {
    const int tid = threadIdx.x;
    for (int i = tid; i < NT * N; i += TK_THREADS) {
        const int t = i / N, n = i - t * N;
        float b = bias ? bias[n] : 0.f;
        out[t * ldo + n] = accumulate ? out[t * ldo + n] + b : b;
    }
    __syncthreads();
    const int pairs = N >> 1;
    int ks = TK_THREADS / pairs;
    if (ks < 1) ks = 1;
    if (ks > K) ks = K;
    const int klen = (K + ks - 1) / ks;
    for (int item = tid; item < pairs * ks; item += TK_THREADS) {
        const int pr = item % pairs, kpart = item / pairs;
        const int k0 = kpart * klen;
        const int k1 = min(K, k0 + klen);
        float a0[NT], a1[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) a0[t] = a1[t] = 0.f;
        const __nv_bfloat16* wp = Wt + (size_t)k0 * N + pr * 2;
#pragma unroll 4
        for (int k = k0; k < k1; ++k, wp += N) {
            const float2 w = ldbf2(wp);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float x = in[t * ldi + k];
                a0[t] = fmaf(x, w.x, a0[t]);
                a1[t] = fmaf(x, w.y, a1[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            atomicAdd(&out[t * ldo + pr * 2], a0[t]);
            atomicAdd(&out[t * ldo + pr * 2 + 1], a1[t]);
        }
    }
    __syncthreads();
    if (relu) {
        for (int i = tid; i < NT * N; i += TK_THREADS) {
            const int t = i / N, n = i - t * N;
            out[t * ldo + n] = fmaxf(out[t * ldo + n], 0.f);
        }
        __syncthreads();
    }
}

// in-place LayerNorm over C=256 for the NT token rows (warp t handles row t)
__device__ void tok_layernorm(float* x, const float* __restrict__ g, const float* __restrict__ b, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < NT) {
        float* r = x + warp * C;
        float v[8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = r[lane + i * 32];
            s += v[i];
        }
        const float mean = warp_sum(s) * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) q = fmaf(v[i] - mean, v[i] - mean, q);
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[lane + i * 32] = (v[i] - mean) * rstd * g[lane + i * 32] + b[lane + i * 32];
    }
    __syncthreads();
}

__device__ void tok_add(float* dst, const float* a, const float* b, int n) {  // dst = a + b
    for (int i = threadIdx.x; i < n; i += TK_THREADS) dst[i] = a[i] + b[i];
    __syncthreads();
}

// token self-attention core: Q,K,V [NT][256] in smem, 8 heads x 32 -> out [NT][256]
__device__ void tok_self_attn_core(const float* Q, const float* K, const float* V, float* out, float* scratch /* >= 8*NT*NT */) {
    const int tid = threadIdx.x;
    for (int i = tid; i < 8 * NT * NT; i += TK_THREADS) {
        const int h = i / (NT * NT), qi = (i / NT) % NT, kj = i % NT;
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e) d = fmaf(Q[qi * C + h * 32 + e], K[kj * C + h * 32 + e], d);
        scratch[i] = d * 0.17677669529663687f;  // 1/sqrt(32)
    }
    __syncthreads();
    for (int i = tid; i < 8 * NT; i += TK_THREADS) {
        float* r = scratch + i * NT;
        float mx = r[0];
#pragma unroll
        for (int j = 1; j < NT; ++j) mx = fmaxf(mx, r[j]);
        float e[NT], s = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            e[j] = __expf(r[j] - mx);
            s += e[j];
        }
        const float inv = 1.0f / s;
#pragma unroll
        for (int j = 0; j < NT; ++j) r[j] = e[j] * inv;
    }
    __syncthreads();
    for (int i = tid; i < NT * C; i += TK_THREADS) {
        const int qi = i / C, c = i % C, h = c / 32;
        const float* p = scratch + (h * NT + qi) * NT;
        float o = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) o = fmaf(p[j], V[j * C + c], o);
        out[i] = o;
    }
    __syncthreads();
}

// token -> image attention: Qt [NT][128] (smem), K/V rows in global (bf16, row pitch ldkv), 8 heads x 16.  A [NT][128] out (smem)
__device__ void tok_t2i_attention(const float* Qt, const __nv_bfloat16* __restrict__ kv, int ldkv, int k_off, int v_off, int hw, float* A,
                                  float* sc /* [NT][hw] */) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < NT * CI; i += TK_THREADS) A[i] = 0.f;
    for (int h = 0; h < 8; ++h) {
        float qh[NT][16];
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int e = 0; e < 16; ++e) qh[t][e] = Qt[t * CI + h * 16 + e];
        for (int key = tid; key < hw; key += TK_THREADS) {
            const uint4* kr = reinterpret_cast<const uint4*>(kv + (size_t)key * ldkv + k_off + h * 16);
            float kf[16];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint4 u = kr[c];
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float2 f = __bfloat1622float2(hh[e]);
                    kf[c * 8 + e * 2] = f.x;
                    kf[c * 8 + e * 2 + 1] = f.y;
                }
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                float d = 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e) d = fmaf(qh[t][e], kf[e], d);
                sc[t * hw + key] = d * 0.25f;  // 1/sqrt(16)
            }
        }
        __syncthreads();
        if (warp < NT) {
            float* r = sc + warp * hw;
            float mx = -INFINITY;
            for (int k = lane; k < hw; k += 32) mx = fmaxf(mx, r[k]);
            mx = warp_max(mx);
            float s = 0.f;
            for (int k = lane; k < hw; k += 32) {
                float e = __expf(r[k] - mx);
                r[k] = e;
                s += e;
            }
            s = warp_sum(s);
            const float inv = 1.0f / s;
            for (int k = lane; k < hw; k += 32) r[k] *= inv;
        }
        __syncthreads();
        {
            const int d = tid & 15, grp = tid >> 4;  // 16 dims x 16 key groups
            float o[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) o[t] = 0.f;
            for (int key = grp; key < hw; key += 16) {
                const float v = __bfloat162float(kv[(size_t)key * ldkv + v_off + h * 16 + d]);
#pragma unroll
                for (int t = 0; t < NT; ++t) o[t] = fmaf(sc[t * hw + key], v, o[t]);
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) atomicAdd(&A[t * CI + h * 16 + d], o[t]);
        }
        __syncthreads();
    }
}

struct TokArgs {
    int phase, hw, n_mask_tokens;
    wg_twoway_layer L;                   // weights of this layer (phases 0/1)
    // final-phase weights
    const void* fin_wq_t; const float* fin_bq; const void* fin_wo_t; const float* fin_bo; const float* nf_g; const float* nf_b;
    const void *hyp_w0_t, *hyp_w1_t, *hyp_w2_t; const float *hyp_b0, *hyp_b1, *hyp_b2;
    const void *iou_w0_t, *iou_w1_t, *iou_w2_t; const float *iou_b0, *iou_b1, *iou_b2;
    const float* out_tokens;   // [1+n_mask][256]
    const float* sparse_add;   // [256] or null
    const float* txt;          // [P][256]
    const __nv_bfloat16* kv;   // image-side projections of this phase, [P*hw][ldkv]
    int ldkv;
    float* Tq;                 // [P][NT][256] running queries
    float* Tpe;                // [P][NT][256] token positional term (= initial tokens)
    float* KT; float* VT;      // [P][NT][128] token-side K/V for image->token attention
    float* hyper;              // [P][n_mask][32]
    float* iou;                // [P][n_mask]
};

__global__ void __launch_bounds__(TK_THREADS) decoder_token_kernel(const TokArgs a) {
    extern __shared__ float sm[];
    float* q = sm;                 // [NT][256]
    float* qpe = q + NT * C;       // [NT][256]
    float* b0 = qpe + NT * C;      // scratch [NT][256]
    float* b1 = b0 + NT * C;
    float* b2 = b1 + NT * C;
    float* b3 = b2 + NT * C;
    float* big = b3 + NT * C;      // [NT][2048]
    float* sc = big + NT * 2048;   // [NT][hw]
    const int p = blockIdx.x;
    const int tid = threadIdx.x;
    const __nv_bfloat16* kv = a.kv + (size_t)p * a.hw * a.ldkv;
    const __nv_bfloat16* W;

    if (a.phase == 0) {
        for (int i = tid; i < NT * C; i += TK_THREADS) {
            const int t = i / C, c = i % C;
            float v = (t < NT - 1) ? a.out_tokens[t * C + c] : a.txt[(size_t)p * C + c] + (a.sparse_add ? a.sparse_add[c] : 0.f);
            q[i] = v;
            qpe[i] = v;
            a.Tpe[(size_t)p * NT * C + i] = v;
        }
    } else {
        for (int i = tid; i < NT * C; i += TK_THREADS) {
            q[i] = a.Tq[(size_t)p * NT * C + i];
            qpe[i] = a.Tpe[(size_t)p * NT * C + i];
        }
    }
    __syncthreads();

    if (a.phase < 2) {
        const wg_twoway_layer& L = a.L;
        // ---- (1) self attention on the tokens
        if (a.phase == 0) {
            for (int i = tid; i < NT * C; i += TK_THREADS) b3[i] = q[i];
            __syncthreads();
        } else {
            tok_add(b3, q, qpe, NT * C);
        }
        tok_linear(b3, C, C, (const __nv_bfloat16*)L.sa_wq_t, L.sa_bq, b0, C, C, false, false);
        tok_linear(b3, C, C, (const __nv_bfloat16*)L.sa_wk_t, L.sa_bk, b1, C, C, false, false);
        tok_linear(q, C, C, (const __nv_bfloat16*)L.sa_wv_t, L.sa_bv, b2, C, C, false, false);
        tok_self_attn_core(b0, b1, b2, b3, sc);
        if (a.phase == 0) {
            tok_linear(b3, C, C, (const __nv_bfloat16*)L.sa_wo_t, L.sa_bo, q, C, C, false, false);  // replaces the queries
        } else {
            tok_linear(b3, C, C, (const __nv_bfloat16*)L.sa_wo_t, L.sa_bo, q, C, C, true, false);   // residual
        }
        tok_layernorm(q, L.n1_g, L.n1_b, 1e-5f);
        // ---- (2) tokens attend to the image
        tok_add(b3, q, qpe, NT * C);
        tok_linear(b3, C, C, (const __nv_bfloat16*)L.t2i_wq_t, L.t2i_bq, b0, CI, CI, false, false);
        tok_t2i_attention(b0, kv, a.ldkv, 0, CI, a.hw, b1, sc);
        tok_linear(b1, CI, CI, (const __nv_bfloat16*)L.t2i_wo_t, L.t2i_bo, q, C, C, true, false);
        tok_layernorm(q, L.n2_g, L.n2_b, 1e-5f);
        // ---- (3) MLP
        tok_linear(q, C, C, (const __nv_bfloat16*)L.mlp_w1_t, L.mlp_b1, big, 2048, 2048, false, true);
        tok_linear(big, 2048, 2048, (const __nv_bfloat16*)L.mlp_w2_t, L.mlp_b2, q, C, C, true, false);
        tok_layernorm(q, L.n3_g, L.n3_b, 1e-5f);
        // ---- (4) token-side K/V for the image->token attention
        tok_add(b3, q, qpe, NT * C);
        tok_linear(b3, C, C, (const __nv_bfloat16*)L.i2t_wk_t, L.i2t_bk, b0, CI, CI, false, false);
        tok_linear(q, C, C, (const __nv_bfloat16*)L.i2t_wv_t, L.i2t_bv, b1, CI, CI, false, false);
        for (int i = tid; i < NT * CI; i += TK_THREADS) {
            a.KT[(size_t)p * NT * CI + i] = b0[i];
            a.VT[(size_t)p * NT * CI + i] = b1[i];
        }
        for (int i = tid; i < NT * C; i += TK_THREADS) a.Tq[(size_t)p * NT * C + i] = q[i];
    } else {
        // ---- final token->image attention, LayerNorm, hypernetworks, IoU head
        tok_add(b3, q, qpe, NT * C);
        tok_linear(b3, C, C, (const __nv_bfloat16*)a.fin_wq_t, a.fin_bq, b0, CI, CI, false, false);
        tok_t2i_attention(b0, kv, a.ldkv, 0, CI, a.hw, b1, sc);
        tok_linear(b1, CI, CI, (const __nv_bfloat16*)a.fin_wo_t, a.fin_bo, q, C, C, true, false);
        tok_layernorm(q, a.nf_g, a.nf_b, 1e-5f);
        for (int i = tid; i < NT * C; i += TK_THREADS) a.Tq[(size_t)p * NT * C + i] = q[i];
        // hypernetwork MLP m acts on mask token (1+m).  tok_linear works on NT rows; we feed row-replicated inputs
        // so each call evaluates one MLP for its own token (rows other than 0 are ignored).
        for (int m = 0; m < a.n_mask_tokens; ++m) {
            for (int i = tid; i < NT * C; i += TK_THREADS) b3[i] = q[(1 + m) * C + (i % C)];
            __syncthreads();
            W = (const __nv_bfloat16*)a.hyp_w0_t + (size_t)m * C * C;
            tok_linear(b3, C, C, W, a.hyp_b0 + m * C, b0, C, C, false, true);
            W = (const __nv_bfloat16*)a.hyp_w1_t + (size_t)m * C * C;
            tok_linear(b0, C, C, W, a.hyp_b1 + m * C, b1, C, C, false, true);
            W = (const __nv_bfloat16*)a.hyp_w2_t + (size_t)m * C * 32;
            tok_linear(b1, C, C, W, a.hyp_b2 + m * 32, b2, 32, 32, false, false);
            for (int i = tid; i < 32; i += TK_THREADS) a.hyper[((size_t)p * a.n_mask_tokens + m) * 32 + i] = b2[i];
            __syncthreads();
        }
        for (int i = tid; i < NT * C; i += TK_THREADS) b3[i] = q[i % C];  // iou token = row 0
        __syncthreads();
        tok_linear(b3, C, C, (const __nv_bfloat16*)a.iou_w0_t, a.iou_b0, b0, C, C, false, true);
        tok_linear(b0, C, C, (const __nv_bfloat16*)a.iou_w1_t, a.iou_b1, b1, C, C, false, true);
        tok_linear(b1, C, C, (const __nv_bfloat16*)a.iou_w2_t, a.iou_b2, b2, 8, a.n_mask_tokens, false, false);
        for (int i = tid; i < a.n_mask_tokens; i += TK_THREADS) a.iou[(size_t)p * a.n_mask_tokens + i] = b2[i];
    }
}

// keys0[p, pos, :] = img_emb[img(p), pos, :] + no_mask_embed      (bf16, channels-last)
__global__ void __launch_bounds__(256) expand_keys_kernel(const __nv_bfloat16* __restrict__ emb, const int* __restrict__ prompt_img,
                                                          const float* __restrict__ no_mask, __nv_bfloat16* __restrict__ keys, int P, int hw) {
    const long long total = (long long)P * hw * (C / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % (C / 8));
        const long long row = i / (C / 8);
        const int p = (int)(row / hw), pos = (int)(row % hw);
        const int img = prompt_img[p];
        uint4 u = *reinterpret_cast<const uint4*>(emb + ((size_t)img * hw + pos) * C + c8 * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        float4 n0 = *reinterpret_cast<const float4*>(no_mask + c8 * 8);
        float4 n1 = *reinterpret_cast<const float4*>(no_mask + c8 * 8 + 4);
        float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]), c = __bfloat1622float2(h[2]), d = __bfloat1622float2(h[3]);
        uint4 o;
        o.x = pack_bf16x2(a.x + n0.x, a.y + n0.y);
        o.y = pack_bf16x2(b.x + n0.z, b.y + n0.w);
        o.z = pack_bf16x2(c.x + n1.x, c.y + n1.y);
        o.w = pack_bf16x2(d.x + n1.z, d.y + n1.w);
        *reinterpret_cast<uint4*>(keys + row * C + c8 * 8) = o;
    }
}

// image -> token attention, one thread per image position (6 keys, 8 heads x 16): out bf16 [P*hw][128]
__global__ void __launch_bounds__(256) i2t_attention_kernel(const __nv_bfloat16* __restrict__ kvq, int ldkv, int q_off, const float* __restrict__ KT,
                                                            const float* __restrict__ VT, __nv_bfloat16* __restrict__ out, int hw) {
    __shared__ float kt[NT * CI], vt[NT * CI];
    const int p = blockIdx.y;
    for (int i = threadIdx.x; i < NT * CI; i += 256) {
        kt[i] = KT[(size_t)p * NT * CI + i];
        vt[i] = VT[(size_t)p * NT * CI + i];
    }
    __syncthreads();
    const int pos = blockIdx.x * 256 + threadIdx.x;
    if (pos >= hw) return;
    const size_t row = (size_t)p * hw + pos;
    const uint4* qr = reinterpret_cast<const uint4*>(kvq + row * ldkv + q_off);
    uint4* orow = reinterpret_cast<uint4*>(out + row * CI);
#pragma unroll 1
    for (int h = 0; h < 8; ++h) {
        float qf[16];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint4 u = qr[h * 2 + c];
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 f = __bfloat1622float2(hh[e]);
                qf[c * 8 + e * 2] = f.x;
                qf[c * 8 + e * 2 + 1] = f.y;
            }
        }
        float s[NT], mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            float d = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) d = fmaf(qf[e], kt[t * CI + h * 16 + e], d);
            s[t] = d * 0.25f;
            mx = fmaxf(mx, s[t]);
        }
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            s[t] = __expf(s[t] - mx);
            sum += s[t];
        }
        const float inv = 1.0f / sum;
        float o[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) o[e] = 0.f;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const float pt = s[t] * inv;
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = fmaf(pt, vt[t * CI + h * 16 + e], o[e]);
        }
        uint4 u0, u1;
        u0.x = pack_bf16x2(o[0], o[1]); u0.y = pack_bf16x2(o[2], o[3]); u0.z = pack_bf16x2(o[4], o[5]); u0.w = pack_bf16x2(o[6], o[7]);
        u1.x = pack_bf16x2(o[8], o[9]); u1.y = pack_bf16x2(o[10], o[11]); u1.z = pack_bf16x2(o[12], o[13]); u1.w = pack_bf16x2(o[14], o[15]);
        orow[h * 2] = u0;
        orow[h * 2 + 1] = u1;
    }
}

// U fp32 [P*hw][4*32] (ConvTranspose2d output, sub-pixel major) -> LayerNorm2d(32) -> GELU -> dot with hypernetwork vectors
//   low_res[p, m, 2y+dy, 2x+dx] = sum_c hyper[p, mask_start+m, c] * gelu(LN(U[p, (y,x), (dy,dx), :]))_c
// Optionally also accumulates the depth-head pooling: pooled[p, 0..31] += sigmoid(logit_0) * up_c, pooled[p, 32] += sigmoid(logit_0)
__global__ void __launch_bounds__(256) upscale_mask_kernel(const float* __restrict__ U, const float* __restrict__ g, const float* __restrict__ b,
                                                           const float* __restrict__ hyper, int n_mask_tokens, int mask_start, int n_out,
                                                           float* __restrict__ low_res, float* __restrict__ pooled, int hw, int gw, long long rows) {
    const long long row = (long long)blockIdx.x * 64 + (threadIdx.x >> 2);
    const int sub = threadIdx.x & 3;
    __shared__ float pool_sm[33];
    const bool do_pool = pooled != nullptr;  // all rows of a block belong to one prompt when hw % 64 == 0
    if (do_pool) {
        if (threadIdx.x < 33) pool_sm[threadIdx.x] = 0.f;
        __syncthreads();
    }
    if (row < rows) {
        const int p = (int)(row / hw), pos = (int)(row % hw);
        const int y = pos / gw, x = pos % gw;
        float v[32];
        const float4* ur = reinterpret_cast<const float4*>(U + row * 128 + sub * 32);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 t = ur[i];
            v[i * 4] = t.x; v[i * 4 + 1] = t.y; v[i * 4 + 2] = t.z; v[i * 4 + 3] = t.w;
            s += (t.x + t.y) + (t.z + t.w);
        }
        const float mean = s * (1.0f / 32);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) q = fmaf(v[i] - mean, v[i] - mean, q);
        const float rstd = 1.0f / sqrtf(q * (1.0f / 32) + 1e-6f);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf((v[i] - mean) * rstd * g[i] + b[i]);
        const int dy = sub >> 1, dx = sub & 1;
        const int OW = 2 * gw;
        const long long opix = (long long)(2 * y + dy) * OW + (2 * x + dx);
        const long long plane = (long long)4 * hw;
        float logit0 = 0.f;
        for (int m = 0; m < n_out; ++m) {
            const float* hv = hyper + ((size_t)p * n_mask_tokens + mask_start + m) * 32;
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) d = fmaf(hv[i], v[i], d);
            low_res[((long long)p * n_out + m) * plane + opix] = d;
            if (m == 0) logit0 = d;
        }
        if (do_pool) {
            const float wgt = 1.0f / (1.0f + __expf(-logit0));
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(&pool_sm[i], wgt * v[i]);
            atomicAdd(&pool_sm[32], wgt);
        }
    }
    if (do_pool) {
        __syncthreads();
        const long long row0 = (long long)blockIdx.x * 64;
        if (threadIdx.x < 33 && row0 < rows) atomicAdd(&pooled[(row0 / hw) * 33 + threadIdx.x], pool_sm[threadIdx.x]);
    }
}

struct DecBuffers {
    __nv_bfloat16 *keysA, *keysB, *kvq, *a2;
    float *U, *Tq, *Tpe, *KT, *VT, *hyper, *iou_all;
};

bool carve(Workspace& ws, int P, int hw, DecBuffers& d) {
    bool ok = true;
    auto take = [&](size_t bytes) {
        void* p = ws.take(bytes);
        ok = ok && p != nullptr;
        return p;
    };
    const size_t rows = (size_t)P * hw;
    d.keysA = (__nv_bfloat16*)take(rows * C * 2);
    d.keysB = (__nv_bfloat16*)take(rows * C * 2);
    d.kvq = (__nv_bfloat16*)take(rows * 384 * 2);
    d.a2 = (__nv_bfloat16*)take(rows * CI * 2);
    d.U = (float*)take(rows * 128 * 4);
    d.Tq = (float*)take((size_t)P * NT * C * 4);
    d.Tpe = (float*)take((size_t)P * NT * C * 4);
    d.KT = (float*)take((size_t)P * NT * CI * 4);
    d.VT = (float*)take((size_t)P * NT * CI * 4);
    d.hyper = (float*)take((size_t)P * 4 * 32 * 4);
    d.iou_all = (float*)take((size_t)P * 4 * 4);
    return ok;
}

}  // namespace
}  // namespace wg

using namespace wg;

extern "C" size_t wg_mask_decoder_workspace_bytes(int P, int hw) {
    if (P <= 0 || hw <= 0) return 0;
    Workspace ws(nullptr, 0);
    DecBuffers d;
    carve(ws, P, hw, d);
    return ws.used();
}

extern "C" int wg_mask_decoder_forward(const wg_mask_decoder_weights* w, const void* img_emb_tokens_bf16, const float* txt_emb,
                                       const int32_t* prompt_img, int P, int multimask_output, float* low_res_out, float* iou_out,
                                       float* depth_pool_out, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(w != nullptr, "wg_mask_decoder_forward: null weights");
    if (P == 0) return WG_OK;
    WG_REQUIRE(img_emb_tokens_bf16 && txt_emb && prompt_img && low_res_out && iou_out && workspace, "wg_mask_decoder_forward: null pointer");
    WG_REQUIRE(P > 0, "wg_mask_decoder_forward: P=%d", P);
    WG_REQUIRE(w->n_mask_tokens == 4 && w->up_stages == 1, "wg_mask_decoder_forward: only the multi-scale decoder head (4 mask tokens, one ConvTranspose) is built");
    const int hw = w->grid_h * w->grid_w;
    WG_REQUIRE(hw > 0 && hw <= 4096, "wg_mask_decoder_forward: grid %dx%d unsupported", w->grid_h, w->grid_w);
    WG_REQUIRE(depth_pool_out == nullptr || hw % 64 == 0, "wg_mask_decoder_forward: depth pooling needs hw %% 64 == 0");
    if (!device_is_sm100()) {
        set_error("wg_mask_decoder_forward: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    Workspace ws(workspace, workspace_bytes);
    DecBuffers d;
    WG_REQUIRE(carve(ws, P, hw, d), "wg_mask_decoder_forward: workspace too small (%zu given, %zu needed)", workspace_bytes,
               wg_mask_decoder_workspace_bytes(P, hw));
    const long long rows = (long long)P * hw;
    WG_REQUIRE(rows < (1ll << 31), "wg_mask_decoder_forward: too many prompt tokens");
    const int n_out = multimask_output ? w->n_mask_tokens : 1;
    const int mask_start = 0;  // MaskDecoderMultiScale keeps index 0 in both modes (mask_decoder_multi_scale.py:126-132)

    long long blocks = (rows * (C / 8) + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    {
        Prof prof("dec_expand_keys", s, 0.0, (double)rows * C * 4.0);
    expand_keys_kernel<<<(unsigned)blocks, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(img_emb_tokens_bf16), prompt_img, w->no_mask, d.keysA, P, hw);
    }
    WG_CHECK_CUDA(cudaGetLastError());

    const size_t tk_smem = (size_t)(6 * NT * C + NT * 2048 + NT * hw) * sizeof(float);
    WG_CHECK_CUDA(cudaFuncSetAttribute(decoder_token_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    WG_REQUIRE(tk_smem <= 227 * 1024, "wg_mask_decoder_forward: token kernel shared memory %zu too large", tk_smem);

    TokArgs ta = {};
    ta.hw = hw;
    ta.n_mask_tokens = w->n_mask_tokens;
    ta.fin_wq_t = w->fin_wq_t; ta.fin_bq = w->fin_bq; ta.fin_wo_t = w->fin_wo_t; ta.fin_bo = w->fin_bo; ta.nf_g = w->nf_g; ta.nf_b = w->nf_b;
    ta.hyp_w0_t = w->hyp_w0_t; ta.hyp_w1_t = w->hyp_w1_t; ta.hyp_w2_t = w->hyp_w2_t;
    ta.hyp_b0 = w->hyp_b0; ta.hyp_b1 = w->hyp_b1; ta.hyp_b2 = w->hyp_b2;
    ta.iou_w0_t = w->iou_w0_t; ta.iou_w1_t = w->iou_w1_t; ta.iou_w2_t = w->iou_w2_t;
    ta.iou_b0 = w->iou_b0; ta.iou_b1 = w->iou_b1; ta.iou_b2 = w->iou_b2;
    ta.out_tokens = w->out_tokens; ta.sparse_add = w->sparse_add; ta.txt = txt_emb;
    ta.Tq = d.Tq; ta.Tpe = d.Tpe; ta.KT = d.KT; ta.VT = d.VT; ta.hyper = d.hyper; ta.iou = d.iou_all;

    // bring-up aid: WG_DEBUG_DECODER_STOP=<n> returns after the n-th launch group so intermediates can be inspected
    const char* dbg = getenv("WG_DEBUG_DECODER_STOP");
    const int stop_at = dbg ? atoi(dbg) : 0;
    int step = 0;
#define WG_DBG_STEP()                         \
    do {                                      \
        if (stop_at && ++step == stop_at) return WG_OK; \
    } while (0)

    __nv_bfloat16* keys = d.keysA;
    __nv_bfloat16* keys_next = d.keysB;
    for (int l = 0; l < 2; ++l) {
        const wg_twoway_layer& L = w->layers[l];
        {   // [K_t2i | V_t2i | Q_i2t] with the positional term as a per-position bias table
            wg_gemm_args a = {};
            a.A = keys; a.lda = C; a.W = L.w_img; a.ldw = C; a.M = (int)rows; a.N = 384; a.K = C;
            a.bias = L.b_img; a.bias_period = hw; a.out_mode = WG_OUT_BF16; a.out = d.kvq; a.ldo = 384;
            WG_TRY(wg_gemm(&a, s));
        }
        WG_DBG_STEP();
        ta.phase = l; ta.L = L; ta.kv = d.kvq; ta.ldkv = 384;
        {
            Prof prof("dec_token", s, (double)P * 2.0 * 10.2e6, (double)P * (hw * 512.0 + 3.0e6));
            decoder_token_kernel<<<P, TK_THREADS, tk_smem, s>>>(ta);
        }
        WG_CHECK_CUDA(cudaGetLastError());
        WG_DBG_STEP();
        {
            Prof prof("dec_i2t_attention", s, (double)rows * 8 * 6 * 16 * 4.0, (double)rows * 512.0);
            i2t_attention_kernel<<<dim3((hw + 255) / 256, P), 256, 0, s>>>(d.kvq, 384, 256, d.KT, d.VT, d.a2, hw);
        }
        WG_CHECK_CUDA(cudaGetLastError());
        WG_DBG_STEP();
        {   // keys' = LayerNorm4(keys + attn W_o^T + b)
            wg_gemm_args a = {};
            a.A = d.a2; a.lda = CI; a.W = L.i2t_wo; a.ldw = CI; a.M = (int)rows; a.N = C; a.K = CI;
            a.bias = L.i2t_bo; a.bias_period = 1; a.out_mode = WG_OUT_BF16_LN; a.out = keys_next; a.ldo = C; a.resid = keys;
            a.ln_gamma = L.n4_g; a.ln_beta = L.n4_b; a.ln_eps = 1e-5f;
            WG_TRY(wg_gemm(&a, s));
        }
        WG_DBG_STEP();
        __nv_bfloat16* t = keys; keys = keys_next; keys_next = t;
    }
    {   // final token->image attention: [K | V]
        wg_gemm_args a = {};
        a.A = keys; a.lda = C; a.W = w->w_img_fin; a.ldw = C; a.M = (int)rows; a.N = 256; a.K = C;
        a.bias = w->b_img_fin; a.bias_period = hw; a.out_mode = WG_OUT_BF16; a.out = d.kvq; a.ldo = 256;
        WG_TRY(wg_gemm(&a, s));
    }
    ta.phase = 2; ta.kv = d.kvq; ta.ldkv = 256;
    {
        Prof prof("dec_token", s, (double)P * 2.0 * 3.0e6, (double)P * (hw * 512.0 + 1.5e6));
        decoder_token_kernel<<<P, TK_THREADS, tk_smem, s>>>(ta);
    }
    WG_CHECK_CUDA(cudaGetLastError());
    // ConvTranspose2d(256 -> 32, k=2, s=2) as a GEMM over positions (N = 4 sub-pixels x 32 channels), fp32 out
    WG_TRY(gemm_f32_out(keys, C, w->w_up, (int)rows, 128, C, w->b_up, WG_ACT_NONE, d.U, 128, nullptr, s));
    if (depth_pool_out) WG_CHECK_CUDA(cudaMemsetAsync(depth_pool_out, 0, (size_t)P * 33 * sizeof(float), s));
    {
        Prof prof("dec_upscale_mask", s, (double)rows * 128 * 12.0, (double)rows * (512.0 + 16.0 * n_out));
    upscale_mask_kernel<<<(unsigned)((rows + 63) / 64), 256, 0, s>>>(d.U, w->up_ln_g, w->up_ln_b, d.hyper, w->n_mask_tokens, mask_start, n_out,
                                                                  low_res_out, depth_pool_out, hw, w->grid_w, rows);
    }
    WG_CHECK_CUDA(cudaGetLastError());
    // iou predictions for the selected masks
    WG_CHECK_CUDA(cudaMemcpy2DAsync(iou_out, n_out * sizeof(float), d.iou_all + mask_start, w->n_mask_tokens * sizeof(float), n_out * sizeof(float), P,
                                    cudaMemcpyDeviceToDevice, s));
    return WG_OK;
}
