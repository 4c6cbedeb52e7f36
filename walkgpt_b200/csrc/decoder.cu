// Numerics: this stage is ~1 % of the path's FLOPs but dominates the mask-logit error when run in plain bf16, so it runs at
// near-fp32 accuracy: image-side activations are kept as split-bf16 (hi | lo) and multiplied against [W_hi | W_hi | W_lo]
// on the tensor cores (3 bf16 products ~ 16-bit mantissas), projections consumed by CUDA cores are stored in fp32, and the
// token-side weights are fp32.
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
constexpr int TK_THREADS = 512;  // one CTA per prompt: the kernel streams fp32 weights from L2, more threads = more loads in flight
constexpr int TK_KEY_GROUPS = TK_THREADS / 16;  // token->image attention: 16 dims x this many key groups


// out[t][n] = act( base + sum_k in[t][k] * Wt[k][n] ), t < NT.   in/out in shared memory, Wt fp32 [K][N] in global.
// base = bias[n] (+ out[t][n] when `accumulate`).  Split-K over thread groups; the partial sums go through `part` and are added
// in a fixed order, so results are run-to-run deterministic.  The kernel is bound by the latency of the weight stream from L2
// (every CTA reads the layer's 6 MB): VEC = 4 issues 16-byte loads, four per thread in flight, and needs 2x the partial-sum
// buffer; it is used whenever that fits next to the score buffer (hw <= 2048), VEC = 2 otherwise.
template <int VEC>
__device__ void tok_linear_v(const float* in, int ldi, int K, const float* __restrict__ Wt, const float* __restrict__ bias, float* out,
                             int ldo, int N, bool accumulate, bool relu, float* part, int part_floats) {
    const int tid = threadIdx.x;
    const int groups = N / VEC;
    int ks = TK_THREADS / groups;
    if (ks < 1) ks = 1;
    if (ks > K) ks = K;
    if (ks > 1 && ks * NT * N > part_floats) ks = part_floats / (NT * N);
    const int klen = (K + ks - 1) / ks;
    for (int item = tid; item < groups * ks; item += TK_THREADS) {
        const int gr = item % groups, kpart = item / groups;
        const int k0 = kpart * klen;
        const int k1 = min(K, k0 + klen);
        float acc[VEC][NT];
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[v][t] = 0.f;
        const float* wp = Wt + (size_t)k0 * N + gr * VEC;
#pragma unroll 4
        for (int k = k0; k < k1; ++k, wp += N) {
            float w[VEC];
            if constexpr (VEC == 4) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(wp));
                w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
            } else {
                const float2 u = __ldg(reinterpret_cast<const float2*>(wp));
                w[0] = u.x; w[1] = u.y;
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float x = in[t * ldi + k];
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v][t] = fmaf(x, w[v], acc[v][t]);
            }
        }
        if (ks == 1) {  // this thread owns the whole dot product
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const int n = gr * VEC + v;
                    const float r = acc[v][t] + (bias ? bias[n] : 0.f) + (accumulate ? out[t * ldo + n] : 0.f);
                    out[t * ldo + n] = relu ? fmaxf(r, 0.f) : r;
                }
        } else {
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int v = 0; v < VEC; ++v) part[(kpart * NT + t) * N + gr * VEC + v] = acc[v][t];
        }
    }
    __syncthreads();
    if (ks > 1) {
        for (int i = tid; i < NT * N; i += TK_THREADS) {
            const int t = i / N, n = i - t * N;
            float v = (bias ? bias[n] : 0.f) + (accumulate ? out[t * ldo + n] : 0.f);
            for (int kp = 0; kp < ks; ++kp) v += part[(kp * NT + t) * N + n];
            out[t * ldo + n] = relu ? fmaxf(v, 0.f) : v;
        }
        __syncthreads();
    }
}
struct TokPart {
    float* buf;
    int floats;  // TK_THREADS * 2 * NT (VEC = 2) or twice that (VEC = 4)
};
__device__ __forceinline__ void tok_linear(const float* in, int ldi, int K, const float* __restrict__ Wt, const float* __restrict__ bias, float* out,
                                           int ldo, int N, bool accumulate, bool relu, TokPart part) {
    if (part.floats >= TK_THREADS * 4 * NT && (N & 3) == 0) tok_linear_v<4>(in, ldi, K, Wt, bias, out, ldo, N, accumulate, relu, part.buf, part.floats);
    else tok_linear_v<2>(in, ldi, K, Wt, bias, out, ldo, N, accumulate, relu, part.buf, part.floats);
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

// token -> image attention: Qt [NT][128] (smem), K/V rows in global (fp32, row pitch ldkv), 8 heads x 16.  A [NT][128] out (smem)
__device__ void tok_t2i_attention(const float* Qt, const float* __restrict__ kv, int ldkv, int k_off, int v_off, int hw, float* A,
                                  float* sc /* [NT][hw] */, float* part) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int h = 0; h < 8; ++h) {
        float qh[NT][16];
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int e = 0; e < 16; ++e) qh[t][e] = Qt[t * CI + h * 16 + e];
        for (int key = tid; key < hw; key += TK_THREADS) {
            const float4* kr = reinterpret_cast<const float4*>(kv + (size_t)key * ldkv + k_off + h * 16);
            float kf[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4 u = __ldg(kr + c);
                kf[c * 4] = u.x; kf[c * 4 + 1] = u.y; kf[c * 4 + 2] = u.z; kf[c * 4 + 3] = u.w;
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
            const int d = tid & 15, grp = tid >> 4;  // 16 dims x TK_KEY_GROUPS key groups
            float o[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) o[t] = 0.f;
            for (int key = grp; key < hw; key += TK_KEY_GROUPS) {
                const float v = __ldg(kv + (size_t)key * ldkv + v_off + h * 16 + d);
#pragma unroll
                for (int t = 0; t < NT; ++t) o[t] = fmaf(sc[t * hw + key], v, o[t]);
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) part[(grp * NT + t) * 16 + d] = o[t];
        }
        __syncthreads();
        if (tid < NT * 16) {  // fixed-order reduction over the 16 key groups
            const int t = tid >> 4, d = tid & 15;
            float v = 0.f;
            for (int gI = 0; gI < TK_KEY_GROUPS; ++gI) v += part[(gI * NT + t) * 16 + d];
            A[t * CI + h * 16 + d] = v;
        }
        __syncthreads();
    }
}

// token -> image attention as its own kernel, split over KEYS: one CTA per (128-key chunk, prompt) handles all 8 heads, so every
// global read is a whole 512-byte K or V row (the eight serial heads inside the one-CTA-per-prompt token kernel took 227 us per layer
// at 192 prompts; a (head, prompt) grid read 64-byte slices at a 1.5 KB stride and reached 1.3 TB/s).  Each chunk emits flash-style
// partials -- per (token, head) the chunk maximum m, the sum l of exp(s - m) and the un-normalised o[16] -- which the token kernel's
// next part combines in a fixed chunk order (deterministic).  Qt [P][NT][128]; K / V = columns [k_off, +128) / [v_off, +128) of the
// fp32 image-side projections kv [P*hw][ldkv].
constexpr int T2I_THREADS = 256;
constexpr int T2I_CHUNK = 128;                 // keys per CTA
constexpr int T2I_ROWS = NT * 8;               // (token, head) score rows
constexpr int T2I_LD = T2I_CHUNK + 4;          // padded row, still 16-byte aligned: the P V pass reads four keys per LDS.128
__global__ void __launch_bounds__(T2I_THREADS) t2i_partial_kernel(const float* __restrict__ Qt, const float* __restrict__ kvall, int ldkv, int k_off,
                                                                   int v_off, int hw, float* __restrict__ Pm, float* __restrict__ Pl,
                                                                   float* __restrict__ Po, const int* __restrict__ kv_img) {
    __shared__ __align__(16) float qs[NT * CI];
    __shared__ __align__(16) float sc[T2I_ROWS * T2I_LD];
    const int chunk = blockIdx.x, nchunk = gridDim.x, p = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = chunk * T2I_CHUNK;
    // kv_img != null (layer 0): the projections exist once per image, kv rows of prompt p are those of image kv_img[p]
    const float* kv = kvall + ((size_t)(kv_img != nullptr ? kv_img[p] : p) * hw + k0) * ldkv;
    for (int i = tid; i < NT * CI; i += T2I_THREADS) qs[i] = Qt[(size_t)p * NT * CI + i];
    __syncthreads();
    {   // scores: thread = (key, group of 4 heads): 256 contiguous bytes of the key's K row
        const int key = tid & (T2I_CHUNK - 1), hg = tid >> 7;
        const bool valid = k0 + key < hw;
        float kf[64];
        if (valid) {
            const float4* kr = reinterpret_cast<const float4*>(kv + (size_t)key * ldkv + k_off + hg * 64);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float4 u = __ldg(kr + c);
                kf[c * 4] = u.x; kf[c * 4 + 1] = u.y; kf[c * 4 + 2] = u.z; kf[c * 4 + 3] = u.w;
            }
        }
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
            const int h = hg * 4 + hh;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                float d = 0.f;
#pragma unroll
                for (int e = 0; e < 16; e += 4) {  // the query slice as four broadcast LDS.128 (scalar reads made this loop LDS-bound)
                    const float4 qv = *reinterpret_cast<const float4*>(&qs[t * CI + h * 16 + e]);
                    d = fmaf(qv.x, kf[hh * 16 + e], d);
                    d = fmaf(qv.y, kf[hh * 16 + e + 1], d);
                    d = fmaf(qv.z, kf[hh * 16 + e + 2], d);
                    d = fmaf(qv.w, kf[hh * 16 + e + 3], d);
                }
                sc[(t * 8 + h) * T2I_LD + key] = valid ? d * 0.25f : -INFINITY;  // 1/sqrt(16)
            }
        }
    }
    __syncthreads();
    for (int row = warp; row < T2I_ROWS; row += T2I_THREADS / 32) {  // chunk-local softmax numerators
        float* r = sc + row * T2I_LD;
        float v[4];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i] = r[lane + 32 * i];
            mx = fmaxf(mx, v[i]);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float e = __expf(v[i] - mx);  // -inf (keys past hw) -> 0; a chunk always holds at least one real key
            r[lane + 32 * i] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        if (lane == 0) {
            Pm[((size_t)p * nchunk + chunk) * T2I_ROWS + row] = mx;
            Pl[((size_t)p * nchunk + chunk) * T2I_ROWS + row] = sum;
        }
    }
    __syncthreads();
    {   // o[t][c] = sum_key p[t][head(c)][key] * V[key][c].  A warp owns every 8th key and reads its whole 512-byte V row with one
        // LDG.128 per lane (four columns each); four rows in flight per warp keep enough bytes outstanding (one float per lane did not)
        const int c4 = lane * 4, h = lane >> 2;
        float o[NT][4];
#pragma unroll
        for (int t = 0; t < NT; ++t) o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f;
        const int kend = min(T2I_CHUNK, hw - k0);
        const float* vp = kv + v_off + c4;
#pragma unroll 4
        for (int key = warp; key < T2I_CHUNK; key += T2I_THREADS / 32) {
            if (key < kend) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(vp + (size_t)key * ldkv));
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const float pt = sc[(t * 8 + h) * T2I_LD + key];
                    o[t][0] = fmaf(pt, v.x, o[t][0]);
                    o[t][1] = fmaf(pt, v.y, o[t][1]);
                    o[t][2] = fmaf(pt, v.z, o[t][2]);
                    o[t][3] = fmaf(pt, v.w, o[t][3]);
                }
            }
        }
        __syncthreads();  // every warp is done with the probabilities: their storage now takes the per-warp partial sums
        float* part = sc;  // [warps][NT * CI]  (8 * 768 <= 48 * 132 floats)
#pragma unroll
        for (int t = 0; t < NT; ++t) *reinterpret_cast<float4*>(&part[warp * (NT * CI) + t * CI + c4]) = make_float4(o[t][0], o[t][1], o[t][2], o[t][3]);
        __syncthreads();
        for (int i = tid; i < NT * CI; i += T2I_THREADS) {  // fixed-order sum over the warps
            float v = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < T2I_THREADS / 32; ++w8) v += part[w8 * (NT * CI) + i];
            Po[((size_t)p * nchunk + chunk) * NT * CI + i] = v;
        }
    }
}

// combine the chunk partials of t2i_partial_kernel for prompt p into A [NT][128] (shared memory), fixed chunk order
__device__ void tok_t2i_combine(const float* __restrict__ Pm, const float* __restrict__ Pl, const float* __restrict__ Po, int p, int nchunk, float* A) {
    for (int i = threadIdx.x; i < NT * CI; i += TK_THREADS) {
        const int t = i / CI, c = i - t * CI, row = t * 8 + (c >> 4);
        const float* pm = Pm + (size_t)p * nchunk * T2I_ROWS + row;
        const float* pl = Pl + (size_t)p * nchunk * T2I_ROWS + row;
        float M = -INFINITY;
        for (int ch = 0; ch < nchunk; ++ch) M = fmaxf(M, pm[ch * T2I_ROWS]);
        float num = 0.f, den = 0.f;
        for (int ch = 0; ch < nchunk; ++ch) {
            const float w = __expf(pm[ch * T2I_ROWS] - M);
            num = fmaf(Po[(((size_t)p * nchunk + ch) * NT + t) * CI + c], w, num);
            den = fmaf(pl[ch * T2I_ROWS], w, den);
        }
        A[i] = num / den;
    }
    __syncthreads();
}

struct TokArgs {
    int phase, hw, n_mask_tokens;
    int part;                            // phases 0/1: 0 = whole layer, 1 = up to norm2 (then the MLP runs as GEMMs), 2 = from norm3 on;
                                         // 3 = up to the token->image query projection (Qt; the attention then runs as its own kernel over
                                         // (head, prompt)), 4 = from the attention output At on (out-proj, norm2, hand-over to the MLP GEMMs);
                                         // final phase: 3 = query projection only, 4 = from At on
    int part_floats;                     // split-K buffer of tok_linear: TK_THREADS * NT * 2 or * 4 floats
    wg_twoway_layer L;                   // weights of this layer (phases 0/1)
    // final-phase weights
    const void* fin_wq_t; const float* fin_bq; const void* fin_wo_t; const float* fin_bo; const float* nf_g; const float* nf_b;
    const void *hyp_w0_t, *hyp_w1_t, *hyp_w2_t; const float *hyp_b0, *hyp_b1, *hyp_b2;
    const void *iou_w0_t, *iou_w1_t, *iou_w2_t; const float *iou_b0, *iou_b1, *iou_b2;
    const float* out_tokens;   // [1+n_mask][256]
    const float* sparse_add;   // [256] or null
    const float* txt;          // [P][256]
    const float* kv;           // image-side projections of this phase, fp32 [P*hw][ldkv]
    int ldkv;
    float* Tq;                 // [P][NT][256] running queries
    __nv_bfloat16* Xs;         // part 1: the queries after norm2 once more as split-bf16 [P*NT][512] (A operand of the MLP GEMMs)
    float* Tpe;                // [P][NT][256] token positional term (= initial tokens)
    float* KT; float* VT;      // [P][NT][128] token-side K/V for image->token attention
    float* Qt;                 // [P][NT][128] token->image queries (part 3)
    const float *Pm, *Pl, *Po; int nchunk;  // chunk partials of t2i_partial_kernel (part 4 combines them)
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
    const TokPart part = {big + NT * 2048, a.part_floats};  // split-K partial sums
    float* sc = part.buf + a.part_floats;                   // [NT][max(hw, 48)]
    const int p = blockIdx.x;
    const int tid = threadIdx.x;
    const float* kv = a.kv + (size_t)p * a.hw * a.ldkv;
    const float* W;

    if (a.phase == 0 && a.part != 2 && a.part != 4) {
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
        if (a.part != 2) {
        if (a.part != 4) {
        // ---- (1) self attention on the tokens
        if (a.phase == 0) {
            for (int i = tid; i < NT * C; i += TK_THREADS) b3[i] = q[i];
            __syncthreads();
        } else {
            tok_add(b3, q, qpe, NT * C);
        }
        tok_linear(b3, C, C, (const float*)L.sa_wq_t, L.sa_bq, b0, C, C, false, false, part);
        tok_linear(b3, C, C, (const float*)L.sa_wk_t, L.sa_bk, b1, C, C, false, false, part);
        tok_linear(q, C, C, (const float*)L.sa_wv_t, L.sa_bv, b2, C, C, false, false, part);
        tok_self_attn_core(b0, b1, b2, b3, sc);
        if (a.phase == 0) {
            tok_linear(b3, C, C, (const float*)L.sa_wo_t, L.sa_bo, q, C, C, false, false, part);  // replaces the queries
        } else {
            tok_linear(b3, C, C, (const float*)L.sa_wo_t, L.sa_bo, q, C, C, true, false, part);   // residual
        }
        tok_layernorm(q, L.n1_g, L.n1_b, 1e-5f);
        // ---- (2) tokens attend to the image
        tok_add(b3, q, qpe, NT * C);
        tok_linear(b3, C, C, (const float*)L.t2i_wq_t, L.t2i_bq, b0, CI, CI, false, false, part);
        if (a.part == 3) {  // hand the queries over: the attention runs as t2i_attention_kernel over (head, prompt)
            for (int i = tid; i < NT * CI; i += TK_THREADS) a.Qt[(size_t)p * NT * CI + i] = b0[i];
            for (int i = tid; i < NT * C; i += TK_THREADS) a.Tq[(size_t)p * NT * C + i] = q[i];
            return;
        }
        tok_t2i_attention(b0, kv, a.ldkv, 0, CI, a.hw, b1, sc, part.buf);
        } else {  // part 4: resume with the attention output (combined from the key-chunk partials)
            tok_t2i_combine(a.Pm, a.Pl, a.Po, p, a.nchunk, b1);
        }
        tok_linear(b1, CI, CI, (const float*)L.t2i_wo_t, L.t2i_bo, q, C, C, true, false, part);
        tok_layernorm(q, L.n2_g, L.n2_b, 1e-5f);
        if (a.part == 1 || a.part == 4) {
            // the MLP runs as two GEMMs over all prompts' token rows: hand the queries over in fp32 (residual) and split-bf16 (operand)
            for (int i = tid; i < NT * C; i += TK_THREADS) {
                const float v = q[i];
                const __nv_bfloat16 hi = __float2bfloat16(v);
                a.Tq[(size_t)p * NT * C + i] = v;
                __nv_bfloat16* xr = a.Xs + ((size_t)p * NT + i / C) * (2 * C) + (i % C);
                xr[0] = hi;
                xr[C] = __float2bfloat16(v - __bfloat162float(hi));
            }
            return;
        }
        // ---- (3) MLP
        tok_linear(q, C, C, (const float*)L.mlp_w1_t, L.mlp_b1, big, 2048, 2048, false, true, part);
        tok_linear(big, 2048, 2048, (const float*)L.mlp_w2_t, L.mlp_b2, q, C, C, true, false, part);
        }
        tok_layernorm(q, L.n3_g, L.n3_b, 1e-5f);
        // ---- (4) token-side K/V for the image->token attention
        tok_add(b3, q, qpe, NT * C);
        tok_linear(b3, C, C, (const float*)L.i2t_wk_t, L.i2t_bk, b0, CI, CI, false, false, part);
        tok_linear(q, C, C, (const float*)L.i2t_wv_t, L.i2t_bv, b1, CI, CI, false, false, part);
        for (int i = tid; i < NT * CI; i += TK_THREADS) {
            a.KT[(size_t)p * NT * CI + i] = b0[i];
            a.VT[(size_t)p * NT * CI + i] = b1[i];
        }
        for (int i = tid; i < NT * C; i += TK_THREADS) a.Tq[(size_t)p * NT * C + i] = q[i];
    } else {
        // ---- final token->image attention, LayerNorm, hypernetworks, IoU head
        if (a.part != 4) {
            tok_add(b3, q, qpe, NT * C);
            tok_linear(b3, C, C, (const float*)a.fin_wq_t, a.fin_bq, b0, CI, CI, false, false, part);
            if (a.part == 3) {
                for (int i = tid; i < NT * CI; i += TK_THREADS) a.Qt[(size_t)p * NT * CI + i] = b0[i];
                return;
            }
            tok_t2i_attention(b0, kv, a.ldkv, 0, CI, a.hw, b1, sc, part.buf);
        } else {
            tok_t2i_combine(a.Pm, a.Pl, a.Po, p, a.nchunk, b1);
        }
        tok_linear(b1, CI, CI, (const float*)a.fin_wo_t, a.fin_bo, q, C, C, true, false, part);
        tok_layernorm(q, a.nf_g, a.nf_b, 1e-5f);
        for (int i = tid; i < NT * C; i += TK_THREADS) a.Tq[(size_t)p * NT * C + i] = q[i];
        // hypernetwork MLP m acts on mask token (1+m).  tok_linear works on NT rows; we feed row-replicated inputs
        // so each call evaluates one MLP for its own token (rows other than 0 are ignored).
        for (int m = 0; m < a.n_mask_tokens; ++m) {
            for (int i = tid; i < NT * C; i += TK_THREADS) b3[i] = q[(1 + m) * C + (i % C)];
            __syncthreads();
            W = (const float*)a.hyp_w0_t + (size_t)m * C * C;
            tok_linear(b3, C, C, W, a.hyp_b0 + m * C, b0, C, C, false, true, part);
            W = (const float*)a.hyp_w1_t + (size_t)m * C * C;
            tok_linear(b0, C, C, W, a.hyp_b1 + m * C, b1, C, C, false, true, part);
            W = (const float*)a.hyp_w2_t + (size_t)m * C * 32;
            tok_linear(b1, C, C, W, a.hyp_b2 + m * 32, b2, 32, 32, false, false, part);
            for (int i = tid; i < 32; i += TK_THREADS) a.hyper[((size_t)p * a.n_mask_tokens + m) * 32 + i] = b2[i];
            __syncthreads();
        }
        for (int i = tid; i < NT * C; i += TK_THREADS) b3[i] = q[i % C];  // iou token = row 0
        __syncthreads();
        tok_linear(b3, C, C, (const float*)a.iou_w0_t, a.iou_b0, b0, C, C, false, true, part);
        tok_linear(b0, C, C, (const float*)a.iou_w1_t, a.iou_b1, b1, C, C, false, true, part);
        tok_linear(b1, C, C, (const float*)a.iou_w2_t, a.iou_b2, b2, 8, a.n_mask_tokens, false, false, part);
        for (int i = tid; i < a.n_mask_tokens; i += TK_THREADS) a.iou[(size_t)p * a.n_mask_tokens + i] = b2[i];
    }
}

__device__ __forceinline__ void split_store8(__nv_bfloat16* hi_ptr, __nv_bfloat16* lo_ptr, const float (&v)[8]) {
    uint4 h, l;
    uint32_t* hp = reinterpret_cast<uint32_t*>(&h);
    uint32_t* lp = reinterpret_cast<uint32_t*>(&l);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat16 h0 = __float2bfloat16(v[2 * e]), h1 = __float2bfloat16(v[2 * e + 1]);
        hp[e] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
        lp[e] = pack_bf16x2(v[2 * e] - __bfloat162float(h0), v[2 * e + 1] - __bfloat162float(h1));
    }
    *reinterpret_cast<uint4*>(hi_ptr) = h;
    *reinterpret_cast<uint4*>(lo_ptr) = l;
}
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float2 t = __bfloat1622float2(h[e]);
        v[2 * e] = t.x;
        v[2 * e + 1] = t.y;
    }
}

// keys0[p, pos, :] = gate[p, pos] * img_emb[img(p), pos, :] + no_mask_embed      (split-bf16 in [B*hw, 512], split-bf16 out [P*hw, 512])
// gate = 1, or sigmoid(mean_m previous_masks[p, m, pos]) + 1 at level > 0 (mask_decoder_multi_scale.py:168-169)
__global__ void __launch_bounds__(256) expand_keys_kernel(const __nv_bfloat16* __restrict__ emb, const int* __restrict__ prompt_img,
                                                          const float* __restrict__ no_mask, __nv_bfloat16* __restrict__ keys, int P, int hw,
                                                          const float* __restrict__ prev_masks, int n_prev) {
    const long long total = (long long)P * hw * (C / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % (C / 8));
        const long long row = i / (C / 8);
        const int p = (int)(row / hw), pos = (int)(row % hw);
        const int img = prompt_img != nullptr ? prompt_img[p] : p;  // null: one key set per IMAGE (layer-0 projections shared by its prompts)
        const __nv_bfloat16* src = emb + ((size_t)img * hw + pos) * (2 * C) + c8 * 8;
        float gate = 1.0f;
        if (prev_masks != nullptr) {
            float m = 0.f;
            for (int k = 0; k < n_prev; ++k) m += prev_masks[((size_t)p * n_prev + k) * hw + pos];
            m /= (float)n_prev;
            gate = 1.0f / (1.0f + expf(-m)) + 1.0f;
        }
        float hi[8], lo[8], v[8];
        load8_bf16(src, hi);
        load8_bf16(src + C, lo);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = gate * (hi[e] + lo[e]) + no_mask[c8 * 8 + e];
        split_store8(keys + row * (2 * C) + c8 * 8, keys + row * (2 * C) + C + c8 * 8, v);
    }
}

// level > 0 of the multi-scale decoder: upsample_2x = ConvTranspose2d(256 -> 256, k2, s2) (a GEMM) -> LayerNorm2d(256, eps 1e-6)
// -> GELU on the image embedding.  U fp32 [B*hw, 4 sub-pixels x 256] -> split-bf16 [B, 4*hw raster positions of the 2h x 2w grid, 512].
// One warp per (position, sub-pixel); lane owns 8 channels.
__global__ void __launch_bounds__(256) up2x_ln_gelu_kernel(const float* __restrict__ U, const float* __restrict__ g, const float* __restrict__ b,
                                                           __nv_bfloat16* __restrict__ out, long long rows, int hw, int gw) {
    const long long item = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);  // (row, sub)
    const int lane = threadIdx.x & 31;
    if (item >= rows * 4) return;
    const long long row = item >> 2;
    const int sub = (int)(item & 3);
    const float4* src = reinterpret_cast<const float4*>(U + row * 1024 + sub * 256 + lane * 8);
    const float4 a0 = src[0], a1 = src[1];
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float s1 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s1 += v[e];
    const float mean = warp_sum(s1) * (1.0f / 256);
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) q = fmaf(v[e] - mean, v[e] - mean, q);
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / 256) + 1e-6f);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = gelu_erf((v[e] - mean) * rstd * g[lane * 8 + e] + b[lane * 8 + e]);
    const long long img = row / hw;
    const int pos = (int)(row % hw), y = pos / gw, x = pos % gw;
    const long long opos = (long long)(2 * y + (sub >> 1)) * (2 * gw) + (2 * x + (sub & 1));
    __nv_bfloat16* dst = out + (img * 4 * hw + opos) * (2 * C) + lane * 8;
    split_store8(dst, dst + C, o);
}

// image -> token attention, one thread per image position (6 keys, 8 heads x 16): q fp32 in, out split-bf16 [P*hw][2*128]
__global__ void __launch_bounds__(256) i2t_attention_kernel(const float* __restrict__ kvq, int ldkv, int q_off, const float* __restrict__ KT,
                                                            const float* __restrict__ VT, __nv_bfloat16* __restrict__ out, int hw,
                                                            const int* __restrict__ kv_img) {
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
    const size_t qrow = (size_t)(kv_img != nullptr ? kv_img[p] : p) * hw + pos;  // layer 0: image-side queries shared by the image's prompts
    const float4* qr = reinterpret_cast<const float4*>(kvq + qrow * ldkv + q_off);
    __nv_bfloat16* orow = out + row * (2 * CI);
#pragma unroll 1
    for (int h = 0; h < 8; ++h) {
        float qf[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float4 u = __ldg(qr + h * 4 + c);
            qf[c * 4] = u.x; qf[c * 4 + 1] = u.y; qf[c * 4 + 2] = u.z; qf[c * 4 + 3] = u.w;
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
        float o0[8], o1[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            o0[e] = o[e];
            o1[e] = o[8 + e];
        }
        split_store8(orow + h * 16, orow + CI + h * 16, o0);
        split_store8(orow + h * 16 + 8, orow + CI + h * 16 + 8, o1);
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
    float v[32];
    float pool_w = 0.f;
    if (row < rows) {
        const int p = (int)(row / hw), pos = (int)(row % hw);
        const int y = pos / gw, x = pos % gw;
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
        if (do_pool) pool_w = 1.0f / (1.0f + __expf(-logit0));
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= pool_w;
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
    if (do_pool) {
        // sigmoid-weighted channel sums for the depth extension: butterfly transpose-reduction across the warp (31 shuffles:
        // lane l ends up with the warp's sum of channel l), then one shared atomic per lane.  (33 shared atomics per thread
        // onto 33 addresses serialised 256-fold and made this kernel 0.8 ms.)
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
                const float send = upper ? v[i] : v[i + off];
                const float keep = upper ? v[i + off] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        atomicAdd(&pool_sm[lane], v[0]);
        const float wsum = warp_sum(pool_w);
        if (lane == 0) atomicAdd(&pool_sm[32], wsum);
    }
    if (do_pool) {
        __syncthreads();
        const long long row0 = (long long)blockIdx.x * 64;
        if (threadIdx.x < 33 && row0 < rows) atomicAdd(&pooled[(row0 / hw) * 33 + threadIdx.x], pool_sm[threadIdx.x]);
    }
}

// ---- SAM head (Path B), stage 1: U1 fp32 [rows, 4 sub-pixels x 64] -> LayerNorm2d(64, eps 1e-6) -> GELU -> split-bf16
// [rows * 4, hi 64 | lo 64]  (row index = position * 4 + sub-pixel)
__global__ void __launch_bounds__(256) up1_ln_gelu_kernel(const float* __restrict__ U1, const float* __restrict__ g, const float* __restrict__ b,
                                                          __nv_bfloat16* __restrict__ A2, long long rows4) {
    const long long r4 = (long long)blockIdx.x * 256 + threadIdx.x;
    if (r4 >= rows4) return;
    const float4* src = reinterpret_cast<const float4*>(U1 + r4 * 64);
    float v[64];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float4 t = src[i];
        v[i * 4] = t.x; v[i * 4 + 1] = t.y; v[i * 4 + 2] = t.z; v[i * 4 + 3] = t.w;
        s += (t.x + t.y) + (t.z + t.w);
    }
    const float mean = s * (1.0f / 64);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) q = fmaf(v[i] - mean, v[i] - mean, q);
    const float rstd = 1.0f / sqrtf(q * (1.0f / 64) + 1e-6f);
    __nv_bfloat16* dst = A2 + r4 * 128;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = gelu_erf((v[c8 * 8 + e] - mean) * rstd * g[c8 * 8 + e] + b[c8 * 8 + e]);
        split_store8(dst + c8 * 8, dst + 64 + c8 * 8, o);
    }
}
// ---- SAM head, stage 2: U2 fp32 [rows4, 4 sub-pixels x 32] -> GELU -> hypernetwork dot -> masks [P, n_out, 4h, 4w]
__global__ void __launch_bounds__(256) upscale_mask2_kernel(const float* __restrict__ U2, const float* __restrict__ hyper, int n_mask_tokens,
                                                            int mask_start, int n_out, float* __restrict__ low_res, int hw, int gw, long long rows4) {
    const long long r4 = (long long)blockIdx.x * 64 + (threadIdx.x >> 2);
    const int sub2 = threadIdx.x & 3;
    if (r4 >= rows4) return;
    const long long row = r4 >> 2;
    const int sub1 = (int)(r4 & 3);
    const int p = (int)(row / hw), pos = (int)(row % hw);
    const int y = pos / gw, x = pos % gw;
    float v[32];
    const float4* ur = reinterpret_cast<const float4*>(U2 + r4 * 128 + sub2 * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = ur[i];
        v[i * 4] = gelu_erf(t.x); v[i * 4 + 1] = gelu_erf(t.y); v[i * 4 + 2] = gelu_erf(t.z); v[i * 4 + 3] = gelu_erf(t.w);
    }
    const int OW = 4 * gw;
    const long long oy = 4 * y + 2 * (sub1 >> 1) + (sub2 >> 1), ox = 4 * x + 2 * (sub1 & 1) + (sub2 & 1);
    const long long plane = (long long)16 * hw;
    for (int m = 0; m < n_out; ++m) {
        const float* hv = hyper + ((size_t)p * n_mask_tokens + mask_start + m) * 32;
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) d = fmaf(hv[i], v[i], d);
        low_res[((long long)p * n_out + m) * plane + oy * OW + ox] = d;
    }
}

struct DecBuffers {
    __nv_bfloat16 *keysA, *keysB, *a2;   // split-bf16: [rows, 512], [rows, 512], [rows, 256]
    float *kvq, *U, *Tq, *Tpe, *KT, *VT, *Qt, *Pm, *Pl, *Po, *hyper, *iou_all;
    __nv_bfloat16 *Xs, *Hs;  // token rows of all prompts as split-bf16: MLP input [P*NT, 512], hidden [P*NT, 4096]
};

bool carve(Workspace& ws, int P, int hw, int up_stages, DecBuffers& d) {
    bool ok = true;
    auto take = [&](size_t bytes) {
        void* p = ws.take(bytes);
        ok = ok && p != nullptr;
        return p;
    };
    const size_t rows = (size_t)P * hw;
    d.keysA = (__nv_bfloat16*)take(rows * 2 * C * 2);
    d.keysB = (__nv_bfloat16*)take(rows * 2 * C * 2);
    d.kvq = (float*)take(rows * 384 * 4);
    d.a2 = (__nv_bfloat16*)take(rows * 2 * CI * 2);
    d.U = (float*)take(rows * 128 * 4 * (up_stages == 2 ? 4 : 1));  // up_stages 2: second ConvT output [rows*4, 128]
    d.Tq = (float*)take((size_t)P * NT * C * 4);
    d.Tpe = (float*)take((size_t)P * NT * C * 4);
    d.KT = (float*)take((size_t)P * NT * CI * 4);
    d.VT = (float*)take((size_t)P * NT * CI * 4);
    d.Qt = (float*)take((size_t)P * NT * CI * 4);
    const size_t nchunk = ((size_t)hw + T2I_CHUNK - 1) / T2I_CHUNK;
    d.Pm = (float*)take((size_t)P * nchunk * T2I_ROWS * 4);
    d.Pl = (float*)take((size_t)P * nchunk * T2I_ROWS * 4);
    d.Po = (float*)take((size_t)P * nchunk * NT * CI * 4);
    d.hyper = (float*)take((size_t)P * 4 * 32 * 4);
    d.iou_all = (float*)take((size_t)P * 4 * 4);
    d.Xs = (__nv_bfloat16*)take((size_t)P * NT * 2 * C * 2);
    d.Hs = (__nv_bfloat16*)take((size_t)P * NT * 2 * 2048 * 2);
    return ok;
}

}  // namespace
}  // namespace wg

using namespace wg;

extern "C" size_t wg_mask_decoder_workspace_bytes_ex(int P, int hw, int up_stages) {
    if (P <= 0 || hw <= 0) return 0;
    Workspace ws(nullptr, 0);
    DecBuffers d;
    carve(ws, P, hw, up_stages, d);
    return ws.used();
}
extern "C" size_t wg_mask_decoder_workspace_bytes(int P, int hw) { return wg_mask_decoder_workspace_bytes_ex(P, hw, 1); }

extern "C" int wg_mask_decoder_forward(const wg_mask_decoder_weights* w, const void* img_emb_tokens_bf16, const float* txt_emb,
                                       const int32_t* prompt_img, int P, int multimask_output, float* low_res_out, float* iou_out,
                                       float* depth_pool_out, void* workspace, size_t workspace_bytes, void* stream_) {
    return wg_mask_decoder_forward_level(w, img_emb_tokens_bf16, txt_emb, prompt_img, P, multimask_output, nullptr, 0, low_res_out, iou_out,
                                         depth_pool_out, workspace, workspace_bytes, stream_);
}

extern "C" size_t wg_upsample2x_workspace_bytes(int B, int hw) { return (B > 0 && hw > 0) ? (size_t)B * hw * 1024 * sizeof(float) : 0; }

extern "C" int wg_upsample2x_embedding(const void* w_up, int split_terms, const float* b_up, const float* ln_g, const float* ln_b,
                                       const void* emb_tokens_bf16, int B, int grid_h, int grid_w, void* out_tokens_bf16, void* workspace,
                                       size_t workspace_bytes, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    if (B == 0) return WG_OK;
    WG_REQUIRE(w_up && b_up && ln_g && ln_b && emb_tokens_bf16 && out_tokens_bf16 && workspace, "wg_upsample2x_embedding: null pointer");
    WG_REQUIRE(B > 0 && grid_h > 0 && grid_w > 0 && (split_terms == 2 || split_terms == 3), "wg_upsample2x_embedding: bad arguments");
    const int hw = grid_h * grid_w;
    const long long rows = (long long)B * hw;
    WG_REQUIRE(rows * 4 < (1ll << 31), "wg_upsample2x_embedding: too many positions");
    WG_REQUIRE(workspace_bytes >= wg_upsample2x_workspace_bytes(B, hw), "wg_upsample2x_embedding: workspace too small");
    if (!device_is_sm100()) {
        set_error("wg_upsample2x_embedding: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    float* U = static_cast<float*>(workspace);
    const int T = split_terms;
    {
        wg_gemm_args a = {};
        a.A = emb_tokens_bf16; a.lda = 2 * C; a.W = w_up; a.ldw = T * C; a.M = (int)rows; a.N = 4 * C; a.K = T * C;
        a.a_k_wrap = T == 3 ? 2 * C : 0;
        a.bias = b_up; a.bias_period = 1; a.out_mode = WG_OUT_F32; a.out = U; a.ldo = 4 * C;
        WG_TRY(wg_gemm(&a, s));
    }
    {
        Prof prof("dec_up2x_ln_gelu", s, 0.0, (double)rows * (4096.0 + 4096.0));
        up2x_ln_gelu_kernel<<<(unsigned)((rows * 4 + 7) / 8), 256, 0, s>>>(U, ln_g, ln_b, static_cast<__nv_bfloat16*>(out_tokens_bf16), rows, hw, grid_w);
    }
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

extern "C" int wg_mask_decoder_forward_level(const wg_mask_decoder_weights* w, const void* img_emb_tokens_bf16, const float* txt_emb,
                                             const int32_t* prompt_img, int P, int multimask_output, const float* prev_masks, int n_prev,
                                             float* low_res_out, float* iou_out, float* depth_pool_out, void* workspace, size_t workspace_bytes,
                                             void* stream_) {
    return wg_mask_decoder_forward_images(w, img_emb_tokens_bf16, 0, txt_emb, prompt_img, P, multimask_output, prev_masks, n_prev, low_res_out, iou_out,
                                          depth_pool_out, workspace, workspace_bytes, stream_);
}

extern "C" int wg_mask_decoder_forward_images(const wg_mask_decoder_weights* w, const void* img_emb_tokens_bf16, int n_images, const float* txt_emb,
                                              const int32_t* prompt_img, int P, int multimask_output, const float* prev_masks, int n_prev,
                                              float* low_res_out, float* iou_out, float* depth_pool_out, void* workspace, size_t workspace_bytes,
                                              void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(prev_masks == nullptr || n_prev > 0, "wg_mask_decoder_forward_level: previous masks need n_prev > 0");
    WG_REQUIRE(w != nullptr, "wg_mask_decoder_forward: null weights");
    if (P == 0) return WG_OK;
    WG_REQUIRE(img_emb_tokens_bf16 && txt_emb && prompt_img && low_res_out && iou_out && workspace, "wg_mask_decoder_forward: null pointer");
    WG_REQUIRE(P > 0, "wg_mask_decoder_forward: P=%d", P);
    WG_REQUIRE(w->split_terms == 2 || w->split_terms == 3, "wg_mask_decoder_forward: split_terms must be 2 or 3");
    WG_REQUIRE(w->n_mask_tokens == 4 && (w->up_stages == 1 || w->up_stages == 2), "wg_mask_decoder_forward: 4 mask tokens and 1 or 2 ConvTranspose stages are built");
    WG_REQUIRE(w->up_stages == 1 || (w->w_up2 && w->b_up2 && depth_pool_out == nullptr), "wg_mask_decoder_forward: the SAM head needs w_up2 / b_up2 and has no depth pooling");
    WG_REQUIRE(w->multimask_first == 0 || w->multimask_first == 1, "wg_mask_decoder_forward: multimask_first must be 0 or 1");
    const int hw = w->grid_h * w->grid_w;
    WG_REQUIRE(hw > 0 && hw <= 4096, "wg_mask_decoder_forward: grid %dx%d unsupported", w->grid_h, w->grid_w);
    WG_REQUIRE(depth_pool_out == nullptr || hw % 64 == 0, "wg_mask_decoder_forward: depth pooling needs hw %% 64 == 0");
    if (!device_is_sm100()) {
        set_error("wg_mask_decoder_forward: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    Workspace ws(workspace, workspace_bytes);
    DecBuffers d;
    WG_REQUIRE(carve(ws, P, hw, w->up_stages, d), "wg_mask_decoder_forward: workspace too small (%zu given, %zu needed)", workspace_bytes,
               wg_mask_decoder_workspace_bytes_ex(P, hw, w->up_stages));
    const long long rows = (long long)P * hw;
    WG_REQUIRE(rows < (1ll << 31), "wg_mask_decoder_forward: too many prompt tokens");
    // MaskDecoderMultiScale keeps index 0 in both modes (mask_decoder_multi_scale.py:126-132); SAM returns masks 1.. with
    // multimask_output and mask 0 without (mask_decoder.py:106-111)
    const int mask_start = multimask_output ? w->multimask_first : 0;
    const int n_out = multimask_output ? w->n_mask_tokens - mask_start : 1;
    const int T = w->split_terms;

    long long blocks = (rows * (C / 8) + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    {
        Prof prof("dec_expand_keys", s, 0.0, (double)rows * C * 8.0);
    expand_keys_kernel<<<(unsigned)blocks, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(img_emb_tokens_bf16), prompt_img, w->no_mask, d.keysA, P, hw,
                                                        prev_masks, n_prev);
    }
    WG_CHECK_CUDA(cudaGetLastError());

    const size_t tk_fixed = (size_t)(6 * NT * C + NT * 2048 + NT * (hw > 48 ? hw : 48)) * sizeof(float);
    const int tk_part = (tk_fixed + (size_t)TK_THREADS * 4 * NT * sizeof(float) <= 227 * 1024) ? TK_THREADS * 4 * NT : TK_THREADS * 2 * NT;
    const size_t tk_smem = tk_fixed + (size_t)tk_part * sizeof(float);
    WG_SMEM_OPT_IN(decoder_token_kernel, 227 * 1024);
    WG_REQUIRE(tk_smem <= 227 * 1024, "wg_mask_decoder_forward: token kernel shared memory %zu too large", tk_smem);

    // token->image attention as its own (head, prompt) kernel whenever P fits a grid dimension (WG_DEC_T2I_SPLIT=0: inside the token kernel)
    static const bool t2i_split_enabled = [] { const char* e = getenv("WG_DEC_T2I_SPLIT"); return e == nullptr || atoi(e) != 0; }();
    const int t2i_chunks = (hw + T2I_CHUNK - 1) / T2I_CHUNK;
    const bool split_t2i = t2i_split_enabled && P <= 65535;
    // Layer 0 of the two-way transformer sees the SAME keys for every prompt of an image (the keys only become prompt-specific after
    // the first image->token attention, transformer.py:151-242): its image-side projections [K_t2i | V_t2i | Q_i2t] are computed once
    // per image and indexed through prompt_img by the attention kernels.  Needs the image count (n_images > 0), no previous-level gate.
    static const bool share_enabled = [] { const char* e = getenv("WG_DEC_SHARE_L0"); return e == nullptr || atoi(e) != 0; }();
    const bool share0 = share_enabled && n_images > 0 && n_images < P && prev_masks == nullptr && split_t2i;

    TokArgs ta = {};
    ta.hw = hw;
    ta.part_floats = tk_part;
    ta.Xs = d.Xs;
    ta.n_mask_tokens = w->n_mask_tokens;
    ta.fin_wq_t = w->fin_wq_t; ta.fin_bq = w->fin_bq; ta.fin_wo_t = w->fin_wo_t; ta.fin_bo = w->fin_bo; ta.nf_g = w->nf_g; ta.nf_b = w->nf_b;
    ta.hyp_w0_t = w->hyp_w0_t; ta.hyp_w1_t = w->hyp_w1_t; ta.hyp_w2_t = w->hyp_w2_t;
    ta.hyp_b0 = w->hyp_b0; ta.hyp_b1 = w->hyp_b1; ta.hyp_b2 = w->hyp_b2;
    ta.iou_w0_t = w->iou_w0_t; ta.iou_w1_t = w->iou_w1_t; ta.iou_w2_t = w->iou_w2_t;
    ta.iou_b0 = w->iou_b0; ta.iou_b1 = w->iou_b1; ta.iou_b2 = w->iou_b2;
    ta.out_tokens = w->out_tokens; ta.sparse_add = w->sparse_add; ta.txt = txt_emb;
    ta.Tq = d.Tq; ta.Tpe = d.Tpe; ta.KT = d.KT; ta.VT = d.VT; ta.Qt = d.Qt; ta.Pm = d.Pm; ta.Pl = d.Pl; ta.Po = d.Po; ta.nchunk = t2i_chunks; ta.hyper = d.hyper; ta.iou = d.iou_all;

    // bring-up aid, compiled out of release builds (-DWG_DEBUG): WG_DEBUG_DECODER_STOP=<n> stops after the n-th launch group so
    // intermediates can be inspected; the call then reports WG_ERR_INVALID because the outputs were not produced
#ifdef WG_DEBUG
    const char* dbg = getenv("WG_DEBUG_DECODER_STOP");
    const int stop_at = dbg ? atoi(dbg) : 0;
    int step = 0;
#define WG_DBG_STEP()                                                                                   \
    do {                                                                                                \
        if (stop_at && ++step == stop_at) {                                                             \
            set_error("wg_mask_decoder_forward: stopped after launch group %d (WG_DEBUG_DECODER_STOP)", step); \
            return WG_ERR_INVALID;                                                                      \
        }                                                                                               \
    } while (0)
#else
#define WG_DBG_STEP() do { } while (0)
#endif

    __nv_bfloat16* keys = d.keysA;
    __nv_bfloat16* keys_next = d.keysB;
    for (int l = 0; l < 2; ++l) {
        const wg_twoway_layer& L = w->layers[l];
        const bool shared = share0 && l == 0 && L.mlp_w1_split != nullptr && L.mlp_w2_split != nullptr;  // (the all-in-one token kernel indexes kv per prompt)
        const int32_t* kv_img = shared ? prompt_img : nullptr;
        if (shared) {  // the images' own key sets (embedding + no-mask term) in the not yet used second key buffer
            long long blk = ((long long)n_images * hw * (C / 8) + 255) / 256;
            if (blk > 148 * 32) blk = 148 * 32;
            Prof prof("dec_expand_keys", s, 0.0, (double)n_images * hw * C * 8.0);
            expand_keys_kernel<<<(unsigned)blk, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(img_emb_tokens_bf16), nullptr, w->no_mask, keys_next, n_images, hw,
                                                         nullptr, 0);
        }
        {   // [K_t2i | V_t2i | Q_i2t] with the positional term as a per-position bias table
            wg_gemm_args a = {};
            a.A = shared ? keys_next : keys; a.lda = 2 * C; a.W = L.w_img; a.ldw = T * C; a.M = shared ? n_images * hw : (int)rows; a.N = 384; a.K = T * C;
            a.a_k_wrap = T == 3 ? 2 * C : 0;
            a.bias = L.b_img; a.bias_period = hw; a.out_mode = WG_OUT_F32; a.out = d.kvq; a.ldo = 384;
            WG_TRY(wg_gemm(&a, s));
        }
        WG_DBG_STEP();
        ta.phase = l; ta.L = L; ta.kv = d.kvq; ta.ldkv = 384;
        if (L.mlp_w1_split != nullptr && L.mlp_w2_split != nullptr) {
            // token MLP as two GEMMs over the P * NT token rows (hidden = relu(x W1^T + b1) in split-bf16; queries += hidden W2^T + b2
            // in place in fp32), between the two halves of the token kernel
            if (split_t2i) {
                ta.part = 3;
                {
                    Prof prof("dec_token", s, (double)P * 2.0 * 2.3e6, (double)P * 1.0e6);
                    decoder_token_kernel<<<P, TK_THREADS, tk_smem, s>>>(ta);
                }
                {
                    Prof prof("dec_t2i_attention", s, (double)P * 2.0 * 2.0 * NT * CI * hw, (double)P * hw * 1024.0);
                    t2i_partial_kernel<<<dim3(t2i_chunks, P), T2I_THREADS, 0, s>>>(d.Qt, d.kvq, 384, 0, CI, hw, d.Pm, d.Pl, d.Po, kv_img);
                }
                ta.part = 4;
                {
                    Prof prof("dec_token", s, (double)P * 2.0 * 0.3e6, (double)P * 0.2e6);
                    decoder_token_kernel<<<P, TK_THREADS, tk_smem, s>>>(ta);
                }
            } else {
                ta.part = 1;
                Prof prof("dec_token", s, (double)P * 2.0 * 3.0e6, (double)P * (hw * 512.0 + 1.0e6));
                decoder_token_kernel<<<P, TK_THREADS, tk_smem, s>>>(ta);
            }
            WG_CHECK_CUDA(cudaGetLastError());
            wg_gemm_args g1 = {};
            g1.A = d.Xs; g1.lda = 2 * C; g1.W = L.mlp_w1_split; g1.ldw = (long long)T * C; g1.M = P * NT; g1.N = 2048; g1.K = T * C;
            g1.a_k_wrap = T == 3 ? 2 * C : 0;
            g1.bias = L.mlp_b1; g1.act = WG_ACT_RELU; g1.out_mode = WG_OUT_BF16; g1.split_out = 1; g1.out = d.Hs; g1.ldo = 2 * 2048;
            WG_TRY(wg_gemm(&g1, s));
            wg_gemm_args g2 = {};
            g2.A = d.Hs; g2.lda = 2 * 2048; g2.W = L.mlp_w2_split; g2.ldw = (long long)T * 2048; g2.M = P * NT; g2.N = C; g2.K = T * 2048;
            g2.a_k_wrap = T == 3 ? 2 * 2048 : 0;
            g2.bias = L.mlp_b2; g2.out_mode = WG_OUT_F32; g2.out = d.Tq; g2.ldo = C; g2.resid = d.Tq;
            WG_TRY(wg_gemm(&g2, s));
            ta.part = 2;
            {
                Prof prof("dec_token", s, (double)P * 2.0 * 0.4e6, (double)P * 0.3e6);
                decoder_token_kernel<<<P, TK_THREADS, tk_smem, s>>>(ta);
            }
            ta.part = 0;
        } else {
            ta.part = 0;
            Prof prof("dec_token", s, (double)P * 2.0 * 10.2e6, (double)P * (hw * 512.0 + 3.0e6));
            decoder_token_kernel<<<P, TK_THREADS, tk_smem, s>>>(ta);
        }
        WG_CHECK_CUDA(cudaGetLastError());
        WG_DBG_STEP();
        {
            Prof prof("dec_i2t_attention", s, (double)rows * 8 * 6 * 16 * 4.0, (double)rows * 512.0);
            i2t_attention_kernel<<<dim3((hw + 255) / 256, P), 256, 0, s>>>(d.kvq, 384, 256, d.KT, d.VT, d.a2, hw, kv_img);
        }
        WG_CHECK_CUDA(cudaGetLastError());
        WG_DBG_STEP();
        {   // keys' = LayerNorm4(keys + attn W_o^T + b)
            wg_gemm_args a = {};
            a.A = d.a2; a.lda = 2 * CI; a.W = L.i2t_wo; a.ldw = T * CI; a.M = (int)rows; a.N = C; a.K = T * CI;
            a.a_k_wrap = T == 3 ? 2 * CI : 0;
            a.bias = L.i2t_bo; a.bias_period = 1; a.out_mode = WG_OUT_BF16_LN; a.split_out = 1; a.out = keys_next; a.ldo = 2 * C; a.resid = keys;
            a.ln_gamma = L.n4_g; a.ln_beta = L.n4_b; a.ln_eps = 1e-5f;
            WG_TRY(wg_gemm(&a, s));
        }
        WG_DBG_STEP();
        __nv_bfloat16* t = keys; keys = keys_next; keys_next = t;
    }
    {   // final token->image attention: [K | V]
        wg_gemm_args a = {};
        a.A = keys; a.lda = 2 * C; a.W = w->w_img_fin; a.ldw = T * C; a.M = (int)rows; a.N = 256; a.K = T * C;
        a.a_k_wrap = T == 3 ? 2 * C : 0;
        a.bias = w->b_img_fin; a.bias_period = hw; a.out_mode = WG_OUT_F32; a.out = d.kvq; a.ldo = 256;
        WG_TRY(wg_gemm(&a, s));
    }
    ta.phase = 2; ta.kv = d.kvq; ta.ldkv = 256;
    if (split_t2i) {
        ta.part = 3;
        {
            Prof prof("dec_token", s, (double)P * 2.0 * 0.2e6, (double)P * 0.2e6);
            decoder_token_kernel<<<P, TK_THREADS, tk_smem, s>>>(ta);
        }
        {
            Prof prof("dec_t2i_attention", s, (double)P * 2.0 * 2.0 * NT * CI * hw, (double)P * hw * 1024.0);
            t2i_partial_kernel<<<dim3(t2i_chunks, P), T2I_THREADS, 0, s>>>(d.Qt, d.kvq, 256, 0, CI, hw, d.Pm, d.Pl, d.Po, nullptr);
        }
        ta.part = 4;
        {
            Prof prof("dec_token", s, (double)P * 2.0 * 2.8e6, (double)P * 1.5e6);
            decoder_token_kernel<<<P, TK_THREADS, tk_smem, s>>>(ta);
        }
        ta.part = 0;
    } else {
        ta.part = 0;
        Prof prof("dec_token", s, (double)P * 2.0 * 3.0e6, (double)P * (hw * 512.0 + 1.5e6));
        decoder_token_kernel<<<P, TK_THREADS, tk_smem, s>>>(ta);
    }
    WG_CHECK_CUDA(cudaGetLastError());
    if (w->up_stages == 1) {
        // ConvTranspose2d(256 -> 32, k=2, s=2) as a GEMM over positions (N = 4 sub-pixels x 32 channels), fp32 out
        {
            wg_gemm_args a = {};
            a.A = keys; a.lda = 2 * C; a.W = w->w_up; a.ldw = T * C; a.M = (int)rows; a.N = 128; a.K = T * C;
            a.a_k_wrap = T == 3 ? 2 * C : 0;
            a.bias = w->b_up; a.bias_period = 1; a.out_mode = WG_OUT_F32; a.out = d.U; a.ldo = 128;
            WG_TRY(wg_gemm(&a, s));
        }
        if (depth_pool_out) WG_CHECK_CUDA(cudaMemsetAsync(depth_pool_out, 0, (size_t)P * 33 * sizeof(float), s));
        {
            Prof prof("dec_upscale_mask", s, (double)rows * 128 * 12.0, (double)rows * (512.0 + 16.0 * n_out));
            upscale_mask_kernel<<<(unsigned)((rows + 63) / 64), 256, 0, s>>>(d.U, w->up_ln_g, w->up_ln_b, d.hyper, w->n_mask_tokens, mask_start, n_out,
                                                                          low_res_out, depth_pool_out, hw, w->grid_w, rows);
        }
        WG_CHECK_CUDA(cudaGetLastError());
    } else {
        // SAM head (mask_decoder.py:53-63): ConvT(256 -> 64) -> LayerNorm2d -> GELU -> ConvT(64 -> 32) -> GELU, both as GEMMs.
        // The buffers of the finished transformer are reused: kvq holds the first ConvT output, keys_next its activations.
        WG_REQUIRE(rows * 4 < (1ll << 31), "wg_mask_decoder_forward: too many up-sampled positions");
        float* U1 = d.kvq;                 // fp32 [rows, 4 * 64]
        __nv_bfloat16* A2 = keys_next;     // split-bf16 [rows * 4, 2 * 64]
        {
            wg_gemm_args a = {};
            a.A = keys; a.lda = 2 * C; a.W = w->w_up; a.ldw = T * C; a.M = (int)rows; a.N = 256; a.K = T * C;
            a.a_k_wrap = T == 3 ? 2 * C : 0;
            a.bias = w->b_up; a.bias_period = 1; a.out_mode = WG_OUT_F32; a.out = U1; a.ldo = 256;
            WG_TRY(wg_gemm(&a, s));
        }
        {
            Prof prof("dec_up1_ln_gelu", s, 0.0, (double)rows * (1024.0 + 1024.0));
            up1_ln_gelu_kernel<<<(unsigned)((rows * 4 + 255) / 256), 256, 0, s>>>(U1, w->up_ln_g, w->up_ln_b, A2, rows * 4);
        }
        WG_CHECK_CUDA(cudaGetLastError());
        {
            wg_gemm_args a = {};
            a.A = A2; a.lda = 128; a.W = w->w_up2; a.ldw = T * 64; a.M = (int)(rows * 4); a.N = 128; a.K = T * 64;
            a.a_k_wrap = T == 3 ? 128 : 0;
            a.bias = w->b_up2; a.bias_period = 1; a.out_mode = WG_OUT_F32; a.out = d.U; a.ldo = 128;
            WG_TRY(wg_gemm(&a, s));
        }
        {
            Prof prof("dec_upscale_mask", s, (double)rows * 4 * 128 * 6.0, (double)rows * 4 * (512.0 + 16.0 * n_out));
            upscale_mask2_kernel<<<(unsigned)((rows * 4 + 63) / 64), 256, 0, s>>>(d.U, d.hyper, w->n_mask_tokens, mask_start, n_out, low_res_out, hw,
                                                                               w->grid_w, rows * 4);
        }
        WG_CHECK_CUDA(cudaGetLastError());
    }
    // iou predictions for the selected masks
    WG_CHECK_CUDA(cudaMemcpy2DAsync(iou_out, n_out * sizeof(float), d.iou_all + mask_start, w->n_mask_tokens * sizeof(float), n_out * sizeof(float), P,
                                    cudaMemcpyDeviceToDevice, s));
    return WG_OK;
}
