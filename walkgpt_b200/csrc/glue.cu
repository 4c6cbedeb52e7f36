// SURVEY §8(f) rows 2 and 4 -- the small steps on either side of the grounding path that the reference runs as chains of
// eager PyTorch ops with host synchronisation:
//   F2a  visual-token resample handed to the LLM        (model/llava_walkgpt/model/llava_arch.py:252-259)
//   F2b  [SEG]-row extraction from the LLM hidden states (model/walkgpt.py:287-306, 406-420)
//   F4   histogram intersection / union scoring          (utils/utils.py:192-204, intersectionAndUnionGPU)
// All three are bandwidth-bound index / byte work: coalesced 16-byte accesses, integer arithmetic, bit-exact results.
#include "internal.h"
#include "ptx.cuh"

namespace wg {
namespace {

// PyTorch upsample_bilinear2d source index (align_corners = False)
__device__ __forceinline__ void src_index_ac0(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float s = fmaxf(scale * (dst + 0.5f) - 0.5f, 0.f);
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = s - (float)i0;
}

// ---- F2a: tokens [n, p*p, C] (channels last) -> [n, t*t, C]; fp32 arithmetic in PyTorch's order, output in the input dtype
template <typename T>
__global__ void __launch_bounds__(256) resample_tokens_kernel(const T* __restrict__ in, T* __restrict__ out, int p, int t, int C, float scale) {
    // one CTA per output row (oy) of one image: threads sweep (ox, 4-channel group), so neighbouring threads read and write
    // neighbouring 8/16-byte pieces
    const int n = blockIdx.y, oy = blockIdx.x;
    int y0, y1;
    float ly;
    src_index_ac0(oy, scale, p, y0, y1, ly);
    const T* base = in + (size_t)n * p * p * C;
    const int cv = C / 4;
    for (int idx = threadIdx.x; idx < t * cv; idx += blockDim.x) {
        const int ox = idx / cv, ch = (idx - ox * cv) * 4;
        int x0, x1;
        float lx;
        src_index_ac0(ox, scale, p, x0, x1, lx);
        const T* a = base + (size_t)(y0 * p + x0) * C;
        const T* b = base + (size_t)(y0 * p + x1) * C;
        const T* c = base + (size_t)(y1 * p + x0) * C;
        const T* d = base + (size_t)(y1 * p + x1) * C;
        T* o = out + ((size_t)n * t * t + (size_t)oy * t + ox) * C;
        float va[4], vb[4], vc[4], vd[4], r[4];
        if constexpr (sizeof(T) == 2) {
            const uint2 ua = *reinterpret_cast<const uint2*>(a + ch), ub = *reinterpret_cast<const uint2*>(b + ch);
            const uint2 uc = *reinterpret_cast<const uint2*>(c + ch), ud = *reinterpret_cast<const uint2*>(d + ch);
            auto unpack = [](uint2 u, float (&v)[4]) {
                const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
                const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
                v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
            };
            unpack(ua, va); unpack(ub, vb); unpack(uc, vc); unpack(ud, vd);
        } else {
            const float4 fa = *reinterpret_cast<const float4*>(a + ch), fb = *reinterpret_cast<const float4*>(b + ch);
            const float4 fc = *reinterpret_cast<const float4*>(c + ch), fd = *reinterpret_cast<const float4*>(d + ch);
            va[0] = fa.x; va[1] = fa.y; va[2] = fa.z; va[3] = fa.w;
            vb[0] = fb.x; vb[1] = fb.y; vb[2] = fb.z; vb[3] = fb.w;
            vc[0] = fc.x; vc[1] = fc.y; vc[2] = fc.z; vc[3] = fc.w;
            vd[0] = fd.x; vd[1] = fd.y; vd[2] = fd.z; vd[3] = fd.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float top = (1.f - lx) * va[e] + lx * vb[e];
            const float bot = (1.f - lx) * vc[e] + lx * vd[e];
            r[e] = (1.f - ly) * top + ly * bot;
        }
        if constexpr (sizeof(T) == 2) {
            uint2 u;
            u.x = pack_bf16x2(r[0], r[1]);
            u.y = pack_bf16x2(r[2], r[3]);
            *reinterpret_cast<uint2*>(o + ch) = u;
        } else {
            *reinterpret_cast<float4*>(o + ch) = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
}

// ---- F2b: [SEG]-row extraction
// hidden position q of text row r is a [SEG] row iff q >= shift and i = q - shift + 1 < Lin and input_ids[r][i] is one of the
// seg ids  (model/walkgpt.py:287-306: mask over input_ids[:, 1:], one false appended, `shift` = 255 falses prepended).
__device__ __forceinline__ bool is_seg(const long long* __restrict__ ids, int Lin, int shift, int q, const long long* seg_ids, int n_ids) {
    const int i = q - shift + 1;
    if (q < shift || i >= Lin) return false;
    const long long v = ids[i];
    bool hit = false;
    for (int k = 0; k < n_ids; ++k) hit = hit || (v == seg_ids[k]);
    return hit;
}

struct SegIds {
    long long v[8];
    int n;
};

// one CTA per text row: counts[r] = number of [SEG] positions
__global__ void __launch_bounds__(256) seg_count_kernel(const long long* __restrict__ ids, int Lin, int L, int shift, SegIds seg, int* __restrict__ counts) {
    const int r = blockIdx.x;
    const long long* row = ids + (size_t)r * Lin;
    int c = 0;
    for (int q = threadIdx.x; q < L; q += 256) c += is_seg(row, Lin, shift, q, seg.v, seg.n) ? 1 : 0;
    __shared__ int red[8];
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        counts[r] = t;
    }
}
// exclusive scan of the row counts (single CTA), image offsets = row_offsets[offset[b]], total
__global__ void __launch_bounds__(1024) seg_scan_kernel(const int* __restrict__ counts, int rows, int* __restrict__ row_offsets,
                                                        const int* __restrict__ img_rows, int n_img_p1, int* __restrict__ img_offsets) {
    __shared__ int carry;
    __shared__ int wsum[32];
    if (threadIdx.x == 0) {
        carry = 0;
        row_offsets[0] = 0;
    }
    __syncthreads();
    for (int base = 0; base < rows; base += 1024) {
        const int i = base + threadIdx.x;
        int v = i < rows ? counts[i] : 0;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const int before = carry + (warp > 0 ? wsum[warp - 1] : 0);
        if (i < rows) row_offsets[i + 1] = before + incl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + incl;
        __syncthreads();
    }
    if (img_rows != nullptr)
        for (int b = threadIdx.x; b < n_img_p1; b += 1024) img_offsets[b] = row_offsets[img_rows[b]];
}
// one CTA per text row: ordered compaction of its [SEG] positions, then the gather of their hidden rows
template <typename T>
__global__ void __launch_bounds__(256) seg_gather_kernel(const long long* __restrict__ ids, int Lin, int L, int shift, SegIds seg,
                                                         const int* __restrict__ row_offsets, const T* __restrict__ hidden, int H,
                                                         T* __restrict__ out, int max_out) {
    const int r = blockIdx.x;
    const long long* row = ids + (size_t)r * Lin;
    const int first = row_offsets[r], cnt = row_offsets[r + 1] - first;
    if (cnt == 0) return;
    extern __shared__ int pos[];  // [cnt] hidden positions of this row's [SEG] tokens, in order
    __shared__ int wtot[8];
    __shared__ int running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < L; base += 256) {
        const int q = base + threadIdx.x;
        const bool hit = q < L && is_seg(row, Lin, shift, q, seg.v, seg.n);
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) wtot[warp] = __popc(bal);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += wtot[w];
        if (hit) pos[before + __popc(bal & ((1u << lane) - 1u))] = q;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += wtot[w];
            running += t;
        }
        __syncthreads();
    }
    // gather: 16-byte pieces of the selected hidden rows
    constexpr int VEC = 16 / sizeof(T);
    const int pieces = H / VEC;
    for (int k = 0; k < cnt; ++k) {
        const int orow = first + k;
        if (orow >= max_out) break;
        const uint4* src = reinterpret_cast<const uint4*>(hidden + ((size_t)r * L + pos[k]) * H);
        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)orow * H);
        for (int i = threadIdx.x; i < pieces; i += 256) dst[i] = src[i];
    }
}

// ---- F4: per-mask class histograms.  acc[n][3][K] (int): intersection, output, target; ignore_index pixels are dropped from all
__global__ void __launch_bounds__(256) iou_hist_kernel(const uint8_t* __restrict__ output, const uint8_t* __restrict__ target, long long pixels, int K,
                                                       int ignore_index, int* __restrict__ acc) {
    const int n = blockIdx.y;
    const uint8_t* o = output + (size_t)n * pixels;
    const uint8_t* t = target + (size_t)n * pixels;
    __shared__ int hist[3 * 8];
    if (threadIdx.x < 3 * 8) hist[threadIdx.x] = 0;
    __syncthreads();
    int li[8], lo[8], lt[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) li[k] = lo[k] = lt[k] = 0;
    auto tally = [&](int ov, int tv) {
        // utils/utils.py:196-201: output[target == ignore] = ignore; histc over [0, K-1] drops everything outside the range
        if (tv == ignore_index) ov = ignore_index;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < K) {
                lo[k] += (ov == k);
                lt[k] += (tv == k);
                li[k] += (ov == k && tv == k);
            }
        }
    };
    const bool aligned = ((reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(t)) & 15) == 0;
    const long long vecs = aligned ? pixels / 16 : 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < vecs; i += (long long)gridDim.x * 256) {
        const uint4 uo = __ldg(reinterpret_cast<const uint4*>(o) + i), ut = __ldg(reinterpret_cast<const uint4*>(t) + i);
        const uint8_t* bo = reinterpret_cast<const uint8_t*>(&uo);
        const uint8_t* bt = reinterpret_cast<const uint8_t*>(&ut);
#pragma unroll
        for (int e = 0; e < 16; ++e) tally(bo[e], bt[e]);
    }
    for (long long i = vecs * 16 + (long long)blockIdx.x * 256 + threadIdx.x; i < pixels; i += (long long)gridDim.x * 256) tally(o[i], t[i]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < K) {
            const int a = __reduce_add_sync(0xffffffffu, li[k]), b = __reduce_add_sync(0xffffffffu, lo[k]), c = __reduce_add_sync(0xffffffffu, lt[k]);
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&hist[k], a);
                atomicAdd(&hist[8 + k], b);
                atomicAdd(&hist[16 + k], c);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 * K) {
        const int which = threadIdx.x / K, k = threadIdx.x % K;
        atomicAdd(&acc[((size_t)n * 3 + which) * K + k], hist[which * 8 + k]);
    }
}
// [n][3][K] ints -> floats: intersection, union = output + target - intersection, target  (histc returns float tensors)
__global__ void iou_finalize_kernel(const int* __restrict__ acc, float* __restrict__ out, int n, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * K) return;
    const int m = i / K, k = i % K;
    const int inter = acc[((size_t)m * 3 + 0) * K + k], ao = acc[((size_t)m * 3 + 1) * K + k], at = acc[((size_t)m * 3 + 2) * K + k];
    out[((size_t)m * 3 + 0) * K + k] = (float)inter;
    out[((size_t)m * 3 + 1) * K + k] = (float)(ao + at - inter);
    out[((size_t)m * 3 + 2) * K + k] = (float)at;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Matching cost of match_pred (utils/matcher.py:93-128): every predicted and every target mask is sampled at the same P points
// (point_sample = F.grid_sample, bilinear, zero padding, align_corners=False), then
//   C[n, m] = (sum_p pos[n,p] t[m,p] + neg[n,p] (1 - t[m,p])) / P  +  1 - (2 sum_p sig[n,p] t[m,p] + 1) / (sum_p sig[n,p] + sum_p t[m,p] + 1)
// with pos / neg = BCE-with-logits against 1 / 0 (matcher.py:33-58) and sig = sigmoid (matcher.py:10-26).
// One CTA per chunk of MATCH_PTS points: each thread owns one point (tap offsets and weights computed once, reused for every
// mask), the sampled values go to shared memory, then the (n, m) pairs are spread over the threads.  Per-chunk partial sums are
// written to the workspace and added in chunk order by the finalize kernel: no floating-point atomics, the same bits every run.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int MATCH_PTS = 128;
constexpr int MATCH_LD = MATCH_PTS + 1;  // row pitch of the shared arrays: threads of a warp read different rows at the same point

struct Taps {
    int o00, o01, o10, o11;    // offsets of the four neighbours (clamped; a neighbour outside the map has weight 0)
    float w00, w01, w10, w11;  // nw, ne, sw, se
};
__device__ __forceinline__ Taps make_taps(float u, float v, int H, int W) {
    // point_sample: grid = 2 p - 1; grid_sample(align_corners=False): x = ((grid + 1) W - 1) / 2
    const float gx = 2.0f * u - 1.0f, gy = 2.0f * v - 1.0f;
    const float x = ((gx + 1.0f) * W - 1.0f) * 0.5f, y = ((gy + 1.0f) * H - 1.0f) * 0.5f;
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = (int)xf, y0 = (int)yf, x1 = x0 + 1, y1 = y0 + 1;
    // weights as grid_sample forms them: differences against the integer corners, not (1 - fraction)
    Taps t;
    t.w00 = ((float)x1 - x) * ((float)y1 - y);
    t.w01 = (x - (float)x0) * ((float)y1 - y);
    t.w10 = ((float)x1 - x) * (y - (float)y0);
    t.w11 = (x - (float)x0) * (y - (float)y0);
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
    if (!(vx0 && vy0)) t.w00 = 0.f;
    if (!(vx1 && vy0)) t.w01 = 0.f;
    if (!(vx0 && vy1)) t.w10 = 0.f;
    if (!(vx1 && vy1)) t.w11 = 0.f;
    const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1), cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
    t.o00 = cy0 * W + cx0; t.o01 = cy0 * W + cx1; t.o10 = cy1 * W + cx0; t.o11 = cy1 * W + cx1;
    return t;
}
template <typename T>
__device__ __forceinline__ float sample4(const T* __restrict__ img, const Taps& t) {
    float acc = (float)__ldg(img + t.o00) * t.w00;
    acc += (float)__ldg(img + t.o01) * t.w01;
    acc += (float)__ldg(img + t.o10) * t.w10;
    acc += (float)__ldg(img + t.o11) * t.w11;
    return acc;
}

// partial layout per chunk: [n_pred * n_tgt] ce sums | [n_pred * n_tgt] dice numerators | [n_pred] sigmoid sums | [n_tgt] target sums
template <typename TT>
__global__ void __launch_bounds__(MATCH_PTS) match_sample_kernel(const float* __restrict__ pred, const TT* __restrict__ tgt, const float* __restrict__ pts,
                                                                  int n_pred, int n_tgt, int H, int W, int P, float* __restrict__ partial) {
    extern __shared__ float sm[];
    float* s_pos = sm;                          // [n_pred][MATCH_LD]
    float* s_neg = s_pos + n_pred * MATCH_LD;
    float* s_sig = s_neg + n_pred * MATCH_LD;
    float* s_t = s_sig + n_pred * MATCH_LD;     // [n_tgt][MATCH_LD]
    const int tid = threadIdx.x;
    const int p = blockIdx.x * MATCH_PTS + tid;
    const bool live = p < P;
    const size_t hw = (size_t)H * W;
    if (live) {
        const Taps t = make_taps(pts[2 * p], pts[2 * p + 1], H, W);
        for (int n = 0; n < n_pred; ++n) {
            const float x = sample4(pred + n * hw, t);
            // binary_cross_entropy_with_logits(x, y) = (1 - y) x + max(-x, 0) + log1p(exp(-|x|))
            const float sp = fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
            s_pos[n * MATCH_LD + tid] = sp;
            s_neg[n * MATCH_LD + tid] = x + sp;
            s_sig[n * MATCH_LD + tid] = 1.0f / (1.0f + expf(-x));
        }
        for (int m = 0; m < n_tgt; ++m) s_t[m * MATCH_LD + tid] = sample4(tgt + m * hw, t);
    } else {  // past the last point: contributes nothing to any sum
        for (int n = 0; n < n_pred; ++n) { s_pos[n * MATCH_LD + tid] = 0.f; s_neg[n * MATCH_LD + tid] = 0.f; s_sig[n * MATCH_LD + tid] = 0.f; }
        for (int m = 0; m < n_tgt; ++m) s_t[m * MATCH_LD + tid] = 0.f;
    }
    __syncthreads();
    const int pairs = n_pred * n_tgt;
    float* out = partial + (size_t)blockIdx.x * (2 * pairs + n_pred + n_tgt);
    for (int pr = tid; pr < pairs; pr += MATCH_PTS) {
        const float* ps = s_pos + (pr / n_tgt) * MATCH_LD;
        const float* ng = s_neg + (pr / n_tgt) * MATCH_LD;
        const float* sg = s_sig + (pr / n_tgt) * MATCH_LD;
        const float* tt = s_t + (pr % n_tgt) * MATCH_LD;
        float ce = 0.f, dn = 0.f;
#pragma unroll 8
        for (int i = 0; i < MATCH_PTS; ++i) {
            const float t = tt[i];
            ce += ps[i] * t + ng[i] * (1.0f - t);
            dn += sg[i] * t;
        }
        out[pr] = ce;
        out[pairs + pr] = dn;
    }
    for (int r = tid; r < n_pred + n_tgt; r += MATCH_PTS) {
        const float* row = r < n_pred ? s_sig + r * MATCH_LD : s_t + (r - n_pred) * MATCH_LD;
        float acc = 0.f;
#pragma unroll 8
        for (int i = 0; i < MATCH_PTS; ++i) acc += row[i];
        out[2 * pairs + r] = acc;
    }
}
__global__ void match_finalize_kernel(const float* __restrict__ partial, int chunks, int n_pred, int n_tgt, int P, float* __restrict__ cost) {
    const int pr = blockIdx.x * blockDim.x + threadIdx.x;
    const int pairs = n_pred * n_tgt;
    if (pr >= pairs) return;
    const int n = pr / n_tgt, m = pr % n_tgt;
    const size_t stride = 2 * (size_t)pairs + n_pred + n_tgt;
    float ce = 0.f, dn = 0.f, ss = 0.f, st = 0.f;
    for (int c = 0; c < chunks; ++c) {
        const float* q = partial + c * stride;
        ce += q[pr];
        dn += q[pairs + pr];
        ss += q[2 * pairs + n];
        st += q[2 * pairs + n_pred + m];
    }
    cost[pr] = ce / (float)P + (1.0f - (2.0f * dn + 1.0f) / (ss + st + 1.0f));
}

}  // namespace
}  // namespace wg

using namespace wg;

extern "C" int wg_resample_tokens(const void* tokens, int is_bf16, int n, int p, int C, int t, void* out, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    if (n == 0) return WG_OK;
    WG_REQUIRE(tokens && out, "wg_resample_tokens: null pointer");
    WG_REQUIRE(n > 0 && n <= 65535 && p > 0 && t > 0 && C > 0 && C % 4 == 0, "wg_resample_tokens: bad sizes n=%d p=%d t=%d C=%d", n, p, t, C);
    WG_REQUIRE((reinterpret_cast<uintptr_t>(tokens) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "wg_resample_tokens: pointers must be 16-byte aligned");
    if (!device_is_sm100()) {
        set_error("wg_resample_tokens: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const float scale = (float)p / (float)t;
    Prof prof("resample_tokens", s, 0.0, (double)n * ((double)p * p + (double)t * t) * C * (is_bf16 ? 2.0 : 4.0));
    if (is_bf16)
        resample_tokens_kernel<__nv_bfloat16><<<dim3(t, n), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(tokens), static_cast<__nv_bfloat16*>(out), p, t, C, scale);
    else
        resample_tokens_kernel<float><<<dim3(t, n), 256, 0, s>>>(static_cast<const float*>(tokens), static_cast<float*>(out), p, t, C, scale);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

// prompt -> image index from the per-image [SEG] offsets, on the device (no host list, no synchronisation).  One thread per
// prompt: binary search for the image whose [offsets[b], offsets[b+1]) holds it.  Offsets that are not a non-decreasing
// sequence from 0 to P cannot index out of bounds: the result is clamped to [0, B-1] and *status is set to 1.
namespace wg {
__global__ void __launch_bounds__(256) prompt_index_kernel(const int* __restrict__ offs, int B, int P, int* __restrict__ prompt_img, int* __restrict__ status) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0 && status != nullptr) {
        int bad = (offs[0] != 0) || (offs[B] != P);
        for (int b = 0; b < B; ++b) bad |= (offs[b + 1] < offs[b]);
        *status = bad;
    }
    if (p >= P) return;
    int lo = 0, hi = B;  // largest b in [0, B-1] with offs[b] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offs[mid] <= p) lo = mid; else hi = mid;
    }
    prompt_img[p] = lo;
}
}  // namespace wg

extern "C" int wg_prompt_index(const int32_t* seg_offsets, int B, int P, int32_t* prompt_img, int32_t* status, void* stream_) {
    using namespace wg;
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(seg_offsets && prompt_img && B > 0 && P >= 0, "wg_prompt_index: bad arguments");
    if (!device_is_sm100()) {
        set_error("wg_prompt_index: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    Prof prof("prompt_index", s);
    prompt_index_kernel<<<(P > 0 ? (P + 255) / 256 : 1), 256, 0, s>>>(seg_offsets, B, P, prompt_img, status);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

extern "C" int wg_seg_gather(const int64_t* input_ids, int rows, int Lin, const void* hidden, int hidden_is_bf16, int L, int H, const int64_t* seg_ids,
                             int n_seg_ids, int shift, const int32_t* img_rows, int n_img_p1, void* out, int max_out, int32_t* counts,
                             int32_t* row_offsets, int32_t* img_offsets, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(input_ids && hidden && seg_ids && out && counts && row_offsets, "wg_seg_gather: null pointer");
    WG_REQUIRE(rows > 0 && Lin > 0 && L > 0 && H > 0 && max_out >= 0, "wg_seg_gather: bad sizes");
    WG_REQUIRE(n_seg_ids >= 1 && n_seg_ids <= 8, "wg_seg_gather: 1..8 seg token ids (got %d)", n_seg_ids);
    WG_REQUIRE(shift >= 0 && L == Lin + shift, "wg_seg_gather: hidden length %d must equal input length %d + shift %d (model/walkgpt.py:295-306)", L, Lin, shift);
    WG_REQUIRE((H * (hidden_is_bf16 ? 2 : 4)) % 16 == 0 && (reinterpret_cast<uintptr_t>(hidden) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "wg_seg_gather: rows must be multiples of 16 bytes and 16-byte aligned");
    WG_REQUIRE(img_rows == nullptr || (img_offsets != nullptr && n_img_p1 > 0), "wg_seg_gather: img_offsets missing");
    if (!device_is_sm100()) {
        set_error("wg_seg_gather: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    SegIds seg;
    seg.n = n_seg_ids;
    for (int k = 0; k < 8; ++k) seg.v[k] = k < n_seg_ids ? seg_ids[k] : 0;  // seg_ids is a HOST array
    const long long* ids = reinterpret_cast<const long long*>(input_ids);
    Prof prof("seg_gather", s, 0.0, (double)rows * Lin * 8.0 + 2.0 * (double)max_out * H * (hidden_is_bf16 ? 2.0 : 4.0), 3);
    seg_count_kernel<<<rows, 256, 0, s>>>(ids, Lin, L, shift, seg, counts);
    seg_scan_kernel<<<1, 1024, 0, s>>>(counts, rows, row_offsets, img_rows, n_img_p1, img_offsets);
    const size_t smem = (size_t)L * sizeof(int);
    WG_REQUIRE(smem <= 200 * 1024, "wg_seg_gather: sequence length %d too long", L);
    if (hidden_is_bf16) {
        WG_SMEM_OPT_IN((seg_gather_kernel<__nv_bfloat16>), 200 * 1024);
        seg_gather_kernel<__nv_bfloat16><<<rows, 256, smem, s>>>(ids, Lin, L, shift, seg, row_offsets, static_cast<const __nv_bfloat16*>(hidden), H,
                                                               static_cast<__nv_bfloat16*>(out), max_out);
    } else {
        WG_SMEM_OPT_IN((seg_gather_kernel<float>), 200 * 1024);
        seg_gather_kernel<float><<<rows, 256, smem, s>>>(ids, Lin, L, shift, seg, row_offsets, static_cast<const float*>(hidden), H, static_cast<float*>(out), max_out);
    }
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

extern "C" size_t wg_intersection_and_union_workspace_bytes(int n_masks, int K) { return (size_t)(n_masks > 0 ? n_masks : 0) * 3 * (K > 0 ? K : 0) * sizeof(int); }

extern "C" int wg_intersection_and_union(const uint8_t* output, const uint8_t* target, int n_masks, int64_t pixels, int K, int ignore_index, float* out,
                                         void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    if (n_masks == 0) return WG_OK;
    WG_REQUIRE(output && target && out && workspace, "wg_intersection_and_union: null pointer");
    WG_REQUIRE(n_masks > 0 && n_masks <= 65535 && pixels > 0 && K >= 1 && K <= 8, "wg_intersection_and_union: bad sizes (K must be 1..8)");
    WG_REQUIRE(pixels < (1ll << 31), "wg_intersection_and_union: at most 2^31 - 1 pixels per mask");
    WG_REQUIRE(workspace_bytes >= wg_intersection_and_union_workspace_bytes(n_masks, K), "wg_intersection_and_union: workspace too small");
    if (!device_is_sm100()) {
        set_error("wg_intersection_and_union: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    int* acc = static_cast<int*>(workspace);
    WG_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)n_masks * 3 * K * sizeof(int), s));
    long long chunks = (pixels / 16 + 255) / 256;
    const long long want = (8LL * device_sm_count() + n_masks - 1) / n_masks;  // ~8 CTAs per SM over all masks
    if (chunks > want) chunks = want;
    if (chunks < 1) chunks = 1;
    Prof prof("intersection_and_union", s, 0.0, 2.0 * (double)n_masks * (double)pixels, 2);
    iou_hist_kernel<<<dim3((unsigned)chunks, n_masks), 256, 0, s>>>(output, target, pixels, K, ignore_index, acc);
    iou_finalize_kernel<<<(n_masks * K + 255) / 256, 256, 0, s>>>(acc, out, n_masks, K);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}


static size_t match_chunks(int P) { return (size_t)((P > 0 ? P : 0) + MATCH_PTS - 1) / MATCH_PTS; }

extern "C" size_t wg_match_cost_workspace_bytes(int n_pred, int n_tgt, int num_points) {
    if (n_pred <= 0 || n_tgt <= 0) return 0;
    return match_chunks(num_points) * (2 * (size_t)n_pred * n_tgt + n_pred + n_tgt) * sizeof(float);
}

extern "C" int wg_match_cost(const float* pred_logits, const void* tgt_masks, int tgt_is_u8, const float* point_coords, int n_pred, int n_tgt, int H, int W,
                             int num_points, float* cost, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    if (n_pred == 0 || n_tgt == 0) return WG_OK;  // empty cost matrix (linear_sum_assignment returns no pairs)
    WG_REQUIRE(pred_logits && tgt_masks && point_coords && cost && workspace, "wg_match_cost: null pointer");
    WG_REQUIRE(n_pred > 0 && n_tgt > 0 && n_pred <= 64 && n_tgt <= 64, "wg_match_cost: 1..64 predicted and target masks (got %d, %d)", n_pred, n_tgt);
    WG_REQUIRE(H > 0 && W > 0 && (long long)H * W < (1ll << 31) && num_points > 0, "wg_match_cost: bad sizes H=%d W=%d points=%d", H, W, num_points);
    WG_REQUIRE(workspace_bytes >= wg_match_cost_workspace_bytes(n_pred, n_tgt, num_points), "wg_match_cost: workspace too small");
    if (!device_is_sm100()) {
        set_error("wg_match_cost: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const int chunks = (int)match_chunks(num_points);
    const size_t smem = (size_t)(3 * n_pred + n_tgt) * MATCH_LD * sizeof(float);
    float* partial = static_cast<float*>(workspace);
    Prof prof("match_cost", s, 0.0, 16.0 * (double)(n_pred + n_tgt) * num_points, 2);
    if (tgt_is_u8) {
        WG_SMEM_OPT_IN((match_sample_kernel<uint8_t>), 4 * 64 * MATCH_LD * 4);
        match_sample_kernel<uint8_t><<<chunks, MATCH_PTS, smem, s>>>(pred_logits, static_cast<const uint8_t*>(tgt_masks), point_coords, n_pred, n_tgt, H, W, num_points, partial);
    } else {
        WG_SMEM_OPT_IN((match_sample_kernel<float>), 4 * 64 * MATCH_LD * 4);
        match_sample_kernel<float><<<chunks, MATCH_PTS, smem, s>>>(pred_logits, static_cast<const float*>(tgt_masks), point_coords, n_pred, n_tgt, H, W, num_points, partial);
    }
    match_finalize_kernel<<<(n_pred * n_tgt + 127) / 128, 128, 0, s>>>(partial, chunks, n_pred, n_tgt, num_points, cost);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}
