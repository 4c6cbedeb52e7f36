// Bandwidth-bound tails:
//   A8  postprocess_masks: bilinear(align_corners=False) low-res logits -> T x T, crop to input_size, bilinear -> original_size
//       (model/walkgpt.py:749-790 with vision_tower_for_mask=True; segment_anything/modeling/sam.py:137-172)
//   A9  binary mask = logits > 0 and mask_score = sum(sigmoid(m) * [m>0]) / (sum([m>0]) + 1e-6)  (model/walkgpt.py:541,742)
//   A10 relative-depth head -- THIS REPO'S EXTENSION (no reference implementation exists; see oracle/path_a.py:depth_head).
// The last resize is fused with the threshold, the u8 mask write and the score reduction: each logit is written once
// (16-byte stores) and never re-read.
#include "internal.h"
#include "ptx.cuh"

namespace wg {
namespace {

// PyTorch upsample_bilinear2d source index (align_corners = False)
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float s = fmaxf(scale * (dst + 0.5f) - 0.5f, 0.f);
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = s - (float)i0;
}

// dst[n, y, x] = bilinear(src[n]) for y < dh, x < dw, where the virtual full output is (full_h, full_w) (crop = top-left).
// One thread owns VEC consecutive x of a BAND of rows: the column taps / weights are computed once per thread, the source
// (a few KB per mask) stays in L1, every logit is written exactly once with 16-byte streaming stores, and the score partial
// sums are reduced once per CTA (2 atomics per band instead of per row).  grid (x groups, bands, masks).
// SCORE: also write the u8 mask and accumulate (sum sigmoid*[>0], count [>0]) per mask.
constexpr int BIL_THREADS = 128;
template <int VEC, bool SCORE>
__global__ void __launch_bounds__(BIL_THREADS) bilinear_kernel(const float* __restrict__ src, int sh, int sw, float* __restrict__ dst, int dh, int dw,
                                                                float scale_y, float scale_x, uint8_t* __restrict__ mask, float* __restrict__ accum,
                                                                int rows_per_band) {
    const int n = blockIdx.z;
    const int ya = blockIdx.y * rows_per_band;
    const int yb = min(dh, ya + rows_per_band);
    const int xg = blockIdx.x * blockDim.x + threadIdx.x;  // group of VEC pixels
    const int x0 = xg * VEC;
    float ssum = 0.f, scnt = 0.f;
    const unsigned row_lanes = __ballot_sync(0xffffffffu, x0 < dw);  // lanes of this warp that own pixels (the row loop's votes)
    if (x0 < dw) {
        int xa[VEC], xb[VEC];
        float lx[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int x = x0 + v;
            src_index(x < dw ? x : dw - 1, scale_x, sw, xa[v], xb[v], lx[v]);
        }
        const float* sbase = src + (size_t)n * sh * sw;
        const bool full_vec = (VEC == 4) && (x0 + 3 < dw);
        // horizontally interpolated source rows y0 / y1, recomputed only when the source row pair changes (an up-sampling
        // band reuses each pair for several output rows): 4 * VEC L1 loads per CHANGE instead of per output row
        float h0[VEC], h1[VEC];
        int cur0 = -1, cur1 = -1;
        float* drow = dst + ((size_t)n * dh + ya) * dw + x0;
        uint8_t* mrow = mask + ((size_t)n * dh + ya) * dw + x0;  // only dereferenced when mask != nullptr
        for (int y = ya; y < yb; ++y) {
            int y0, y1;
            float ly;
            src_index(y, scale_y, sh, y0, y1, ly);
            if (y0 != cur0 || y1 != cur1) {
                if (y0 == cur1) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) h0[v] = h1[v];
                } else {
                    const float* r0 = sbase + (size_t)y0 * sw;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) h0[v] = (1.f - lx[v]) * __ldg(r0 + xa[v]) + lx[v] * __ldg(r0 + xb[v]);
                }
                if (y1 == y0) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) h1[v] = h0[v];
                } else {
                    const float* r1 = sbase + (size_t)y1 * sw;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) h1[v] = (1.f - lx[v]) * __ldg(r1 + xa[v]) + lx[v] * __ldg(r1 + xb[v]);
                }
                cur0 = y0;
                cur1 = y1;
            }
            float o[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) o[v] = (1.f - ly) * h0[v] + ly * h1[v];
            if (full_vec) {
                __stcs(reinterpret_cast<float4*>(drow), make_float4(o[0], o[1], o[2], o[3]));
                if (SCORE && mask) {
                    uchar4 m;
                    m.x = o[0] > 0.f; m.y = o[1] > 0.f; m.z = o[2] > 0.f; m.w = o[3] > 0.f;
                    __stcs(reinterpret_cast<uchar4*>(mrow), m);
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (x0 + v < dw) {
                        drow[v] = o[v];
                        if (SCORE && mask) mrow[v] = o[v] > 0.f;
                    }
            }
            drow += dw;  // one pointer bump per output row instead of a 64-bit multiply-add chain
            mrow += dw;
            if (SCORE && accum != nullptr) {
                // one warp-uniform test per row: real masks are mostly negative, so most warps (128 consecutive pixels of a row)
                // skip the sigmoids altogether; a warp that holds a positive pixel evaluates all of its lanes branch-free.
                // The scoring variant of this kernel is issue-bound, not write-bound (mask-only: 42 us for 192 masks, with the
                // IEEE division 1.0f / (1.0f + __expf(-o)): 92 us): the reciprocal is MUFU.RCP (1 ulp) like the exponential is
                // MUFU.EX2 -- far inside the 1e-5 the score is compared at; exp(-o) = inf gives 1 / inf = 0, no NaN.
                bool any_pos = false;
#pragma unroll
                for (int v = 0; v < VEC; ++v) any_pos = any_pos || (x0 + v < dw && o[v] > 0.f);
                if (__any_sync(row_lanes, any_pos)) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const bool pos = x0 + v < dw && o[v] > 0.f;
                        float sg;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + __expf(-o[v])));
                        ssum += pos ? sg : 0.f;
                        scnt += pos ? 1.f : 0.f;
                    }
                }
            }
        }
    }
    if (SCORE && accum != nullptr) {
        __shared__ float red[2][BIL_THREADS / 32];
        ssum = warp_sum(ssum);
        scnt = warp_sum(scnt);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
            red[0][warp] = ssum;
            red[1][warp] = scnt;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, b = 0.f;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
                a += red[0][i];
                b += red[1][i];
            }
            if (b > 0.f) {
                atomicAdd(accum + 2 * n, a);
                atomicAdd(accum + 2 * n + 1, b);
            }
        }
    }
}

// Eight pixels per thread, for row widths that are a multiple of 8 (every size the path produces): round 1's four-pixel kernel was
// ISSUE-bound, not write-bound (ncu, 192 masks of random logits: issue slots 80 % busy, ALU pipe 69 %, 2.6-3.0 TB/s) -- about 110
// instructions per thread and output row for four pixels, half of them per-row bookkeeping every thread repeated (source-row index
// arithmetic with float <-> int conversions, votes, pointer math).  Here the row table (source rows + weight) is computed once per
// CTA into shared memory, a thread amortises the rest over eight pixels (two 16-byte logit stores, one 8-byte mask store), the
// positive-pixel count comes from a population count of the mask bytes and the score sum from predicated adds.  Same arithmetic per
// pixel as bilinear_kernel (bit-identical logits).
constexpr int BIL8_THREADS = 64;
constexpr int BIL8_MAX_ROWS = 64;
template <bool SCORE>
__global__ void __launch_bounds__(BIL8_THREADS) bilinear8_kernel(const float* __restrict__ src, int sh, int sw, float* __restrict__ dst, int dh, int dw,
                                                                  float scale_y, float scale_x, uint8_t* __restrict__ mask, float* __restrict__ accum,
                                                                  int rows_per_band) {
    __shared__ int4 rowtab[BIL8_MAX_ROWS];  // {y0 * sw, y1 * sw, bits of ly, unused}
    const int n = blockIdx.z;
    const int ya = blockIdx.y * rows_per_band;
    const int yb = min(dh, ya + rows_per_band);
    if (threadIdx.x < yb - ya) {
        int y0, y1;
        float ly;
        src_index(ya + threadIdx.x, scale_y, sh, y0, y1, ly);
        rowtab[threadIdx.x] = make_int4(y0 * sw, y1 * sw, __float_as_int(ly), 0);
    }
    __syncthreads();
    const int x0 = (blockIdx.x * BIL8_THREADS + threadIdx.x) * 8;
    float ssum = 0.f;
    int scnt = 0;
    const unsigned row_lanes = __ballot_sync(0xffffffffu, x0 < dw);
    if (x0 < dw) {
        int xa[8], xb[8];
        float lx[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) src_index(x0 + v, scale_x, sw, xa[v], xb[v], lx[v]);
        const float* sbase = src + (size_t)n * sh * sw;
        float h0[8], h1[8];
        int cur0 = -1, cur1 = -1;
        float* drow = dst + ((size_t)n * dh + ya) * dw + x0;
        uint8_t* mrow = mask + ((size_t)n * dh + ya) * dw + x0;  // only dereferenced when mask != nullptr
        for (int y = ya; y < yb; ++y) {
            const int4 rt = rowtab[y - ya];
            const float ly = __int_as_float(rt.z);
            if (rt.x != cur0 || rt.y != cur1) {
                if (rt.x == cur1) {
#pragma unroll
                    for (int v = 0; v < 8; ++v) h0[v] = h1[v];
                } else {
                    const float* r0 = sbase + rt.x;
#pragma unroll
                    for (int v = 0; v < 8; ++v) h0[v] = (1.f - lx[v]) * __ldg(r0 + xa[v]) + lx[v] * __ldg(r0 + xb[v]);
                }
                if (rt.y == rt.x) {
#pragma unroll
                    for (int v = 0; v < 8; ++v) h1[v] = h0[v];
                } else {
                    const float* r1 = sbase + rt.y;
#pragma unroll
                    for (int v = 0; v < 8; ++v) h1[v] = (1.f - lx[v]) * __ldg(r1 + xa[v]) + lx[v] * __ldg(r1 + xb[v]);
                }
                cur0 = rt.x;
                cur1 = rt.y;
            }
            float o[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) o[v] = (1.f - ly) * h0[v] + ly * h1[v];
            __stcs(reinterpret_cast<float4*>(drow), make_float4(o[0], o[1], o[2], o[3]));
            __stcs(reinterpret_cast<float4*>(drow) + 1, make_float4(o[4], o[5], o[6], o[7]));
            drow += dw;
            if (SCORE) {
                uint32_t m0 = 0, m1 = 0;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    m0 |= (o[v] > 0.f ? 1u : 0u) << (8 * v);
                    m1 |= (o[4 + v] > 0.f ? 1u : 0u) << (8 * v);
                }
                if (mask) {
                    __stcs(reinterpret_cast<uint2*>(mrow), make_uint2(m0, m1));
                    mrow += dw;
                }
                if (accum != nullptr && __any_sync(row_lanes, (m0 | m1) != 0)) {
                    scnt += __popc(m0) + __popc(m1);
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        // sigmoid = 1 / (1 + 2^(-o log2 e)): MUFU.EX2 + MUFU.RCP; ex2.approx.ftz flushes 2^(< -126) to 0 -> 1 / 1
                        float e, sg;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(o[v] * -1.4426950408889634f));
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + e));
                        if (o[v] > 0.f) ssum += sg;
                    }
                }
            }
        }
    }
    if (SCORE && accum != nullptr) {
        __shared__ float red[2][BIL8_THREADS / 32];
        ssum = warp_sum(ssum);
        const float cnt = warp_sum((float)scnt);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
            red[0][warp] = ssum;
            red[1][warp] = cnt;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, b = 0.f;
            for (int i = 0; i < BIL8_THREADS / 32; ++i) {
                a += red[0][i];
                b += red[1][i];
            }
            if (b > 0.f) {
                atomicAdd(accum + 2 * n, a);
                atomicAdd(accum + 2 * n + 1, b);
            }
        }
    }
}

__global__ void finalize_score_kernel(const float* __restrict__ accum, float* __restrict__ score, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) score[i] = accum[2 * i] / (accum[2 * i + 1] + 1e-6f);
}

// depth extension: pooled [P][33] -> raw[p] = W2 . gelu(W1 . pooled_c + b1) + b2 ; min-max normalised inside each image
__global__ void __launch_bounds__(256) depth_head_kernel(const float* __restrict__ pooled, const int* __restrict__ seg_offsets, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                                         float* __restrict__ depth, int max_S) {
    extern __shared__ float raw[];  // [max_S]
    __shared__ float red[8];
    const int img = blockIdx.x;
    const int p0 = seg_offsets[img], p1 = seg_offsets[img + 1];
    if (p1 <= p0 || p0 < 0) return;
    const int S = min(p1 - p0, max_S);  // raw[] holds max_S entries: inconsistent offsets cannot overrun it
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int sidx = 0; sidx < S; ++sidx) {
        const float* pp = pooled + (size_t)(p0 + sidx) * 33;
        const float inv = 1.0f / (pp[32] + 1e-6f);
        float h = b1[tid];
#pragma unroll
        for (int c = 0; c < 32; ++c) h = fmaf(w1[tid * 32 + c], pp[c] * inv, h);
        float part = warp_sum(gelu_erf(h) * w2[tid]);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (tid == 0) {
            float t = b2[0];
            for (int i = 0; i < 8; ++i) t += red[i];
            raw[sidx] = t;
        }
        __syncthreads();
    }
    float lo = INFINITY, hi = -INFINITY;
    for (int i = 0; i < S; ++i) {
        lo = fminf(lo, raw[i]);
        hi = fmaxf(hi, raw[i]);
    }
    for (int i = tid; i < S; i += 256) depth[p0 + i] = (raw[i] - lo) / (hi - lo + 1e-6f);
}


// PositionEmbeddingRandom.forward((h,w)) (prompt_encoder.py:216-229): out[c, y, x], c < 2F: sin | cos of 2*pi*((2*coord-1) @ gauss)
__global__ void dense_pe_kernel(const float* __restrict__ gauss, int F, int h, int w, float* __restrict__ out_chw, float* __restrict__ out_tok) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * h * w) return;
    const int f = i / (h * w), pos = i % (h * w);
    const int y = pos / w, x = pos % w;
    const float cx = 2.f * ((x + 0.5f) / w) - 1.f, cy = 2.f * ((y + 0.5f) / h) - 1.f;
    const float a = 6.283185307179586f * (cx * gauss[f] + cy * gauss[F + f]);
    const float sv = sinf(a), cv = cosf(a);
    if (out_chw) {
        out_chw[(size_t)f * h * w + pos] = sv;
        out_chw[(size_t)(F + f) * h * w + pos] = cv;
    }
    if (out_tok) {
        out_tok[(size_t)pos * 2 * F + f] = sv;
        out_tok[(size_t)pos * 2 * F + F + f] = cv;
    }
}

template <bool SCORE>
int launch_bilinear(const float* src, int n, int sh, int sw, float* dst, int dh, int dw, int full_h, int full_w, uint8_t* mask, float* accum,
                    cudaStream_t s) {
    const float sy = (float)sh / (float)full_h, sx = (float)sw / (float)full_w;
    // algorithmic bytes: read the source once, write fp32 logits (+ u8 mask)
    Prof prof(SCORE ? "postprocess_bilinear_score" : "postprocess_bilinear", s, 0.0,
              (double)n * ((double)sh * sw * 4.0 + (double)dh * dw * (SCORE && mask ? 5.0 : 4.0)));
    // bands of rows: ~32 CTAs per SM, i.e. several waves at the 7-12 CTAs an SM holds (with ~9 per SM the scoring variant ran 1.3
    // waves: a second, 30 % full round), at least 8 rows each so the per-thread column set-up is amortised
    if (dw % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 7) == 0)) {
        const int groups8 = dw / 8;
        const int xb8 = (groups8 + BIL8_THREADS - 1) / BIL8_THREADS;
        int bands8 = (32 * device_sm_count() + n * xb8 - 1) / (n * xb8);
        if (bands8 < 1) bands8 = 1;
        int rpb = (dh + bands8 - 1) / bands8;
        if (rpb < 8) rpb = 8;
        if (rpb > BIL8_MAX_ROWS) rpb = BIL8_MAX_ROWS;
        bands8 = (dh + rpb - 1) / rpb;
        if (bands8 <= 65535) {
            bilinear8_kernel<SCORE><<<dim3(xb8, bands8, n), BIL8_THREADS, 0, s>>>(src, sh, sw, dst, dh, dw, sy, sx, mask, accum, rpb);
            WG_CHECK_CUDA(cudaGetLastError());
            return WG_OK;
        }
    }
    const int vec = (dw % 4 == 0) ? 4 : 1;
    const int groups = (dw + vec - 1) / vec;
    const int xblocks = (groups + BIL_THREADS - 1) / BIL_THREADS;
    int bands = (32 * device_sm_count() + n * xblocks - 1) / (n * xblocks);
    if (bands < 1) bands = 1;
    int rows_per_band = (dh + bands - 1) / bands;
    if (rows_per_band < 8) rows_per_band = 8;
    bands = (dh + rows_per_band - 1) / rows_per_band;
    dim3 grid(xblocks, bands, n);
    if (vec == 4)
        bilinear_kernel<4, SCORE><<<grid, BIL_THREADS, 0, s>>>(src, sh, sw, dst, dh, dw, sy, sx, mask, accum, rows_per_band);
    else
        bilinear_kernel<1, SCORE><<<grid, BIL_THREADS, 0, s>>>(src, sh, sw, dst, dh, dw, sy, sx, mask, accum, rows_per_band);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

}  // namespace
}  // namespace wg

using namespace wg;

extern "C" size_t wg_postprocess_workspace_bytes(int n_masks, int in_h, int in_w) {
    Workspace ws(nullptr, 0);
    ws.take((size_t)n_masks * 2 * sizeof(float));
    ws.take((size_t)n_masks * in_h * in_w * sizeof(float));
    return ws.used();
}

extern "C" int wg_postprocess_masks(const float* low_res, int n_masks, int Hm, int Wm, int target, int in_h, int in_w, int out_h, int out_w,
                                    float* logits_out, uint8_t* mask_out, float* score_out, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    if (n_masks == 0) return WG_OK;
    WG_REQUIRE(low_res && logits_out && workspace, "wg_postprocess_masks: null pointer");
    WG_REQUIRE(n_masks > 0 && n_masks <= 65535 && Hm > 0 && Wm > 0, "wg_postprocess_masks: bad sizes");
    WG_REQUIRE(in_h > 0 && in_w > 0 && in_h <= target && in_w <= target, "wg_postprocess_masks: input_size (%d,%d) must fit target %d", in_h, in_w,
               target);  // asserts at model/walkgpt.py:781-782
    WG_REQUIRE(out_h > 0 && out_w > 0 && out_h <= 65535 && in_h <= 65535, "wg_postprocess_masks: bad output size");
    if (!device_is_sm100()) {
        set_error("wg_postprocess_masks: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    Workspace ws(workspace, workspace_bytes);
    float* accum = static_cast<float*>(ws.take((size_t)n_masks * 2 * sizeof(float)));
    float* mid = static_cast<float*>(ws.take((size_t)n_masks * in_h * in_w * sizeof(float)));
    WG_REQUIRE(accum && mid, "wg_postprocess_masks: workspace too small");
    const bool want_score = score_out != nullptr || mask_out != nullptr;
    if (want_score) WG_CHECK_CUDA(cudaMemsetAsync(accum, 0, (size_t)n_masks * 2 * sizeof(float), s));
    if (in_h == out_h && in_w == out_w) {
        // the second interpolation is the identity (scale 1 => lambda 0): one fused pass
        if (want_score)
            WG_TRY(launch_bilinear<true>(low_res, n_masks, Hm, Wm, logits_out, in_h, in_w, target, target, mask_out, score_out ? accum : nullptr, s));
        else
            WG_TRY(launch_bilinear<false>(low_res, n_masks, Hm, Wm, logits_out, in_h, in_w, target, target, nullptr, nullptr, s));
    } else {
        WG_TRY(launch_bilinear<false>(low_res, n_masks, Hm, Wm, mid, in_h, in_w, target, target, nullptr, nullptr, s));
        if (want_score)
            WG_TRY(launch_bilinear<true>(mid, n_masks, in_h, in_w, logits_out, out_h, out_w, out_h, out_w, mask_out, score_out ? accum : nullptr, s));
        else
            WG_TRY(launch_bilinear<false>(mid, n_masks, in_h, in_w, logits_out, out_h, out_w, out_h, out_w, nullptr, nullptr, s));
    }
    if (score_out) {
        Prof prof("finalize_score", s);
        finalize_score_kernel<<<(n_masks + 255) / 256, 256, 0, s>>>(accum, score_out, n_masks);
        WG_CHECK_CUDA(cudaGetLastError());
    }
    return WG_OK;
}

extern "C" int wg_depth_head(const float* pooled, const int32_t* seg_offsets, int B, int max_S, const float* w1, const float* b1, const float* w2,
                             const float* b2, float* depth_out, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    if (B == 0) return WG_OK;
    WG_REQUIRE(pooled && seg_offsets && w1 && b1 && w2 && b2 && depth_out, "wg_depth_head: null pointer");
    WG_REQUIRE(max_S >= 0 && max_S <= 8192, "wg_depth_head: max_S=%d out of range", max_S);
    if (!device_is_sm100()) {
        set_error("wg_depth_head: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    Prof prof("depth_head", s);
    depth_head_kernel<<<B, 256, (size_t)(max_S > 0 ? max_S : 1) * sizeof(float), s>>>(pooled, seg_offsets, w1, b1, w2, b2, depth_out, max_S > 0 ? max_S : 1);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

extern "C" int wg_dense_pe(const float* gauss, int num_pos_feats, int h, int w, float* out_chw, float* out_tokens, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(gauss && (out_chw || out_tokens) && num_pos_feats > 0 && h > 0 && w > 0, "wg_dense_pe: bad arguments");
    if (!device_is_sm100()) {
        set_error("wg_dense_pe: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const int n = num_pos_feats * h * w;
    Prof prof("dense_pe", s);
    dense_pe_kernel<<<(n + 255) / 256, 256, 0, s>>>(gauss, num_pos_feats, h, w, out_chw, out_tokens);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}
