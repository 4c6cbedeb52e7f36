// A5 Calibrated Text Projector (utils/utils_walkgpt.py:321-327), A3 out_mm_projector MLP (llava_arch.py:38-42) and
// A4 image_feature_neck (model/walkgpt.py:97-113): GEMMs from gemm.cu plus the small row-wise tails defined here.
#include <stdlib.h>

#include "internal.h"
#include "ptx.cuh"

namespace wg {
namespace {

// CTP tail, one warp per row of 256: y = LN(h)*g+b + text_type;  y = y / max(||y||, 1e-12) * exp(log_temp)
template <typename TOut>
__global__ void __launch_bounds__(256) ctp_tail_kernel(const float* __restrict__ h, const float* __restrict__ g, const float* __restrict__ b,
                                                       const float* __restrict__ text_type, const float* __restrict__ log_temp,
                                                       TOut* __restrict__ out, int rows) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float v[8];
    const float* hr = h + (size_t)row * 256;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        float4 t = *reinterpret_cast<const float4*>(hr + (c * 32 + lane) * 4);
        v[c * 4 + 0] = t.x; v[c * 4 + 1] = t.y; v[c * 4 + 2] = t.z; v[c * 4 + 3] = t.w;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / 256);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) q = fmaf(v[i] - mean, v[i] - mean, q);
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / 256) + 1e-5f);
    float n2 = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int col = (c * 32 + lane) * 4 + e;
            float y = (v[c * 4 + e] - mean) * rstd * g[col] + b[col] + text_type[col];
            v[c * 4 + e] = y;
            n2 = fmaf(y, y, n2);
        }
    }
    const float nrm = fmaxf(sqrtf(warp_sum(n2)), 1e-12f);
    const float sc = expf(log_temp[0]) / nrm;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int col = (c * 32 + lane) * 4;
        if constexpr (sizeof(TOut) == 4) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)row * 256 + col) =
                make_float4(v[c * 4] * sc, v[c * 4 + 1] * sc, v[c * 4 + 2] * sc, v[c * 4 + 3] * sc);
        } else {
            uint2 u;
            u.x = pack_bf16x2(v[c * 4] * sc, v[c * 4 + 1] * sc);
            u.y = pack_bf16x2(v[c * 4 + 2] * sc, v[c * 4 + 3] * sc);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)row * 256 + col) = u;
        }
    }
}

// 3x3 / pad 1 im2col on a channels-last token grid: in [B, g, g, in_ld] (C channels starting at the pointer) bf16 ->
// out [B*g*g, out_ld] (9*C columns starting at the pointer), column = (ky*3+kx)*C + c
__global__ void __launch_bounds__(256) im2col3x3_kernel(const __nv_bfloat16* __restrict__ in, int in_ld, __nv_bfloat16* __restrict__ out, int out_ld,
                                                        int B, int g, int C) {
    const int vec_per_row = 9 * C / 8;
    const long long total = (long long)B * g * g * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec_per_row);
        const long long tok = i / vec_per_row;
        const int tap = v / (C / 8), c8 = v % (C / 8);
        const int ky = tap / 3, kx = tap % 3;
        const int b = (int)(tok / (g * g));
        const int pos = (int)(tok % (g * g));
        const int y = pos / g + ky - 1, x = pos % g + kx - 1;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (y >= 0 && y < g && x >= 0 && x < g)
            val = *reinterpret_cast<const uint4*>(in + (((size_t)b * g + y) * g + x) * in_ld + c8 * 8);
        *reinterpret_cast<uint4*>(out + tok * out_ld + v * 8) = val;
    }
}

// token-major [B, L, C] bf16 (or split-bf16 [B, L, 2C]: value = hi + lo) -> NCHW [B, C, L] (fp32 or bf16); 32x32 smem tile transpose
template <typename TOut>
__global__ void __launch_bounds__(256) tokens_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, TOut* __restrict__ out, int L, int C, int split) {
    const int ld = split ? 2 * C : C;
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        int l = l0 + j, c = c0 + tx;
        float v = 0.f;
        if (l < L && c < C) {
            const __nv_bfloat16* pr = in + ((size_t)b * L + l) * ld + c;
            v = __bfloat162float(pr[0]) + (split ? __bfloat162float(pr[C]) : 0.f);
        }
        tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int c = c0 + j, l = l0 + tx;
        if (c < C && l < L) out[((size_t)b * C + c) * L + l] = static_cast<TOut>(tile[tx][j]);
    }
}

}  // namespace

int launch_tokens_to_nchw(const void* tokens_bf16, int split, void* out, int out_is_bf16, int B, int L, int C, cudaStream_t s) {
    dim3 grid((L + 31) / 32, (C + 31) / 32, B);
    Prof prof("tokens_to_nchw", s, 0.0, (double)B * L * C * (2.0 + (out_is_bf16 ? 2.0 : 4.0)));
    if (out_is_bf16)
        tokens_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(tokens_bf16), static_cast<__nv_bfloat16*>(out), L, C, split);
    else
        tokens_to_nchw_kernel<float><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(tokens_bf16), static_cast<float*>(out), L, C, split);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

}  // namespace wg

using namespace wg;

extern "C" size_t wg_ctp_workspace_bytes(int rows, int in_dim) {
    Workspace ws(nullptr, 0);
    ws.take((size_t)rows * in_dim * 2);
    ws.take((size_t)rows * 512 * 2);
    ws.take((size_t)rows * 256 * 4);
    return ws.used();
}

extern "C" int wg_ctp_forward(const wg_ctp_weights* w, const void* x, int x_is_bf16, int rows, void* out, int out_is_bf16, void* workspace,
                              size_t workspace_bytes, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(w && out, "wg_ctp_forward: null pointer");
    if (rows == 0) return WG_OK;
    WG_REQUIRE(x && workspace, "wg_ctp_forward: null pointer");
    WG_REQUIRE(w->out_dim == 256 && w->mid_dim == 512, "wg_ctp_forward: built for mid=512/out=256 (got %d/%d)", w->mid_dim, w->out_dim);
    WG_REQUIRE(w->in_dim % 128 == 0, "wg_ctp_forward: in_dim %d must be a multiple of 128", w->in_dim);
    if (!device_is_sm100()) {
        set_error("wg_ctp_forward: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    Workspace ws(workspace, workspace_bytes);
    void* xn = ws.take((size_t)rows * w->in_dim * 2);
    void* h1 = ws.take((size_t)rows * 512 * 2);
    float* h2 = static_cast<float*>(ws.take((size_t)rows * 256 * 4));
    WG_REQUIRE(xn && h1 && h2, "wg_ctp_forward: workspace too small");
    WG_TRY(wg_layernorm(x, x_is_bf16, w->in_dim, w->ln0_g, w->ln0_b, 1e-5f, xn, w->in_dim, rows, w->in_dim, s));
    WG_TRY(gemm_bf16_out(xn, w->in_dim, w->w1, rows, 512, w->in_dim, w->b1, WG_ACT_GELU_ERF, h1, 512, s));
    WG_TRY(gemm_f32_out(h1, 512, w->w2, rows, 256, 512, w->b2, WG_ACT_NONE, h2, 256, nullptr, s));
    const unsigned grid = (unsigned)((rows + 7) / 8);
    Prof prof("ctp_tail", s, 0.0, (double)rows * 256 * (4.0 + (out_is_bf16 ? 2.0 : 4.0)));
    if (out_is_bf16)
        ctp_tail_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(h2, w->ln4_g, w->ln4_b, w->text_type, w->log_temp, static_cast<__nv_bfloat16*>(out), rows);
    else
        ctp_tail_kernel<float><<<grid, 256, 0, s>>>(h2, w->ln4_g, w->ln4_b, w->text_type, w->log_temp, static_cast<float*>(out), rows);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

// neck, computed at near-fp32 accuracy with split-bf16 operands (see wg_gemm_args): 1x1 conv (GEMM) + LayerNorm2d fused in
// the epilogue; 3x3 conv as im2col GEMM + LayerNorm2d epilogue.  pr: split-bf16 [rows, 2H]; emb: split-bf16 [rows, 512].
// implicit-GEMM 3x3 convolution needs whole map rows per 128-position GEMM tile (wg_gemm_args.conv_grid); WG_NECK_IM2COL=1 forces
// the explicit im2col matrix (A/B switch for tests)
static bool neck_conv_implicit(int g) {
    static const bool off = getenv("WG_NECK_IM2COL") != nullptr && atoi(getenv("WG_NECK_IM2COL")) != 0;
    return !off && g > 0 && 128 % g == 0 && (g * g) % 128 == 0;
}

static int neck_impl(const wg_proj_neck_weights* w, const void* pr, int B, int grid_side, void* emb_tokens_split, void* n1, void* col,
                     cudaStream_t s) {
    const int rows = B * grid_side * grid_side;
    const int H = w->hidden;
    const int T = w->split_terms;  // 2: W' = [W|W] (weights exactly bf16), 3: W' = [W_hi|W_hi|W_lo]
    {
        wg_gemm_args a = {};
        a.A = pr; a.lda = 2 * H; a.W = w->w_conv1; a.ldw = (long long)T * H; a.M = rows; a.N = 256; a.K = T * H;
        a.a_k_wrap = T == 3 ? 2 * H : 0;
        a.out_mode = WG_OUT_BF16_LN; a.split_out = 1; a.out = n1; a.ldo = 512; a.ln_gamma = w->ln1_g; a.ln_beta = w->ln1_b; a.ln_eps = 1e-6f;
        WG_TRY(wg_gemm(&a, s));
    }
    // 3x3 convolution: implicit GEMM (the A operand is the channels-last map itself, read through shifted 4-D TMA boxes) when
    // the grid allows whole rows per 128-position tile; small test grids go through an explicit im2col matrix
    const bool implicit = neck_conv_implicit(grid_side);
    if (!implicit) {
        const long long total = (long long)rows * (9 * 256 / 8);
        long long blocks = (total + 255) / 256;
        if (blocks > 148 * 32) blocks = 148 * 32;
        Prof prof("im2col3x3", s, 0.0, (double)rows * 256 * 2 * 20.0, 2);
        const __nv_bfloat16* nin = static_cast<const __nv_bfloat16*>(n1);
        __nv_bfloat16* cout_ = static_cast<__nv_bfloat16*>(col);
        im2col3x3_kernel<<<(unsigned)blocks, 256, 0, s>>>(nin, 512, cout_, 2 * 2304, B, grid_side, 256);              // hi
        im2col3x3_kernel<<<(unsigned)blocks, 256, 0, s>>>(nin + 256, 512, cout_ + 2304, 2 * 2304, B, grid_side, 256);  // lo
        WG_CHECK_CUDA(cudaGetLastError());
    }
    {
        wg_gemm_args a = {};
        a.W = w->w_conv3; a.ldw = (long long)T * 2304; a.M = rows; a.N = 256; a.K = T * 2304;
        a.a_k_wrap = T == 3 ? 2 * 2304 : 0;
        if (implicit) {
            a.A = n1; a.lda = 512; a.conv_grid = grid_side; a.conv_channels = 256;
        } else {
            a.A = col; a.lda = 2 * 2304;
        }
        a.out_mode = WG_OUT_BF16_LN; a.split_out = 1; a.out = emb_tokens_split; a.ldo = 512; a.ln_gamma = w->ln2_g; a.ln_beta = w->ln2_b;
        a.ln_eps = 1e-6f;
        WG_TRY(wg_gemm(&a, s));
    }
    return WG_OK;
}

extern "C" size_t wg_proj_neck_workspace_bytes(const wg_proj_neck_weights* w, int rows) {
    if (!w) return 0;
    Workspace ws(nullptr, 0);
    ws.take((size_t)rows * 2 * w->hidden * 2);
    ws.take((size_t)rows * 2 * w->hidden * 2);
    ws.take((size_t)rows * 512 * 2);
    ws.take((size_t)rows * 2 * 2304 * 2);
    return ws.used();
}

// feats bf16 [B*g*g, mm_hidden] -> proj bf16 [B*g*g, hidden] (optional output) -> image embedding tokens bf16 [B*g*g, 256]
extern "C" int wg_proj_neck_forward(const wg_proj_neck_weights* w, const void* feats_bf16, int B, int grid_side, void* proj_out_bf16,
                                    void* emb_tokens_bf16, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(w && feats_bf16 && workspace, "wg_proj_neck_forward: null pointer");
    WG_REQUIRE(B > 0 && grid_side > 0, "wg_proj_neck_forward: bad sizes");
    WG_REQUIRE(w->out_chans == 256, "wg_proj_neck_forward: out_chans must be 256");
    WG_REQUIRE(w->split_terms == 2 || w->split_terms == 3, "wg_proj_neck_forward: split_terms must be 2 or 3");
    WG_REQUIRE(w->hidden % 64 == 0, "wg_proj_neck_forward: hidden must be a multiple of 64");
    if (!device_is_sm100()) {
        set_error("wg_proj_neck_forward: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const int rows = B * grid_side * grid_side;
    const int H = w->hidden;
    Workspace ws(workspace, workspace_bytes);
    void* h1 = ws.take((size_t)rows * 2 * H * 2);
    void* pr = ws.take((size_t)rows * 2 * H * 2);   // split-bf16 projector output [rows, 2H]
    void* n1 = ws.take((size_t)rows * 512 * 2);
    void* col = ws.take((size_t)rows * 2 * 2304 * 2);
    WG_REQUIRE(h1 && pr && n1 && col, "wg_proj_neck_forward: workspace too small");
    if (proj_out_bf16) pr = proj_out_bf16;
    // out_mm_projector: Linear(mm_hidden -> 2H) + GELU(erf) + Linear(2H -> H)
    WG_TRY(gemm_bf16_out(feats_bf16, w->mm_hidden, w->w_fc1, rows, 2 * H, w->mm_hidden, w->b_fc1, WG_ACT_GELU_ERF, h1, 2 * H, s));
    {
        wg_gemm_args a = {};
        a.A = h1; a.lda = 2 * H; a.W = w->w_fc2; a.ldw = 2 * H; a.M = rows; a.N = H; a.K = 2 * H;
        a.bias = w->b_fc2; a.bias_period = 1; a.out_mode = WG_OUT_BF16; a.split_out = 1; a.out = pr; a.ldo = 2 * H;
        WG_TRY(wg_gemm(&a, s));
    }
    if (!emb_tokens_bf16) return WG_OK;
    return neck_impl(w, pr, B, grid_side, emb_tokens_bf16, n1, col, s);
}

extern "C" size_t wg_neck_workspace_bytes(int rows) {
    Workspace ws(nullptr, 0);
    ws.take((size_t)rows * 512 * 2);
    ws.take((size_t)rows * 2 * 2304 * 2);
    return ws.used();
}

// proj tokens bf16 [B*g*g, H] (channels-last view of the neck's NCHW input) -> image embedding tokens bf16 [B*g*g, 256]
extern "C" int wg_neck_forward(const wg_proj_neck_weights* w, const void* proj_tokens_bf16, int B, int grid_side, void* emb_tokens_bf16,
                               void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(w && proj_tokens_bf16 && emb_tokens_bf16 && workspace, "wg_neck_forward: null pointer");
    WG_REQUIRE(B > 0 && grid_side > 0 && w->out_chans == 256, "wg_neck_forward: bad sizes");
    if (!device_is_sm100()) {
        set_error("wg_neck_forward: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const int rows = B * grid_side * grid_side;
    Workspace ws(workspace, workspace_bytes);
    void* n1 = ws.take((size_t)rows * 512 * 2);
    void* col = ws.take((size_t)rows * 2 * 2304 * 2);
    WG_REQUIRE(n1 && col, "wg_neck_forward: workspace too small");
    return neck_impl(w, proj_tokens_bf16, B, grid_side, emb_tokens_bf16, n1, col, s);
}

extern "C" int wg_tokens_to_nchw(const void* tokens_bf16, int split, void* out, int out_is_bf16, int B, int L, int C, void* stream_) {
    WG_REQUIRE(tokens_bf16 && out && B > 0 && L > 0 && C > 0, "wg_tokens_to_nchw: bad arguments");
    if (!device_is_sm100()) {
        set_error("wg_tokens_to_nchw: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    return launch_tokens_to_nchw(tokens_bf16, split, out, out_is_bf16, B, L, C, static_cast<cudaStream_t>(stream_));
}
