// the SAM ViT image encoder (segment_anything/modeling/image_encoder.py:17-126 ImageEncoderViT, :128-194 Block,
// built by build_sam.py:15-22 as ViT-H: 1024^2 input, patch 16, dim 1280, 32 blocks of 16 heads x 80, 14 x 14 windows except the
// global blocks 7 / 15 / 23 / 31, 256-channel neck) as a fixed sequence of this library's kernels:
//   patch unfold -> tcgen05 GEMM whose per-position bias table is pos_embed + conv bias (fp32 residual stream, image-major rows)
//   per block:  LayerNorm (windowed blocks: written straight into the zero-padded, WINDOW-major token order, so the QKV GEMM's output
//               already is the partitioned tensor)  ->  QKV GEMM (columns re-ordered at load time into 64- and 16-wide head slices)
//               ->  sam_attention (decomposed relative position inside; un-partitions and crops by writing image-major rows)
//               ->  proj GEMM + fp32 residual in place  ->  LayerNorm  ->  fc1 GEMM + erf-GELU  ->  fc2 GEMM + fp32 residual
//   neck: fp32 -> split-bf16, then the same near-fp32 1x1 conv + LayerNorm2d + implicit 3x3 conv + LayerNorm2d as Path A's neck.
#include "internal.h"
#include "ptx.cuh"

namespace wg {
namespace {

constexpr int SE_G = 64, SE_WS = 14, SE_NW = 5;  // map side, window side, windows per axis (70 / 14)

// LayerNorm of image-major fp32 rows, written to window-major bf16 rows (pad tokens = 0: window_partition pads AFTER norm1,
// image_encoder.py:178-184, 250-270).  One warp per OUTPUT row.
template <int CH>
__global__ void __launch_bounds__(256) ln_to_windows_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, __nv_bfloat16* __restrict__ y, long long out_rows) {
    constexpr int D = CH * 128;
    const long long orow = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (orow >= out_rows) return;
    const int tok = (int)(orow % (SE_WS * SE_WS));
    const long long wl = orow / (SE_WS * SE_WS);
    const int win = (int)(wl % (SE_NW * SE_NW));
    const long long img = wl / (SE_NW * SE_NW);
    const int iy = (win / SE_NW) * SE_WS + tok / SE_WS, ix = (win % SE_NW) * SE_WS + tok % SE_WS;
    __nv_bfloat16* yr = y + orow * D;
    if (iy >= SE_G || ix >= SE_G) {
#pragma unroll
        for (int c = 0; c < CH; ++c) *reinterpret_cast<uint2*>(yr + (c * 32 + lane) * 4) = make_uint2(0u, 0u);
        return;
    }
    const float* xr = x + (img * (SE_G * SE_G) + iy * SE_G + ix) * D;
    float v[CH][4];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(xr + (c * 32 + lane) * 4);
        v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w;
        s += (t.x + t.y) + (t.z + t.w);
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float d = v[c][e] - mean;
            q = fmaf(d, d, q);
        }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int col = (c * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(gamma + col), b = *reinterpret_cast<const float4*>(beta + col);
        uint2 u;
        u.x = pack_bf16x2((v[c][0] - mean) * rstd * g.x + b.x, (v[c][1] - mean) * rstd * g.y + b.y);
        u.y = pack_bf16x2((v[c][2] - mean) * rstd * g.z + b.z, (v[c][3] - mean) * rstd * g.w + b.w);
        *reinterpret_cast<uint2*>(yr + col) = u;
    }
}

// fp32 [rows, D] -> split-bf16 [rows, 2D] (hi | lo): the storage format of the near-fp32 neck GEMMs
__global__ void __launch_bounds__(256) f32_to_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long rows, int D) {
    const long long n4 = rows * (D / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (D / 4);
        const int col = (int)(i % (D / 4)) * 4;
        const float4 t = *reinterpret_cast<const float4*>(x + row * D + col);
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(t.x, t.y), h1 = __floats2bfloat162_rn(t.z, t.w);
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        uint2 hi, lo;
        hi.x = *reinterpret_cast<const uint32_t*>(&h0);
        hi.y = *reinterpret_cast<const uint32_t*>(&h1);
        lo.x = pack_bf16x2(t.x - f0.x, t.y - f0.y);
        lo.y = pack_bf16x2(t.z - f1.x, t.w - f1.y);
        *reinterpret_cast<uint2*>(y + row * 2 * D + col) = hi;
        *reinterpret_cast<uint2*>(y + row * 2 * D + D + col) = lo;
    }
}

int launch_ln_windows(const float* x, const float* g, const float* b, float eps, void* y, long long out_rows, int D, cudaStream_t s) {
    Prof prof("sam_ln_to_windows", s, 0.0, (double)out_rows * D * 6.0 * (4096.0 / 4900.0));
    const unsigned grid = (unsigned)((out_rows + 7) / 8);
    __nv_bfloat16* yy = static_cast<__nv_bfloat16*>(y);
    switch (D / 128) {
        case 10: ln_to_windows_kernel<10><<<grid, 256, 0, s>>>(x, g, b, eps, yy, out_rows); break;  // ViT-H
        case 8: ln_to_windows_kernel<8><<<grid, 256, 0, s>>>(x, g, b, eps, yy, out_rows); break;    // ViT-L
        case 6: ln_to_windows_kernel<6><<<grid, 256, 0, s>>>(x, g, b, eps, yy, out_rows); break;    // ViT-B
        case 5: ln_to_windows_kernel<5><<<grid, 256, 0, s>>>(x, g, b, eps, yy, out_rows); break;    // test width (8 heads)
        default: set_error("wg_sam_encoder_forward: hidden %d is not built (640 / 768 / 1024 / 1280)", D); return WG_ERR_INVALID;
    }
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

struct SamBuffers {
    void* patches;  // bf16 [B*4096, 3*patch^2]
    float* x;       // fp32 [B*4096, D] residual stream (image-major)
    void* ln;       // bf16 [B*4900, D]  (window-major in windowed blocks)
    void* qkv;      // bf16 [B*4900, 3D]
    void* attn;     // bf16 [B*4096, D]
    void* h1;       // bf16 [B*4096, mlp]
    void* xs;       // split-bf16 [B*4096, 2D] neck input
    void* neck_ws;
    size_t neck_ws_bytes;
};

bool carve(Workspace& ws, const wg_sam_encoder_weights* w, int B, SamBuffers& b) {
    const size_t L = (size_t)SE_G * SE_G, LW = (size_t)SE_NW * SE_NW * SE_WS * SE_WS, D = w->hidden;
    b.patches = ws.take((size_t)B * L * 3 * w->patch * w->patch * 2);
    b.x = static_cast<float*>(ws.take((size_t)B * L * D * 4));
    b.ln = ws.take((size_t)B * LW * D * 2);
    b.qkv = ws.take((size_t)B * LW * 3 * D * 2);
    b.attn = ws.take((size_t)B * L * D * 2);
    b.h1 = ws.take((size_t)B * L * w->mlp * 2);
    b.xs = ws.take((size_t)B * L * 2 * D * 2);
    b.neck_ws_bytes = wg_neck_workspace_bytes((int)(B * L));
    b.neck_ws = ws.take(b.neck_ws_bytes);
    return b.patches && b.x && b.ln && b.qkv && b.attn && b.h1 && b.xs && b.neck_ws;
}

}  // namespace
}  // namespace wg

extern "C" size_t wg_sam_encoder_workspace_bytes(const wg_sam_encoder_weights* w, int B) {
    using namespace wg;
    if (!w || B <= 0) return 0;
    Workspace ws(nullptr, 0);
    SamBuffers b;
    carve(ws, w, B, b);
    return ws.used();
}

extern "C" int wg_sam_encoder_forward(const wg_sam_encoder_weights* w, const void* pixels, int pixels_is_bf16, int B, int n_run, void* emb_tokens_split,
                                      float* x_out, void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace wg;
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(w && pixels && workspace && (emb_tokens_split || x_out), "wg_sam_encoder_forward: null pointer");
    WG_REQUIRE(B > 0 && B <= 2621, "wg_sam_encoder_forward: 1..2621 images per call (got %d)", B);
    WG_REQUIRE(w->image == 1024 && w->patch == 16, "wg_sam_encoder_forward: built for 1024-pixel inputs and 16-pixel patches (64 x 64 map; got %d / %d)",
               w->image, w->patch);
    WG_REQUIRE(w->heads > 0 && w->hidden == w->heads * 80 && w->hidden % 128 == 0, "wg_sam_encoder_forward: hidden must be heads * 80 and a multiple of 128 (got %d / %d)",
               w->hidden, w->heads);
    WG_REQUIRE(w->mlp % 8 == 0 && w->out_chans == 256, "wg_sam_encoder_forward: bad mlp / out_chans");
    WG_REQUIRE(n_run >= 0 && n_run <= w->depth && w->blocks, "wg_sam_encoder_forward: n_run=%d exceeds the %d blocks supplied", n_run, w->depth);
    if (!device_is_sm100()) {
        set_error("wg_sam_encoder_forward: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    Workspace ws(workspace, workspace_bytes);
    SamBuffers b;
    WG_REQUIRE(carve(ws, w, B, b), "wg_sam_encoder_forward: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
               wg_sam_encoder_workspace_bytes(w, B));
    const int D = w->hidden, L = SE_G * SE_G, M = B * L, KP = 3 * w->patch * w->patch;
    const long long MW = (long long)B * SE_NW * SE_NW * SE_WS * SE_WS;

    // patch embedding (PatchEmbed, image_encoder.py:395-426) + pos_embed (:107-110): bias table [4096, D] = pos_embed + conv bias
    WG_TRY(launch_im2col_patch(pixels, pixels_is_bf16, b.patches, B, w->image, w->patch, KP, s));
    {
        wg_gemm_args a = {};
        a.A = b.patches; a.lda = KP; a.W = w->patch_w; a.ldw = KP; a.M = M; a.N = D; a.K = KP;
        a.bias = w->pos_bias; a.bias_period = L; a.out_mode = WG_OUT_F32; a.out = b.x; a.ldo = D;
        WG_TRY(wg_gemm(&a, s));
    }
    for (int i = 0; i < n_run; ++i) {
        const wg_sam_block& bk = w->blocks[i];
        const long long rows = bk.is_global ? M : MW;
        WG_REQUIRE(rows <= 0x7fffffffLL, "wg_sam_encoder_forward: too many rows");
        if (bk.is_global) WG_TRY(wg_layernorm(b.x, 0, D, bk.ln1_g, bk.ln1_b, 1e-6f, b.ln, D, M, D, s));
        else WG_TRY(launch_ln_windows(b.x, bk.ln1_g, bk.ln1_b, 1e-6f, b.ln, MW, D, s));
        WG_TRY(gemm_bf16_out(b.ln, D, bk.w_qkv, (int)rows, 3 * D, D, bk.b_qkv, WG_ACT_NONE, b.qkv, 3 * D, s));
        WG_TRY(wg_sam_attention(b.qkv, bk.rel_table, b.attn, B, bk.is_global ? 1 : 0, w->heads, s));
        WG_TRY(gemm_f32_out(b.attn, D, bk.w_proj, M, D, D, bk.b_proj, WG_ACT_NONE, b.x, D, b.x, s));
        WG_TRY(wg_layernorm(b.x, 0, D, bk.ln2_g, bk.ln2_b, 1e-6f, b.ln, D, M, D, s));
        WG_TRY(gemm_bf16_out(b.ln, D, bk.w_fc1, M, w->mlp, D, bk.b_fc1, WG_ACT_GELU_ERF, b.h1, w->mlp, s));
        WG_TRY(gemm_f32_out(b.h1, w->mlp, bk.w_fc2, M, D, w->mlp, bk.b_fc2, WG_ACT_NONE, b.x, D, b.x, s));
    }
    if (x_out) WG_CHECK_CUDA(cudaMemcpyAsync(x_out, b.x, (size_t)M * D * 4, cudaMemcpyDeviceToDevice, s));
    if (!emb_tokens_split) return WG_OK;
    {
        Prof prof("f32_to_split", s, 0.0, (double)M * D * 8.0);
        long long blocks = ((long long)M * (D / 4) + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        f32_to_split_kernel<<<(unsigned)blocks, 256, 0, s>>>(b.x, static_cast<__nv_bfloat16*>(b.xs), M, D);
        WG_CHECK_CUDA(cudaGetLastError());
    }
    return wg_neck_forward(&w->neck, b.xs, B, SE_G, emb_tokens_split, b.neck_ws, b.neck_ws_bytes, s);
}
