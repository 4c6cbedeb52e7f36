// A1: CLIP ViT-L/14@448 tower forward = a fixed sequence of the kernels in gemm.cu / attention.cu / rowwise.cu.
// Residual stream in fp32 (LayerNorm statistics, residual adds); GEMM operands and attention in bf16.
#include "internal.h"
#include <atomic>

namespace wg {
namespace {

struct ClipBuffers {
    void* patches;   // bf16 [B*L, kpad]
    float* patch_emb;  // fp32 [B*L, D]
    float* x;        // fp32 [B*T, D]  residual stream
    void* ln;        // bf16 [B*T, D]
    void* qkv;       // bf16 [B*T, 3D]
    void* attn;      // bf16 [B*T, D]
    void* h1;        // bf16 [B*T, mlp]
    float* ln_stats;     // fused residual + LayerNorm epilogue (gemm2.cu, WG_OUT_F32_LN): per-tile row statistics
    unsigned* ln_flags;  // and the tiles' "published" flags (zeroed at the start of every forward)
    size_t ln_flag_bytes;
};

bool carve(Workspace& ws, const wg_clip_weights* w, int B, ClipBuffers& b) {
    const size_t g = w->image / w->patch, L = g * g, T = L + 1, D = w->hidden;
    b.patches = ws.take((size_t)B * L * w->kpad * 2);
    b.patch_emb = static_cast<float*>(ws.take((size_t)B * L * D * 4));
    b.x = static_cast<float*>(ws.take((size_t)B * T * D * 4));
    b.ln = ws.take((size_t)B * T * D * 2);
    b.qkv = ws.take((size_t)B * T * 3 * D * 2);
    b.attn = ws.take((size_t)B * T * D * 2);
    b.h1 = ws.take((size_t)B * T * w->mlp * 2);
    const size_t row_blocks = 2 * (((size_t)B * T + 255) / 256), nt = D / 256;  // both CTAs of the last pair tile publish, valid rows or not
    b.ln_stats = static_cast<float*>(ws.take(row_blocks * nt * 128 * 2 * sizeof(float)));
    b.ln_flag_bytes = row_blocks * nt * sizeof(unsigned);
    b.ln_flags = static_cast<unsigned*>(ws.take(b.ln_flag_bytes));
    return b.patches && b.patch_emb && b.x && b.ln && b.qkv && b.attn && b.h1 && b.ln_stats && b.ln_flags;
}

std::atomic<int>& fuse_ln_mode() {
    static std::atomic<int> mode([] { const char* e = getenv("WG_CLIP_FUSE_LN"); const int v = e ? atoi(e) : 0; return v < 0 ? 0 : v > 2 ? 2 : v; }());
    return mode;
}

}  // namespace
}  // namespace wg

extern "C" int wg_clip_set_fuse_ln(int mode) {
    std::atomic<int>& m = wg::fuse_ln_mode();
    if (mode < 0) return m.load();
    return m.exchange(mode > 2 ? 2 : mode);
}

extern "C" size_t wg_clip_workspace_bytes(const wg_clip_weights* w, int B) {
    using namespace wg;
    if (!w || B <= 0) return 0;
    Workspace ws(nullptr, 0);
    ClipBuffers b;
    carve(ws, w, B, b);
    return ws.used();
}

extern "C" int wg_clip_forward(const wg_clip_weights* w, const void* pixels, int pixels_is_bf16, const uint8_t* key_valid, int B, int n_run,
                               int mid_index, void* out_last, void* out_mid, int out_is_bf16, void* workspace, size_t workspace_bytes,
                               void* stream_) {
    return wg_clip_forward_ex(w, pixels, pixels_is_bf16, key_valid, B, n_run, mid_index, out_last, out_mid, out_is_bf16, 0, workspace, workspace_bytes,
                              stream_);
}

extern "C" int wg_clip_forward_ex(const wg_clip_weights* w, const void* pixels, int pixels_is_bf16, const uint8_t* key_valid, int B, int n_run,
                                  int mid_index, void* out_last, void* out_mid, int out_is_bf16, int keep_cls, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
    using namespace wg;
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(w && pixels && out_last && workspace, "wg_clip_forward: null pointer");
    WG_REQUIRE(B > 0, "wg_clip_forward: B=%d", B);
    WG_REQUIRE(w->hidden == 1024 && w->heads * 64 == w->hidden, "wg_clip_forward: only hidden=1024 / head_dim=64 is built (got %d/%d)",
               w->hidden, w->heads);
    WG_REQUIRE(w->image % w->patch == 0 && w->kpad % 64 == 0 && w->kpad >= 3 * w->patch * w->patch, "wg_clip_forward: bad patch geometry");
    WG_REQUIRE(n_run >= 0 && n_run <= w->n_layers, "wg_clip_forward: n_run=%d exceeds the %d layers supplied", n_run, w->n_layers);
    WG_REQUIRE(out_mid == nullptr || (mid_index >= 0 && mid_index <= n_run), "wg_clip_forward: mid_index=%d out of range", mid_index);
    if (!device_is_sm100()) {
        set_error("wg_clip_forward: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const int g = w->image / w->patch, L = g * g, T = L + 1, D = w->hidden;
    Workspace ws(workspace, workspace_bytes);
    ClipBuffers b;
    WG_REQUIRE(carve(ws, w, B, b), "wg_clip_forward: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
               wg_clip_workspace_bytes(w, B));
    const int M = B * T;
    const float scale = 0.125f;  // head_dim^-0.5, head_dim = 64

    // embeddings: conv-as-GEMM, + CLS, + position embedding, pre_layrnorm
    WG_TRY(launch_im2col_patch(pixels, pixels_is_bf16, b.patches, B, w->image, w->patch, w->kpad, s));
    WG_TRY(gemm_f32_out(b.patches, w->kpad, w->patch_w, B * L, D, w->kpad, nullptr, WG_ACT_NONE, b.patch_emb, D, nullptr, s));
    WG_TRY(launch_embed_ln(b.patch_emb, w->cls_emb, w->pos_emb, w->pre_ln_g, w->pre_ln_b, 1e-5f, b.x, B, T, D, s));
    if (out_mid && mid_index == 0) WG_TRY(launch_drop_cls_cast(b.x, out_mid, out_is_bf16, B, T, D, s, keep_cls));

    // LayerNorm fused into the epilogue of the GEMM that produces its input (WG_CLIP_FUSE_LN: 0 = separate LayerNorm kernels, 1 = the
    // fc2 GEMM also emits ln1 of the NEXT layer, 2 = and the out-proj GEMM emits ln2); needs a full wave of pair tiles
    const int fuse = gemm_pair_ln_eligible(M, D) ? fuse_ln_mode().load(std::memory_order_relaxed) : 0;
    unsigned epoch = 0;
    if (fuse) WG_CHECK_CUDA(cudaMemsetAsync(b.ln_flags, 0, b.ln_flag_bytes, s));
    auto gemm_resid_ln = [&](const void* A, long long lda, const void* W, int K, const float* bias, const float* g, const float* be) -> int {
        wg_gemm_args a = {};
        a.A = A; a.lda = lda; a.W = W; a.ldw = K; a.M = M; a.N = D; a.K = K;
        a.bias = bias; a.bias_period = 1; a.act = WG_ACT_NONE; a.out_mode = WG_OUT_F32; a.out = b.x; a.ldo = D; a.resid = b.x;
        GemmLnFuse f = {b.ln, g, be, 1e-5f, b.ln_stats, b.ln_flags, ++epoch};
        return launch_gemm_pair_ln(&a, &f, s);
    };
    for (int i = 0; i < n_run; ++i) {
        const wg_clip_layer& ly = w->layers[i];
        if (i == 0 || fuse < 1) WG_TRY(wg_layernorm(b.x, 0, D, ly.ln1_g, ly.ln1_b, 1e-5f, b.ln, D, M, D, s));
        WG_TRY(gemm_bf16_out(b.ln, D, ly.w_qkv, M, 3 * D, D, ly.b_qkv, WG_ACT_NONE, b.qkv, 3 * D, s));
        WG_TRY(wg_attention_d64(b.qkv, b.attn, key_valid, B, T, w->heads, scale, s));
        if (fuse >= 2) {
            WG_TRY(gemm_resid_ln(b.attn, D, ly.w_o, D, ly.b_o, ly.ln2_g, ly.ln2_b));
        } else {
            WG_TRY(gemm_f32_out(b.attn, D, ly.w_o, M, D, D, ly.b_o, WG_ACT_NONE, b.x, D, b.x, s));
            WG_TRY(wg_layernorm(b.x, 0, D, ly.ln2_g, ly.ln2_b, 1e-5f, b.ln, D, M, D, s));
        }
        WG_TRY(gemm_bf16_out(b.ln, D, ly.w_fc1, M, w->mlp, D, ly.b_fc1, WG_ACT_QUICK_GELU, b.h1, w->mlp, s));
        if (fuse >= 1 && i + 1 < n_run) {
            WG_TRY(gemm_resid_ln(b.h1, w->mlp, ly.w_fc2, w->mlp, ly.b_fc2, w->layers[i + 1].ln1_g, w->layers[i + 1].ln1_b));
        } else {
            WG_TRY(gemm_f32_out(b.h1, w->mlp, ly.w_fc2, M, D, w->mlp, ly.b_fc2, WG_ACT_NONE, b.x, D, b.x, s));
        }
        if (out_mid && mid_index == i + 1) WG_TRY(launch_drop_cls_cast(b.x, out_mid, out_is_bf16, B, T, D, s, keep_cls));
    }
    WG_TRY(launch_drop_cls_cast(b.x, out_last, out_is_bf16, B, T, D, s, keep_cls));
    return WG_OK;
}
