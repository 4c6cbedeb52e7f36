// Internal (non-ABI) launchers shared between translation units.
#pragma once
#include "host.h"

namespace wg {

int launch_im2col_patch(const void* pixels, int is_bf16, void* out_bf16, int B, int IMG, int P, int KPAD, cudaStream_t s);
int launch_embed_ln(const float* patch_emb, const float* cls, const float* pos, const float* gamma, const float* beta, float eps, float* x,
                    int B, int T, int D, cudaStream_t s);
int launch_drop_cls_cast(const float* x, void* out, int out_is_bf16, int B, int T, int D, cudaStream_t s, int keep_cls = 0);

// Bump allocator over a caller-provided workspace (256-byte aligned slices).
struct Workspace {
    uint8_t* base;
    size_t size;
    size_t off = 0;
    Workspace(void* p, size_t n) : base(static_cast<uint8_t*>(p)), size(n) {}
    void* take(size_t bytes) {
        size_t a = (off + 255) & ~size_t(255);
        if (base != nullptr && a + bytes > size) return nullptr;
        off = a + bytes;
        return base ? base + a : reinterpret_cast<void*>(uintptr_t(1));  // sizing pass: non-null dummy
    }
    size_t used() const { return (off + 255) & ~size_t(255); }
};

// convenience wrappers around wg_gemm
inline int gemm_bf16_out(const void* A, long long lda, const void* W, int M, int N, int K, const float* bias, int act, void* out, long long ldo,
                         cudaStream_t s) {
    wg_gemm_args a = {};
    a.A = A; a.lda = lda; a.W = W; a.ldw = K; a.M = M; a.N = N; a.K = K;
    a.bias = bias; a.bias_period = 1; a.act = act; a.out_mode = WG_OUT_BF16; a.out = out; a.ldo = ldo;
    return wg_gemm(&a, s);
}
inline int gemm_f32_out(const void* A, long long lda, const void* W, int M, int N, int K, const float* bias, int act, float* out, long long ldo,
                        const float* resid, cudaStream_t s) {
    wg_gemm_args a = {};
    a.A = A; a.lda = lda; a.W = W; a.ldw = K; a.M = M; a.N = N; a.K = K;
    a.bias = bias; a.bias_period = 1; a.act = act; a.out_mode = WG_OUT_F32; a.out = out; a.ldo = ldo; a.resid = resid;
    return wg_gemm(&a, s);
}

// Fused "fp32 residual add + LayerNorm over the full row" epilogue of the CTA-pair GEMM (gemm2.cu, WG_OUT_F32_LN): the fp32 result
// (out, in place on the residual stream) is also emitted as bf16 LayerNorm(out) * gamma + beta.
struct GemmLnFuse {
    void* ln_out;        // bf16 [M, N], ld = N
    const float* gamma;
    const float* beta;
    float eps;
    float* stats;        // scratch [2 * ceil(M / 256)][N / 256][128][2] floats (both CTAs of every pair tile publish)
    unsigned* flags;     // scratch [2 * ceil(M / 256)][N / 256], zeroed (stream-ordered) before the first launch that uses them
    unsigned epoch;      // > 0, unique per launch since the flags were zeroed
};
int launch_gemm_pair_ln(const wg_gemm_args* a, const GemmLnFuse* f, cudaStream_t stream);
inline bool gemm_pair_ln_eligible(int M, int N) {
    const int nt = N / 256, pairs = device_sm_count() / 2;
    return M > 0 && N % 256 == 0 && nt >= 2 && nt <= 8 && pairs >= 2 * nt;
}

}  // namespace wg
