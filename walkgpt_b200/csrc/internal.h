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

}  // namespace wg
