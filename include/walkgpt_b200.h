/*
 * walkgpt_b200 -- C ABI of the B200-native (sm_100a) implementation of WalkGPT's pixel-grounding
 * forward path.  This is the drop-in boundary: plain pointers and sizes, no C++ or
 * torch types.  The reference (rafiibnsultan/WalkGPT) has no FFI of its own -- its seam is the set of
 * nn.Module attributes listed below -- so each module-level entry point cites the reference
 * `forward` it replaces, and INTEGRATION.md shows the Python (ctypes / torch.library) stub that binds it.
 *
 * Conventions
 *   - every function returns WG_OK (0) or a negative error code; the message is available from
 *     wg_last_error() (thread-local).  Nothing throws, nothing calls exit().
 *   - all data pointers are DEVICE pointers unless the name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream, never
 *     synchronise, never allocate device memory: the caller supplies outputs and workspaces.
 *   - bf16 = __nv_bfloat16 (2 bytes), row-major, innermost dimension contiguous unless a leading
 *     dimension (`ld*`, in elements) is given.
 *   - there is NO CPU fallback: on a non-sm_100 device every compute entry point returns
 *     WG_ERR_UNSUPPORTED.
 */
#ifndef WALKGPT_B200_H
#define WALKGPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WG_OK 0
#define WG_ERR_INVALID (-1)
#define WG_ERR_CUDA (-2)
#define WG_ERR_UNSUPPORTED (-3)

#define WG_ABI_VERSION 1

#if defined(__GNUC__)
#define WG_API __attribute__((visibility("default")))
#else
#define WG_API
#endif

WG_API int wg_version(void);
WG_API const char* wg_last_error(void);
/* WG_OK when `device` is an sm_100 (B200) GPU this library can run on. */
WG_API int wg_device_check(int device);

/* ------------------------------------------------------------------------------------------------
 * Kernel-level entry points (used by the module-level functions below; exported so that each kernel
 * can be parity-tested on its own).
 * ---------------------------------------------------------------------------------------------- */

/* activation applied in a GEMM epilogue */
#define WG_ACT_NONE 0
#define WG_ACT_QUICK_GELU 1 /* x*sigmoid(1.702x): HF CLIPMLP "quick_gelu" */
#define WG_ACT_GELU_ERF 2   /* nn.GELU() */
#define WG_ACT_RELU 3

/* GEMM output modes */
#define WG_OUT_BF16 0     /* out_bf16[M,N]  = act(A W^T + bias)                      (TMA store)            */
#define WG_OUT_F32 1      /* out_f32[M,N]   = act(A W^T + bias) (+ resid_f32[M,N] when resid != NULL)       */
#define WG_OUT_BF16_LN 2  /* out_bf16[M,N]  = LayerNorm_N(A W^T + bias + resid_bf16) * gamma + beta, N==256 */

typedef struct wg_gemm_args {
    const void* A;   /* bf16 [M,K], leading dimension lda */
    int64_t lda;
    const void* W;   /* bf16 [N,K] (torch nn.Linear weight layout), leading dimension ldw */
    int64_t ldw;
    int32_t M, N, K;
    const float* bias;    /* fp32; NULL = none.  bias[(row % bias_period) * N + col]; bias_period <= 1: bias[col] */
    int32_t bias_period;
    int32_t act;          /* WG_ACT_* */
    int32_t out_mode;     /* WG_OUT_* */
    void* out;            /* bf16 or fp32 [M,N], leading dimension ldo */
    int64_t ldo;
    const void* resid;    /* WG_OUT_F32: fp32 [M,N] ld=ldo (may alias out); WG_OUT_BF16_LN: bf16 [M,N] ld=ldo; or NULL */
    const float* ln_gamma; /* WG_OUT_BF16_LN only */
    const float* ln_beta;
    float ln_eps;
    int32_t reserved;
} wg_gemm_args;

/* tcgen05/TMEM + TMA bf16 GEMM with fused epilogue.  Requirements: K % 8 == 0, lda/ldw % 8 == 0,
 * 16-byte aligned base pointers; WG_OUT_BF16*: N % 8 == 0 and ldo % 8 == 0. */
WG_API int wg_gemm(const wg_gemm_args* args, void* stream);

/* y_bf16[rows, D] = LayerNorm(x)[rows, D] * gamma + beta ; x is fp32 (x_is_bf16 == 0) or bf16.
 * gamma/beta may be NULL (plain normalisation).  D % 128 == 0, D <= 8192. */
WG_API int wg_layernorm(const void* x, int x_is_bf16, int64_t ldx, const float* gamma, const float* beta, float eps,
                 void* y_bf16, int64_t ldy, int64_t rows, int D, void* stream);

/* Fused multi-head attention over packed QKV (HF CLIPAttention semantics, custom_clip.py:27-38 mask):
 *   qkv   bf16 [B, T, 3*heads*64]  (Q | K | V, head h at columns h*64)
 *   out   bf16 [B, T, heads*64]
 *   key_valid  uint8 [B, T] or NULL: keys with 0 receive an additive -inf (finfo.min in the reference)
 *   scale multiplies Q.K^T (head_dim^-0.5).   head_dim is fixed at 64. */
WG_API int wg_attention_d64(const void* qkv, void* out, const uint8_t* key_valid, int B, int T, int heads, float scale,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Module-level entry points: one per reference nn.Module.forward on the hot path.
 * Weights are passed as plain structs of device pointers.  bf16 matrices use the torch nn.Linear
 * layout [out_features, in_features]; vectors (biases, LayerNorm affine, embeddings) are fp32.
 * The Python host layer (walkgpt_b200/modules.py) builds these structs from the reference state_dict.
 * ---------------------------------------------------------------------------------------------- */

/* A1 -- CLIP ViT-L/14 tower.  Replaces CLIPVisionTower.forward
 * (model/llava_walkgpt/model/multimodal_encoder/clip_encoder.py:71-98) whose arithmetic is HF
 * CLIPVisionModel (embeddings -> pre_layrnorm -> N x CLIPEncoderLayer), with the key-padding mask of
 * custom_clip.py:27-38 / llava_arch.py:160-193. */
typedef struct wg_clip_layer {
    const float* ln1_g; const float* ln1_b;
    const void* w_qkv;  const float* b_qkv;   /* bf16 [3*hidden, hidden] = q_proj | k_proj | v_proj rows ; fp32 [3*hidden] */
    const void* w_o;    const float* b_o;     /* bf16 [hidden, hidden] */
    const float* ln2_g; const float* ln2_b;
    const void* w_fc1;  const float* b_fc1;   /* bf16 [mlp, hidden] */
    const void* w_fc2;  const float* b_fc2;   /* bf16 [hidden, mlp] */
} wg_clip_layer;

typedef struct wg_clip_weights {
    int32_t hidden, heads, mlp, image, patch; /* 1024, 16, 4096, 448, 14 */
    int32_t kpad;                             /* padded im2col width: 3*patch*patch rounded up to a multiple of 64 */
    int32_t n_layers;                         /* entries in `layers` */
    int32_t reserved;
    const void* patch_w;                      /* bf16 [hidden, kpad]: conv weight flattened (c,dy,dx), zero padded */
    const float* cls_emb;                     /* fp32 [hidden] */
    const float* pos_emb;                     /* fp32 [1 + (image/patch)^2, hidden] */
    const float* pre_ln_g; const float* pre_ln_b;
    const wg_clip_layer* layers;              /* HOST array of n_layers entries */
} wg_clip_weights;

WG_API size_t wg_clip_workspace_bytes(const wg_clip_weights* w, int B);

/* pixels: [B,3,image,image] fp32 or bf16.  key_valid: uint8 [B, 1+g*g] or NULL.
 * Runs `n_run` encoder layers (hidden_states[n_run]); out_last receives hidden_states[n_run][:,1:],
 * out_mid (nullable) receives hidden_states[mid_index][:,1:] (mid_index <= n_run).  Outputs are
 * [B, g*g, hidden] in bf16 or fp32 (out_is_bf16). */
WG_API int wg_clip_forward(const wg_clip_weights* w, const void* pixels, int pixels_is_bf16, const uint8_t* key_valid, int B,
                    int n_run, int mid_index, void* out_last, void* out_mid, int out_is_bf16, void* workspace,
                    size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WALKGPT_B200_H */
