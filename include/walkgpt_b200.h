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

/* Launch accounting and an optional CUDA-event profiler (bench.py's roofline leg).
 * wg_launch_count: kernels launched by this library since the last reset.
 * wg_profile_enable(1): every launch is bracketed by CUDA events on its stream; wg_profile_collect synchronises them
 * and writes "<kernel-family> <launches> <total_ms> <algorithmic_flops> <algorithmic_bytes>" lines into `out`. */
WG_API long long wg_launch_count(int reset);
WG_API int wg_profile_enable(int on);
WG_API long long wg_profile_collect(char* out, long long cap);

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
    /* "split-bf16" = an fp32-accurate matrix X stored as bf16 [rows, 2C]: hi = bf16(X) in columns [0,C), lo = bf16(X - hi) in
     * [C,2C).  Used where the path needs near-fp32 accuracy at negligible cost (neck, mask decoder).
     * a_k_wrap = 2C: A is split-bf16 and K = 3C walks [hi | lo | hi] against W' = [W_hi | W_hi | W_lo] (0 = plain A).
     * split_out != 0: bf16 output modes write split-bf16 [M, 2N] (ldo >= 2N); in WG_OUT_BF16_LN `resid` is then split-bf16 too. */
    int32_t a_k_wrap;
    int32_t split_out;
    /* Implicit 3x3 / pad 1 convolution (conv_grid = g > 0): A is a channels-last map bf16 [M / g^2 images, g, g, lda] and the
     * GEMM's k index runs over (term, tap = ky*3+kx, channel < conv_channels): k-block (term, tap, c) of output positions
     * m0..m0+127 is ONE 4-D TMA box of the map shifted by (kx-1, ky-1), out-of-range neighbours zero-filled by the TMA unit --
     * no im2col matrix is ever written.  term selects the hi / lo half of a split-bf16 map (channel offset term * conv_channels,
     * after the a_k_wrap fold, which then counts in units of 9 * conv_channels).  Needs 128 % g == 0, g^2 % 128 == 0,
     * conv_channels % 64 == 0, K % (9 * conv_channels) == 0; 1-CTA kernels only. */
    int32_t conv_grid;
    int32_t conv_channels;
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
/* The same with feature_select's "cls_patch" (clip_encoder.py:61-69): keep_cls != 0 keeps the CLS row, outputs are then
 * [B, 1 + g*g, hidden]; keep_cls == 0 is wg_clip_forward ("patch"). */
WG_API int wg_clip_forward_ex(const wg_clip_weights* w, const void* pixels, int pixels_is_bf16, const uint8_t* key_valid, int B,
                       int n_run, int mid_index, void* out_last, void* out_mid, int out_is_bf16, int keep_cls, void* workspace,
                       size_t workspace_bytes, void* stream);

/* LayerNorm fused into the GEMM that produces its input (CTA-pair kernel, fp32 residual epilogue): mode 0 = separate LayerNorm
 * kernels (default; the faster choice under this pool's power cap), 1 = every fc2 GEMM also emits layer_norm1 of
 * the NEXT encoder layer, 2 = and every out-proj GEMM emits layer_norm2.  The tiles of a 1024-wide row block exchange their row
 * statistics through global memory inside the persistent kernel.  Returns the previous mode; mode < 0 only queries.  The initial
 * mode is the environment variable WG_CLIP_FUSE_LN (unset = 0).  Process-wide. */
WG_API int wg_clip_set_fuse_ln(int mode);

/* A2 -- Multi-Scale Query Projector.  Replaces MultiScaleQFormerProjector.forward (utils/utils_walkgpt.py:259-300). */
typedef struct wg_msqp_block {                 /* one CrossAttnBlock (utils_walkgpt.py:163-185) */
    const float* qn_g; const float* qn_b;      /* q_norm */
    const void* w_q;   const float* b_q;       /* bf16 [d,d] = attn.in_proj_weight rows [0,d) */
    const void* w_o;   const float* b_o;       /* attn.out_proj */
    const float* ffn_ln_g; const float* ffn_ln_b;
    const void* w_f1;  const float* b_f1;      /* bf16 [4d,d] */
    const void* w_f2;  const float* b_f2;      /* bf16 [d,4d] */
} wg_msqp_block;

typedef struct wg_msqp_scale {
    const float* queries;                      /* fp32 [nq, d] learned queries (q_x1 / q_x2 / q_x4 / q_global) */
    int32_t nq; int32_t reserved;
    /* bf16 [4d, d] = (K_layer0 | V_layer0 | K_layer1 | V_layer1) rows of the two blocks' in_proj, with each block's
       kv_norm affine folded in:  W' = W diag(gamma),  b' = b + W beta   (the kernel normalises kv without affine). */
    const void* w_kv;  const float* b_kv;
    wg_msqp_block blocks[2];
} wg_msqp_scale;

typedef struct wg_msqp_weights {
    int32_t sam_dim, llama_dim, d, heads;      /* d = 1024, heads = 8 */
    int32_t n_tokens;                          /* s*s after padding (36 when target_square_side = 6) */
    int32_t reserved;
    const void* w_in;  const float* b_in;      /* sam_to_proj: bf16 [d, sam_dim] */
    const float* gate_ln_g; const float* gate_ln_b;   /* gate.net.0 */
    const void* w_g1;  const float* b_g1;      /* gate.net.1: bf16 [128, d] */
    const float* w_g2; const float* b_g2;      /* gate.net.3: fp32 [128], [1] */
    const float* pad_token;                    /* fp32 [d] */
    const void* w_out; const float* b_out;     /* to_llama: bf16 [llama_dim, d] */
    wg_msqp_scale scales[4];                   /* x1, x2, x4, global */
} wg_msqp_weights;

WG_API size_t wg_msqp_workspace_bytes(const wg_msqp_weights* w, int B, int L);
/* feats bf16 [B, L, sam_dim] (L a perfect square) -> out [B, n_tokens, llama_dim] (bf16 or fp32). */
WG_API int wg_msqp_forward(const wg_msqp_weights* w, const void* feats_bf16, int B, int L, void* out, int out_is_bf16,
                           void* workspace, size_t workspace_bytes, void* stream);

/* A5 -- Calibrated Text Projector.  Replaces CalibratedTextProjector.forward (utils/utils_walkgpt.py:321-327),
 * widen = 2, use_residual = False (model/walkgpt.py:141). */
typedef struct wg_ctp_weights {
    int32_t in_dim, mid_dim, out_dim, reserved; /* H, 512, 256 */
    const float* ln0_g; const float* ln0_b;     /* net.0 */
    const void* w1; const float* b1;            /* net.1: bf16 [mid, in] */
    const void* w2; const float* b2;            /* net.3: bf16 [out, mid] */
    const float* ln4_g; const float* ln4_b;     /* net.4 */
    const float* text_type;                     /* fp32 [out] */
    const float* log_temp;                      /* fp32 [1] */
} wg_ctp_weights;

WG_API size_t wg_ctp_workspace_bytes(int rows, int in_dim);
/* x [rows, in_dim] fp32 or bf16 -> out [rows, 256] fp32 or bf16 (row-wise; gather-then-project == project-then-gather). */
WG_API int wg_ctp_forward(const wg_ctp_weights* w, const void* x, int x_is_bf16, int rows, void* out, int out_is_bf16,
                          void* workspace, size_t workspace_bytes, void* stream);

/* A3 + A4 -- out_mm_projector MLP (llava_arch.py:38-42, used at :197-208) followed by image_feature_neck
 * (model/walkgpt.py:97-113, LayerNorm2d = segment_anything/modeling/common.py:31-43). */
typedef struct wg_proj_neck_weights {
    int32_t mm_hidden, hidden, out_chans;  /* 1024, H, 256 */
    int32_t split_terms;                   /* neck GEMMs run on split-bf16 operands: 2 = W' is [W|W] (weights exactly bf16), 3 = [W_hi|W_hi|W_lo] */
    const void* w_fc1; const float* b_fc1;      /* bf16 [2H, mm_hidden] */
    const void* w_fc2; const float* b_fc2;      /* bf16 [H, 2H] */
    const void* w_conv1;                        /* bf16 [256, split_terms*H]   (1x1 conv, no bias), see wg_gemm_args.a_k_wrap */
    const float* ln1_g; const float* ln1_b;
    const void* w_conv3;                        /* bf16 [256, split_terms*2304]: per term, column (ky*3+kx)*256 + ci  (3x3 conv, pad 1, no bias) */
    const float* ln2_g; const float* ln2_b;
} wg_proj_neck_weights;

WG_API size_t wg_proj_neck_workspace_bytes(const wg_proj_neck_weights* w, int rows);
/* feats bf16 [B*g*g, mm_hidden] -> proj_out split-bf16 [B*g*g, 2H] (hi | lo; nullable) and the image embedding as TOKENS
 * emb_tokens split-bf16 [B*g*g, 512] (channels-last, hi | lo; nullable = projector only). */
WG_API int wg_proj_neck_forward(const wg_proj_neck_weights* w, const void* feats_bf16, int B, int grid_side,
                                void* proj_out_bf16, void* emb_tokens_bf16, void* workspace, size_t workspace_bytes,
                                void* stream);
WG_API size_t wg_neck_workspace_bytes(int rows);
/* image_feature_neck alone: proj tokens split-bf16 [B*g*g, 2H] (channels-last view of the NCHW input) -> emb tokens split-bf16 [B*g*g, 512]. */
WG_API int wg_neck_forward(const wg_proj_neck_weights* w, const void* proj_tokens_bf16, int B, int grid_side,
                           void* emb_tokens_bf16, void* workspace, size_t workspace_bytes, void* stream);
/* tokens bf16 [B, L, C] (split != 0: split-bf16 [B, L, 2C]) -> NCHW [B, C, L] (fp32 or bf16): the layout the reference modules exchange. */
WG_API int wg_tokens_to_nchw(const void* tokens_bf16, int split, void* out, int out_is_bf16, int B, int L, int C, void* stream);

/* A6 + A7 -- PromptEncoder (text_embeds path, prompt_encoder.py:140-186) + MaskDecoderMultiScale.forward
 * (mask_decoder_multi_scale.py:87-213, level_num = 0) with its TwoWayTransformer (transformer.py:62-242),
 * batched over all P prompts of a batch of images.
 * "_t" matrices are TRANSPOSED fp32 [in_features][out_features] (token-side linears, read column-wise by CUDA cores);
 * the image-side GEMM weights are bf16 [out_features][split_terms * in_features] = [W_hi | W_hi | W_lo] (split_terms = 3) or
 * [W | W] (split_terms = 2, weights exactly representable in bf16): the image-side activations are split-bf16 (hi | lo),
 * see wg_gemm_args.a_k_wrap, which gives this stage near-fp32 accuracy on the bf16 tensor cores. */
typedef struct wg_twoway_layer {
    const void* sa_wq_t; const float* sa_bq; const void* sa_wk_t; const float* sa_bk;   /* self_attn: [256][256] */
    const void* sa_wv_t; const float* sa_bv; const void* sa_wo_t; const float* sa_bo;
    const float* n1_g; const float* n1_b;
    const void* t2i_wq_t; const float* t2i_bq;      /* cross_attn_token_to_image.q_proj  [256][128] */
    const void* t2i_wo_t; const float* t2i_bo;      /* cross_attn_token_to_image.out_proj [128][256] */
    const float* n2_g; const float* n2_b;
    const void* mlp_w1_t; const float* mlp_b1;      /* [256][2048] */
    const void* mlp_w2_t; const float* mlp_b2;      /* [2048][256] */
    const float* n3_g; const float* n3_b;
    const void* i2t_wk_t; const float* i2t_bk;      /* cross_attn_image_to_token.k_proj [256][128] */
    const void* i2t_wv_t; const float* i2t_bv;      /* cross_attn_image_to_token.v_proj [256][128] */
    const void* w_img;                              /* bf16 [384][T*256]: rows = t2i.k_proj | t2i.v_proj | i2t.q_proj */
    const float* b_img;                             /* fp32 [hw][384] = (pe Wk^T + bk | bv | pe Wq^T + bq): dense PE folded in */
    const void* i2t_wo; const float* i2t_bo;        /* cross_attn_image_to_token.out_proj: bf16 [256][T*128] */
    const float* n4_g; const float* n4_b;
    /* optional: the token MLP (71 % of the token side's weight bytes) as two tensor-core GEMMs over the token rows of ALL prompts
     * instead of inside the one-CTA-per-prompt token kernel: bf16 [2048][T*256] and [256][T*2048] in the [W_hi | W_hi | W_lo]
     * layout of wg_gemm_args.a_k_wrap.  NULL = the token kernel multiplies with mlp_w1_t / mlp_w2_t itself. */
    const void* mlp_w1_split; const void* mlp_w2_split;
} wg_twoway_layer;

typedef struct wg_mask_decoder_weights {
    int32_t grid_h, grid_w, n_mask_tokens, up_stages;   /* 32, 32, 4, 1 (MaskDecoderMultiScale) ; 64, 64, 4, 2 (SAM MaskDecoder, Path B) */
    int32_t split_terms;                                /* T = 2 or 3 (see above) */
    int32_t multimask_first;                            /* first mask returned with multimask_output: 0 (multi-scale decoder,
                                                           mask_decoder_multi_scale.py:126-132) or 1 (SAM, mask_decoder.py:106-111) */
    const float* out_tokens;     /* fp32 [1 + n_mask_tokens][256] = (iou_token ; mask_tokens) + level_embed[0] */
    const float* sparse_add;     /* fp32 [256] added to each text embedding (level_embed[0]); NULL = none */
    const float* no_mask;        /* fp32 [256] PromptEncoder.no_mask_embed (dense prompt embedding) */
    wg_twoway_layer layers[2];
    const void* fin_wq_t; const float* fin_bq;      /* final_attn_token_to_image.q_proj [256][128] */
    const void* fin_wo_t; const float* fin_bo;      /* final_attn_token_to_image.out_proj [128][256] */
    const float* nf_g; const float* nf_b;           /* norm_final_attn */
    const void* w_img_fin;                          /* bf16 [256][T*256]: rows = final.k_proj | final.v_proj */
    const float* b_img_fin;                         /* fp32 [hw][256] = (pe Wk^T + bk | bv) */
    const void* w_up;                               /* bf16 [4*Cu][T*256]: row (dy*2+dx)*Cu + co = ConvT.weight[ci, co, dy, dx]; Cu = 32
                                                       (up_stages 1) or 64 (up_stages 2) */
    const float* b_up;                              /* fp32 [4*Cu] = ConvT.bias tiled over the 4 sub-pixels */
    const float* up_ln_g; const float* up_ln_b;     /* output_upscaling.1 (LayerNorm2d over Cu channels) */
    const void* w_up2;                              /* up_stages 2: bf16 [4*32][T*64] second ConvTranspose (64 -> 32), same row order */
    const float* b_up2;                             /* up_stages 2: fp32 [128] */
    const void* hyp_w0_t; const float* hyp_b0;      /* hypernetwork MLPs: fp32 [4][256][256], fp32 [4][256] */
    const void* hyp_w1_t; const float* hyp_b1;      /* [4][256][256] */
    const void* hyp_w2_t; const float* hyp_b2;      /* [4][256][32], fp32 [4][32] */
    const void* iou_w0_t; const float* iou_b0;      /* iou_prediction_head: [256][256] */
    const void* iou_w1_t; const float* iou_b1;
    const void* iou_w2_t; const float* iou_b2;      /* [256][4] */
} wg_mask_decoder_weights;

WG_API size_t wg_mask_decoder_workspace_bytes(int P, int hw);                      /* up_stages = 1 */
WG_API size_t wg_mask_decoder_workspace_bytes_ex(int P, int hw, int up_stages);
/* img_emb_tokens split-bf16 [B, hw, 512] (channels-last image embeddings, hi | lo), txt_emb fp32 [P, 256] (CTP outputs = sparse prompt
 * embeddings), prompt_img int32 [P] (image index of each prompt; prompts of one image are contiguous).
 * low_res_out fp32 [P, n_out, u*grid_h, u*grid_w] with u = 2^up_stages, iou_out fp32 [P, n_out]; n_out = 1 (multimask_output = 0)
 * or n_mask_tokens - multimask_first (4 for the multi-scale decoder, 3 for the SAM decoder).
 * depth_pool_out (nullable) fp32 [P, 33]: sigmoid(logit)-weighted sums of the 32 upscaled channels + the weight sum
 * (input of this repo's relative-depth extension; not part of the reference). */
WG_API int wg_mask_decoder_forward(const wg_mask_decoder_weights* w, const void* img_emb_tokens_bf16, const float* txt_emb,
                                   const int32_t* prompt_img, int P, int multimask_output, float* low_res_out, float* iou_out,
                                   float* depth_pool_out, void* workspace, size_t workspace_bytes, void* stream);

/* Level > 0 of MaskDecoderMultiScale (image_feature_scale_num = 2; mask_decoder_multi_scale.py:165-171).
 * wg_upsample2x_embedding: upsample_2x = ConvTranspose2d(256 -> 256, k2, s2) + LayerNorm2d + GELU on the image embeddings,
 *   split-bf16 [B, h*w, 512] -> split-bf16 [B, 4*h*w, 512] (raster order of the 2h x 2w grid); w_up bf16 [4*256][T*256] in the
 *   (dy, dx, co) row order of the other ConvTranspose weights, b_up fp32 [1024] tiled, ln_g / ln_b fp32 [256].
 * wg_mask_decoder_forward_level: wg_mask_decoder_forward whose keys are gated by the previous level's masks,
 *   keys0 = (sigmoid(mean_m prev_masks[p, m]) + 1) * embedding + no_mask;  prev_masks fp32 [P, n_prev, grid_h, grid_w] or NULL.
 *   `w` then carries the level's own transformer, out_tokens / sparse_add with level_embed[level] and bias tables from pe1. */
WG_API size_t wg_upsample2x_workspace_bytes(int B, int hw);
WG_API int wg_upsample2x_embedding(const void* w_up, int split_terms, const float* b_up, const float* ln_g, const float* ln_b,
                                   const void* emb_tokens_bf16, int B, int grid_h, int grid_w, void* out_tokens_bf16,
                                   void* workspace, size_t workspace_bytes, void* stream);
WG_API int wg_mask_decoder_forward_level(const wg_mask_decoder_weights* w, const void* img_emb_tokens_bf16, const float* txt_emb,
                                         const int32_t* prompt_img, int P, int multimask_output, const float* prev_masks, int n_prev,
                                         float* low_res_out, float* iou_out, float* depth_pool_out, void* workspace,
                                         size_t workspace_bytes, void* stream);
/* The same with the number of images behind img_emb_tokens_bf16 (prompt_img values are < n_images): layer 0 of the two-way transformer
 * sees identical keys for all prompts of an image, so its image-side projections are then computed once per image instead of once per
 * prompt (n_images <= 0, or previous-level masks: computed per prompt as wg_mask_decoder_forward_level does).  Same results. */
WG_API int wg_mask_decoder_forward_images(const wg_mask_decoder_weights* w, const void* img_emb_tokens_bf16, int n_images, const float* txt_emb,
                                   const int32_t* prompt_img, int P, int multimask_output, const float* prev_masks, int n_prev,
                                   float* low_res_out, float* iou_out, float* depth_pool_out, void* workspace,
                                   size_t workspace_bytes, void* stream);

/* A6 -- PromptEncoder.get_dense_pe / PositionEmbeddingRandom.forward (prompt_encoder.py:67-76, 216-229).
 * gauss fp32 [2, F]; out_chw fp32 [2F, h, w] (reference layout, nullable); out_tokens fp32 [h*w, 2F] (nullable). */
WG_API int wg_dense_pe(const float* gauss, int num_pos_feats, int h, int w, float* out_chw, float* out_tokens, void* stream);

/* A8 + A9 -- postprocess_masks (model/walkgpt.py:749-790, vision_tower_for_mask=True; Sam.postprocess_masks,
 * segment_anything/modeling/sam.py:137-172) fused with the `> 0` threshold (evaluation_walkgpt.py:937) and the
 * mask_score expression (model/walkgpt.py:541 == :742).
 *   low_res fp32 [n_masks, Hm, Wm] -> bilinear(align_corners=False) to target x target -> crop [:in_h, :in_w]
 *   -> bilinear to (out_h, out_w).  logits_out fp32 [n_masks, out_h, out_w]; mask_out uint8 (0/1, nullable);
 *   score_out fp32 [n_masks] (nullable). */
WG_API size_t wg_postprocess_workspace_bytes(int n_masks, int in_h, int in_w);
WG_API int wg_postprocess_masks(const float* low_res, int n_masks, int Hm, int Wm, int target, int in_h, int in_w, int out_h,
                                int out_w, float* logits_out, uint8_t* mask_out, float* score_out, void* workspace,
                                size_t workspace_bytes, void* stream);

/* A10 -- relative-depth head.  NOT PART OF THE REFERENCE (it emits depth as LLM text, utils/PAVE_dataset.py:181-188);
 * this repo's own extension, defined in oracle/path_a.py:depth_head.  pooled fp32 [P,33] comes from
 * wg_mask_decoder_forward; seg_offsets int32 [B+1] (device); w1 fp32 [256,32], b1 [256], w2 [256], b2 [1];
 * depth_out fp32 [P] in [0,1], min-max normalised over the prompts of each image. */
WG_API int wg_depth_head(const float* pooled, const int32_t* seg_offsets, int B, int max_S, const float* w1, const float* b1,
                         const float* w2, const float* b2, float* depth_out, void* stream);

/* ---- the SAM ViT image encoder (Path B's pixel encoder) ------------------------------------------------- */

/* Attention.forward + add_decomposed_rel_pos (segment_anything/modeling/image_encoder.py:222-247, 325-361), head_dim 80, SAM geometry
 * (64 x 64 map; mode 0 = 14 x 14 windows of the zero-padded 70 x 70 map, mode 1 = global).
 *   qkv bf16 [rows, 3*heads*80], columns [Q main (heads x 64) | K main | V main | Q rem (heads x 16) | K rem | V rem]: head h's vector is
 *       its 64 main columns followed by its 16 rem columns; rows: mode 0 = WINDOW-major [image, 25 windows, 196 tokens] (pad tokens
 *       included -- they are keys, as in the reference), mode 1 = image-major [image, 4096].
 *   rel_table bf16: mode 0 [64, 80] = rel_pos_h (27 rows) at row 0 and rel_pos_w at row 32, other rows zero;
 *                   mode 1 [256, 80] = rel_pos_h (127 rows) at row 0 and rel_pos_w at row 128.
 *   out bf16 [n_images * 4096, heads * 80], image-major (mode 0 un-partitions and drops the pad tokens). */
WG_API int wg_sam_attention(const void* qkv, const void* rel_table, void* out, int n_images, int mode, int heads, void* stream);

typedef struct wg_sam_block {
    const float* ln1_g; const float* ln1_b;
    const void* w_qkv;  const float* b_qkv;   /* bf16 [3*hidden, hidden], fp32 [3*hidden]: rows in the column order wg_sam_attention reads */
    const void* rel_table;                     /* see wg_sam_attention */
    const void* w_proj; const float* b_proj;  /* bf16 [hidden, hidden] */
    const float* ln2_g; const float* ln2_b;
    const void* w_fc1;  const float* b_fc1;   /* bf16 [mlp, hidden]  (MLPBlock lin1, erf GELU) */
    const void* w_fc2;  const float* b_fc2;   /* bf16 [hidden, mlp] */
    int32_t is_global; int32_t reserved;
} wg_sam_block;

typedef struct wg_sam_encoder_weights {
    int32_t hidden, heads, mlp, image, patch, depth, out_chans, reserved;  /* ViT-H: 1280, 16, 5120, 1024, 16, 32, 256 */
    const void* patch_w;                      /* bf16 [hidden, 3*patch*patch]: conv weight flattened (c, dy, dx) */
    const float* pos_bias;                    /* fp32 [4096, hidden] = pos_embed + patch_embed.proj.bias */
    const wg_sam_block* blocks;               /* HOST array of `depth` entries */
    wg_proj_neck_weights neck;                /* neck.0 / neck.1 / neck.2 / neck.3 in the fields w_conv1, ln1, w_conv3, ln2; hidden = this hidden */
} wg_sam_encoder_weights;

WG_API size_t wg_sam_encoder_workspace_bytes(const wg_sam_encoder_weights* w, int B);
/* ImageEncoderViT.forward (image_encoder.py:107-116): pixels [B,3,1024,1024] fp32 or bf16 -> image embedding TOKENS, split-bf16
 * [B*4096, 512] (channels-last, hi | lo; nullable) after `n_run` blocks + neck; x_out (nullable) fp32 [B*4096, hidden] receives the
 * residual stream after the n_run blocks (what the reference's blocks return before the neck). */
WG_API int wg_sam_encoder_forward(const wg_sam_encoder_weights* w, const void* pixels, int pixels_is_bf16, int B, int n_run,
                                  void* emb_tokens_split, float* x_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- SURVEY section 8(f), "next" rows: the steps on either side of the grounding path ------------------------------------- */

/* F2a -- visual-token resample handed to the LLM (model/llava_walkgpt/model/llava_arch.py:252-259):
 *   tokens [n, p*p, C] (channels last) -> [n, t*t, C], F.interpolate(bilinear, align_corners=False) computed in fp32 in
 *   PyTorch's operation order, stored in the input dtype (bf16 when is_bf16, else fp32).  C % 4 == 0. */
WG_API int wg_resample_tokens(const void* tokens, int is_bf16, int n, int p, int C, int t, void* out, void* stream);

/* F2b -- [SEG]-row extraction (model/walkgpt.py:287-306 mask, :406-420 gather and per-image offsets):
 *   input_ids int64 [rows, Lin] (device); hidden [rows, L, H] bf16 or fp32 with L == Lin + shift (shift = 255: the image
 *   token expands to 256 visual tokens).  Hidden position q of row r is selected iff q >= shift, i = q - shift + 1 < Lin and
 *   input_ids[r][i] is one of the n_seg_ids (<= 8) token ids in the HOST array seg_ids.  Selected rows are copied to
 *   out [max_out, H] in (row, position) order (rows past max_out are dropped; the caller compares row_offsets[rows] with
 *   max_out); counts int32 [rows]; row_offsets int32 [rows+1] = exclusive scan; img_offsets[b] = row_offsets[img_rows[b]]
 *   for the n_img_p1 entries of the image -> text-row map `offset` (img_rows may be NULL).  No host synchronisation. */
WG_API int wg_seg_gather(const int64_t* input_ids, int rows, int Lin, const void* hidden, int hidden_is_bf16, int L, int H,
                         const int64_t* seg_ids, int n_seg_ids, int shift, const int32_t* img_rows, int n_img_p1, void* out,
                         int max_out, int32_t* counts, int32_t* row_offsets, int32_t* img_offsets, void* stream);

/* Batched-composition glue (replaces the reference's per-image Python loop over `seg_token_offset`, model/walkgpt.py:414-420,
 * 511-541): prompt_img[p] = the image b with seg_offsets[b] <= p < seg_offsets[b+1]; seg_offsets int32 [B+1] on the device.
 * Inconsistent offsets (not non-decreasing from 0 to P) never index out of bounds: results are clamped to [0, B-1] and
 * *status (device int32, nullable) is set to 1, else 0.  No host synchronisation. */
WG_API int wg_prompt_index(const int32_t* seg_offsets, int B, int P, int32_t* prompt_img, int32_t* status, void* stream);

/* F4 -- intersectionAndUnionGPU (utils/utils.py:192-204) for n_masks (output, target) pairs of `pixels` uint8 class ids:
 *   output[target == ignore_index] = ignore_index; out fp32 [n_masks, 3, K] = per-class histograms (area_intersection,
 *   area_union = area_output + area_target - area_intersection, area_target), exactly as torch.histc(bins=K, min=0, max=K-1)
 *   returns them (values outside [0, K-1] are dropped).  1 <= K <= 8. */
WG_API size_t wg_intersection_and_union_workspace_bytes(int n_masks, int K);
WG_API int wg_intersection_and_union(const uint8_t* output, const uint8_t* target, int n_masks, int64_t pixels, int K,
                                     int ignore_index, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* F4 -- cost matrix of match_pred (utils/matcher.py:93-128), everything up to the host-side linear_sum_assignment:
 *   pred_logits fp32 [n_pred, H, W]; tgt_masks fp32 (or uint8 when tgt_is_u8) [n_tgt, H, W]; point_coords fp32 [num_points, 2]
 *   = (x, y) in [0, 1]^2, the ONE point set all masks share (matcher.py:101-102; the caller draws it, torch.rand in the
 *   reference).  Both mask sets are sampled with point_sample (matcher.py:64-90: F.grid_sample bilinear, zero padding,
 *   align_corners=False); cost fp32 [n_pred, n_tgt] = batch_sigmoid_ce_loss + batch_dice_loss (matcher.py:10-58, both with
 *   weight 1 as in match_pred).  Deterministic (fixed-order partial sums); fp32 sums in a different order than torch.einsum,
 *   so the parity tolerance is 2e-5 absolute on costs of O(1).  1 <= n_pred, n_tgt <= 64. */
WG_API size_t wg_match_cost_workspace_bytes(int n_pred, int n_tgt, int num_points);
WG_API int wg_match_cost(const float* pred_logits, const void* tgt_masks, int tgt_is_u8, const float* point_coords, int n_pred,
                         int n_tgt, int H, int W, int num_points, float* cost, void* workspace, size_t workspace_bytes,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WALKGPT_B200_H */
