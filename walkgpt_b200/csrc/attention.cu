// Fused flash-style multi-head attention for the CLIP ViT-L/14@448 tower: head_dim 64,
// T = 1025 tokens, optional key-padding mask (custom_clip.py:27-38 semantics: masked KEYS never receive weight).
//
// One CTA per (128-query tile, head, image); two CTAs co-reside per SM so that one CTA's softmax overlaps the
// other's tensor work.  Per CTA:
//   warp 0  : TMA producer  (Q once, K/V blocks of 128 keys through a 2-stage ring; 3-D tensor map over
//             [image, token, 3*heads*64] so rows past T are zero-filled by the hardware)
//   warp 1  : TMEM allocation + single-thread tcgen05.mma issue:  S = Q K^T (128x128x64) into TMEM,
//             O_j = P_j V_j (128x64x128, V consumed as an MN-major operand straight from its TMA tile)
//   warps 2-5: softmax warpgroup, ONE THREAD PER QUERY ROW (TMEM lane == row, so row max / row sum need no
//             shuffles): tcgen05.ld S -> online softmax in fp32 (exp2 with folded scale) -> P as bf16 into
//             128B-swizzled smem (A operand of the PV MMA) -> accumulate O_j from TMEM into registers with the
//             running rescale -> final 1/l, bf16, TMA store.
#include "host.h"
#include "ptx.cuh"

namespace wg {
namespace {

constexpr int ATT_BQ = 128;      // queries per CTA
constexpr int ATT_BKV = 128;     // keys per block
constexpr int ATT_D = 64;        // head dim
constexpr int ATT_THREADS = 192;  // warp 0: TMA, warp 1: MMA, warps 2-5: softmax (one thread per query row)
constexpr int ATT_SOFTMAX_THREADS = 128;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: a [128 x 64] bf16 tile
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;          // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;      // 2 stages
constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;      // [128 x 128] bf16 = two K-atoms of 64 keys
constexpr int OFF_BAR = OFF_P + 2 * TILE_BYTES;
constexpr int ATT_NUM_BARS = 12;
constexpr int ATT_SMEM_BYTES = OFF_BAR + ATT_NUM_BARS * 8 + 16;
constexpr int ATT_TMEM_COLS = 256;  // S: [0,128)  O: [128,192)
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units: the running scale is refreshed only when the row max grew by > 2^8

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x on the FMA/ALU pipes (no MUFU): round-to-nearest split x = j + f, f in [-0.5, 0.5], degree-3 minimax polynomial for
// 2^f (max rel. error 1.0e-4, far below the bf16 rounding of P), exponent add for 2^j.  The B200 SFU evaluates ex2 at
// 2 lanes/clk/SMSP, which makes d=64 attention SFU-bound; a fixed fraction of the exponentials is routed here instead.
__device__ __forceinline__ float exp2_poly(float x) {
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;            // 1.5 * 2^23: the low mantissa bits now hold round(x)
    const float f = x - (t - 12582912.0f);
    float pz = fmaf(0.05500893f, f, 0.24221096f);
    pz = fmaf(pz, f, 0.69328293f);
    pz = fmaf(pz, f, 1.0f);
    return __int_as_float(__float_as_int(pz) + (__float_as_int(t) << 23));
}
#ifndef ATT_POLY_EXP
#define ATT_POLY_EXP 0  // measured on B200: with one softmax warp per scheduler the kernel is latency- not SFU-bound; kept for tuning
#endif
// which of the 32 columns of a chunk use the polynomial (3 of 8)
__device__ __forceinline__ constexpr bool use_poly(int i) { return (i & 7) == 1 || (i & 7) == 4 || (i & 7) == 6; }

struct AttnParams {
    int T, heads, num_kv_blocks;
    int n_last;  // keys covered by the last key block, rounded up to 16 (T = 1025: 16 -- the block holds only the 1025th token)
    float scale_log2;
    const uint8_t* key_valid;  // [B, T] or null
};

struct SoftmaxState {
    float m_ref;  // the maximum the current scale refers to
    float l_run;  // running row sum in that scale
};

// One key block for one query row (= one thread): S row (NCH x 32 columns) from TMEM -> registers, hand S back, online softmax
// with the lazily refreshed scale, P row (bf16) -> smem.
template <int NCH>
__device__ __forceinline__ void softmax_block(SoftmaxState& st, int j, const AttnParams& p, const uint8_t* kvalid, uint32_t tmem_S, uint32_t tmem_O,
                                              uint32_t lane_off, int r, uint8_t* p_row, uint64_t* s_free, uint64_t* p_free, uint64_t* p_ready) {
    const int key0 = j * ATT_BKV;
    const bool need_mask = (key0 + NCH * 32 > p.T) || (kvalid != nullptr);
    uint32_t sv[NCH][32];
#pragma unroll
    for (int c = 0; c < NCH; ++c) tmem_ld_32x32b_x32(tmem_S + lane_off + c * 32, sv[c]);
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(s_free);  // S_j is in registers: the tensor core may overwrite it with S_{j+1}
    if (need_mask) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int key = key0 + c * 32 + i;
                const bool ok = key < p.T && (kvalid == nullptr || kvalid[key] != 0);
                if (!ok) sv[c][i] = 0xff800000u;  // -inf
            }
    }
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains (ILP)
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[c][i]));
    const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    // lazily refreshed scale
    float alpha = 1.0f;
    bool refresh = false;
    if (j == 0) {
        st.m_ref = (mx == -INFINITY) ? 0.f : mx;
    } else if ((mx - st.m_ref) * p.scale_log2 > RESCALE_THRESHOLD) {
        alpha = ex2_approx((st.m_ref - mx) * p.scale_log2);
        st.m_ref = mx;
        refresh = true;
    }
    const float neg_m = -st.m_ref * p.scale_log2;
    float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float x = fmaf(__uint_as_float(sv[c][i]), p.scale_log2, neg_m);  // -inf for masked keys -> 0
            const float e = (ATT_POLY_EXP && use_poly(i)) ? exp2_poly(x) : ex2_approx(x);
            rs4[i & 3] += e;
            sv[c][i] = __float_as_uint(e);
        }
    st.l_run = st.l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
    if (j > 0) {
        mbar_wait(p_free, (j - 1) & 1);  // PV_{j-1} finished: P buffer free, O quiescent
        if (__any_sync(0xffffffffu, refresh)) {
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t ov[32];
                tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, ov);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                tmem_st_32x32b_x32(tmem_O + lane_off + c * 32, ov);
            }
            tmem_st_wait();
        }
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        uint8_t* atom_row = p_row + (c >> 1) * TILE_BYTES;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(sv[c][g * 8 + 0]), __uint_as_float(sv[c][g * 8 + 1]));
            u.y = pack_bf16x2(__uint_as_float(sv[c][g * 8 + 2]), __uint_as_float(sv[c][g * 8 + 3]));
            u.z = pack_bf16x2(__uint_as_float(sv[c][g * 8 + 4]), __uint_as_float(sv[c][g * 8 + 5]));
            u.w = pack_bf16x2(__uint_as_float(sv[c][g * 8 + 6]), __uint_as_float(sv[c][g * 8 + 7]));
            const int chunk = ((c & 1) * 4 + g) ^ (r & 7);
            *reinterpret_cast<uint4*>(atom_row + chunk * 16) = u;
        }
    }
    fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core (async proxy)
    tc_fence_before();
    mbar_arrive(p_ready);
}

// Pipeline (per CTA = one 128-query tile of one head of one image; two CTAs share an SM):
//   S_j = Q K_j^T (TMEM)  ->  softmax threads pull their S row into registers and immediately hand the S buffer back
//   (s_free) so the tensor core computes S_{j+1} while the exponentials of block j are evaluated  ->  P_j (bf16, smem)
//   ->  O += P_j V_j accumulated IN TMEM across all key blocks.  The softmax uses a lazily refreshed scale: the reference
//   maximum is only moved (and O / l rescaled through tcgen05.ld/st) when a row's maximum grew by more than 2^8, which is
//   exact arithmetic (any common scale cancels in O / l) and takes the O round trip off the per-block critical path.
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_d64_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzled tiles need 1024B alignment

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* v_full = bars + 3;    // [2]
    uint64_t* kv_empty = bars + 5;  // [2]
    uint64_t* s_full = bars + 7;
    uint64_t* s_free = bars + 8;
    uint64_t* p_ready = bars + 9;
    uint64_t* p_free = bars + 10;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + ATT_NUM_BARS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qt = blockIdx.x;
    const int head = blockIdx.y;
    const int img = blockIdx.z;
    const int HD = p.heads * ATT_D;
    const int nkb = p.num_kv_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmO);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, ATT_SOFTMAX_THREADS);
        mbar_init(p_ready, ATT_SOFTMAX_THREADS);
        mbar_init(p_free, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_S = tmem_base;
    const uint32_t tmem_O = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
            tma_load_3d(smem + OFF_Q, &tmQKV, q_full, head * ATT_D, qt * ATT_BQ, img);
            for (int j = 0; j < nkb; ++j) {
                const int s = j & 1;
                mbar_wait_relaxed(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
                tma_load_3d(smem + OFF_K + s * TILE_BYTES, &tmQKV, &k_full[s], HD + head * ATT_D, j * ATT_BKV, img);
                mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
                tma_load_3d(smem + OFF_V + s * TILE_BYTES, &tmQKV, &v_full[s], 2 * HD + head * ATT_D, j * ATT_BKV, img);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, false, false);  // Q (K-major) x K (K-major)
            constexpr uint32_t IDESC_O = umma_idesc_bf16(128, 64, false, true);    // P (K-major) x V (MN-major)
            // the last key block only spans n_last (multiple of 16) keys: narrower S MMA, fewer PV k-steps
            const uint32_t idesc_s_last = umma_idesc_bf16(128, p.n_last, false, false);
            const int pv_steps_last = p.n_last / 16;
            const uint32_t q_addr = smem_u32(smem + OFF_Q);
            const uint32_t p_addr = smem_u32(smem + OFF_P);
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            {
                const uint32_t k_addr = smem_u32(smem + OFF_K);
                const uint32_t idesc = nkb == 1 ? idesc_s_last : IDESC_S;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16_ss(tmem_S, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idesc, k != 0);
                umma_commit(s_full);
            }
            for (int j = 0; j < nkb; ++j) {
                const int s = j & 1;
                if (j + 1 < nkb) {
                    // S_{j+1} as soon as the softmax threads have pulled S_j out of TMEM: overlaps their exponentials
                    const int s1 = (j + 1) & 1;
                    mbar_wait(s_free, j & 1);
                    mbar_wait(&k_full[s1], ((j + 1) >> 1) & 1);
                    tc_fence_after();
                    const uint32_t k_addr = smem_u32(smem + OFF_K + s1 * TILE_BYTES);
                    const uint32_t idesc = (j + 2 == nkb) ? idesc_s_last : IDESC_S;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_S, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idesc, k != 0);
                    umma_commit(s_full);
                }
                mbar_wait(p_ready, j & 1);
                mbar_wait(&v_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(smem + OFF_V + s * TILE_BYTES);
                const int pv_steps = (j + 1 == nkb) ? pv_steps_last : 8;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // A: P atom (k/4) of 64 keys, 16-key step inside the atom = +32 B;  B: V rows (keys) step 16 rows = +2048 B
                    if (k < pv_steps)
                        umma_f16_ss(tmem_O, umma_desc_sw128(p_addr + (k >> 2) * TILE_BYTES + (k & 3) * 32),
                                    umma_desc_sw128(v_addr + k * 2048), IDESC_O, (j | k) != 0);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(p_free);  // P buffer reusable, O consistent
            }
        }
    } else {
        // ===================== softmax warpgroup: one thread per query row =====================
        // (A variant with two threads per row -- 8 softmax warps per CTA -- was measured slower on B200: 1.06 ms vs 0.80 ms
        //  for B=64, T=1025, 16 heads; the extra named barriers / smem exchange cost more than the added warp parallelism.)
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const uint8_t* kvalid = p.key_valid ? p.key_valid + (size_t)img * p.T : nullptr;
        uint8_t* p_row = smem + OFF_P + r * 128;

        // warps whose 32 query rows all lie past T (last query tile: only the 1025th token is real) skip the arithmetic and only
        // keep the barrier protocol alive; their P / O rows are never stored (the TMA store clips at T).
        const bool warp_active = qt * ATT_BQ + quarter * 32 < p.T;
        SoftmaxState st{0.f, 0.f};
        for (int j = 0; j < nkb; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            if (!warp_active) {
                tc_fence_before();
                mbar_arrive(s_free);
                if (j > 0) mbar_wait(p_free, (j - 1) & 1);
                mbar_arrive(p_ready);
                continue;
            }
            const int ncols = (j + 1 == nkb) ? p.n_last : ATT_BKV;
            if (ncols > 96) softmax_block<4>(st, j, p, kvalid, tmem_S, tmem_O, lane_off, r, p_row, s_free, p_free, p_ready);
            else if (ncols > 64) softmax_block<3>(st, j, p, kvalid, tmem_S, tmem_O, lane_off, r, p_row, s_free, p_free, p_ready);
            else if (ncols > 32) softmax_block<2>(st, j, p, kvalid, tmem_S, tmem_O, lane_off, r, p_row, s_free, p_free, p_ready);
            else softmax_block<1>(st, j, p, kvalid, tmem_S, tmem_O, lane_off, r, p_row, s_free, p_free, p_ready);
        }
        const float l_run = st.l_run;
        // ---- epilogue: O / l, bf16, stage into the (now idle) Q tile, TMA store
        mbar_wait(p_free, (nkb - 1) & 1);
        tc_fence_after();
        const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
        uint8_t* o_row = smem + OFF_Q + r * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (!warp_active) break;
            uint32_t ov[32];
            tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]) * inv_l, __uint_as_float(ov[g * 8 + 1]) * inv_l);
                u.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]) * inv_l, __uint_as_float(ov[g * 8 + 3]) * inv_l);
                u.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]) * inv_l, __uint_as_float(ov[g * 8 + 5]) * inv_l);
                u.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]) * inv_l, __uint_as_float(ov[g * 8 + 7]) * inv_l);
                *reinterpret_cast<uint4*>(o_row + (((c * 4 + g) ^ (r & 7)) * 16)) = u;
            }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, ATT_SOFTMAX_THREADS);
        if (threadIdx.x == 64) {
            tma_store_3d(&tmO, smem + OFF_Q, head * ATT_D, qt * ATT_BQ, img);
            tma_store_commit();
            tma_store_wait_all<0>();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
    }
}

}  // namespace
}  // namespace wg

extern "C" int wg_attention_d64(const void* qkv, void* out, const uint8_t* key_valid, int B, int T, int heads, float scale,
                                void* stream_) {
    using namespace wg;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(qkv && out, "wg_attention_d64: null pointer");
    WG_REQUIRE(B > 0 && T > 0 && heads > 0, "wg_attention_d64: bad sizes B=%d T=%d heads=%d", B, T, heads);
    WG_REQUIRE(heads <= 65535 && B <= 65535, "wg_attention_d64: grid too large");
    if (!device_is_sm100()) {
        set_error("wg_attention_d64: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const uint64_t HD = (uint64_t)heads * ATT_D;
    CUtensorMap tmQKV, tmO;
    {
        uint64_t dims[3] = {3 * HD, (uint64_t)T, (uint64_t)B};
        uint64_t strides[2] = {3 * HD * 2, (uint64_t)T * 3 * HD * 2};
        uint32_t box[3] = {ATT_D, 128, 1};
        WG_TRY(make_tensor_map(&tmQKV, qkv, 2, 3, dims, strides, box));
    }
    {
        uint64_t dims[3] = {HD, (uint64_t)T, (uint64_t)B};
        uint64_t strides[2] = {HD * 2, (uint64_t)T * HD * 2};
        uint32_t box[3] = {ATT_D, 128, 1};
        WG_TRY(make_tensor_map(&tmO, out, 2, 3, dims, strides, box));
    }
    AttnParams p;
    p.T = T;
    p.heads = heads;
    p.num_kv_blocks = (T + ATT_BKV - 1) / ATT_BKV;
    p.n_last = ((T - (p.num_kv_blocks - 1) * ATT_BKV) + 15) / 16 * 16;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.key_valid = key_valid;
    static bool attr_set = false;
    if (!attr_set) {
        WG_CHECK_CUDA(cudaFuncSetAttribute(attention_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid((T + ATT_BQ - 1) / ATT_BQ, heads, B);
    Prof prof("attention_d64", stream, 4.0 * B * heads * (double)T * T * ATT_D, 2.0 * 4.0 * B * (double)T * heads * ATT_D);
    attention_d64_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tmQKV, tmO, p);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}
