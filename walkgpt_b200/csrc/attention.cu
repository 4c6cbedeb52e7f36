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
constexpr int ATT_THREADS = 192;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: a [128 x 64] bf16 tile
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;          // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;      // 2 stages
constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;      // [128 x 128] bf16 = two K-atoms of 64 keys
constexpr int OFF_BAR = OFF_P + 2 * TILE_BYTES;
constexpr int ATT_NUM_BARS = 10;
constexpr int ATT_SMEM_BYTES = OFF_BAR + ATT_NUM_BARS * 8 + 16;
constexpr int ATT_TMEM_COLS = 256;  // S: [0,128)  O_j: [128,192)

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttnParams {
    int T, heads, num_kv_blocks;
    float scale_log2;
    const uint8_t* key_valid;  // [B, T] or null
};

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
    uint64_t* p_ready = bars + 8;
    uint64_t* o_full = bars + 9;
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
        mbar_init(p_ready, 128);
        mbar_init(o_full, 1);
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
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
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
            const uint32_t q_addr = smem_u32(smem + OFF_Q);
            const uint32_t p_addr = smem_u32(smem + OFF_P);
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            {
                const uint32_t k_addr = smem_u32(smem + OFF_K);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16_ss(tmem_S, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), IDESC_S, k != 0);
                umma_commit(s_full);
            }
            for (int j = 0; j < nkb; ++j) {
                const int s = j & 1;
                mbar_wait(p_ready, j & 1);
                mbar_wait(&v_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(smem + OFF_V + s * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // A: P atom (k/4) of 64 keys, 16-key step inside the atom = +32 B;  B: V rows (keys) step 16 rows = +2048 B
                    umma_f16_ss(tmem_O, umma_desc_sw128(p_addr + (k >> 2) * TILE_BYTES + (k & 3) * 32),
                                umma_desc_sw128(v_addr + k * 2048), IDESC_O, k != 0);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(o_full);
                if (j + 1 < nkb) {
                    const int s1 = (j + 1) & 1;
                    mbar_wait(&k_full[s1], ((j + 1) >> 1) & 1);
                    tc_fence_after();
                    const uint32_t k_addr = smem_u32(smem + OFF_K + s1 * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_S, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), IDESC_S, k != 0);
                    umma_commit(s_full);
                }
            }
        }
    } else {
        // ===================== softmax warpgroup: one thread per query row =====================
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const uint8_t* kvalid = p.key_valid ? p.key_valid + (size_t)img * p.T : nullptr;
        float m_run = -INFINITY, l_run = 0.f;
        float o_acc[ATT_D];
#pragma unroll
        for (int i = 0; i < ATT_D; ++i) o_acc[i] = 0.f;
        uint8_t* p_row = smem + OFF_P + r * 128;

        for (int j = 0; j < nkb; ++j) {
            const int key0 = j * ATT_BKV;
            const bool need_mask = (key0 + ATT_BKV > p.T) || (kvalid != nullptr);
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            // pass 1: row maximum
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_S + lane_off + c * 32, v);
                tmem_ld_wait();
                if (need_mask) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        int key = key0 + c * 32 + i;
                        bool ok = key < p.T && (kvalid == nullptr || kvalid[key] != 0);
                        if (ok) mx = fmaxf(mx, __uint_as_float(v[i]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            }
            const float m_new = fmaxf(m_run, mx);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;  // whole row masked so far
            const float alpha = ex2_approx((m_run - m_use) * p.scale_log2);
            const float neg_m = -m_use * p.scale_log2;
            // pass 2: probabilities -> bf16 P tile (A operand of the PV MMA), row sum
            float rowsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                float f[32];
                tmem_ld_32x32b_x32(tmem_S + lane_off + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m));
                if (need_mask) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        int key = key0 + c * 32 + i;
                        bool ok = key < p.T && (kvalid == nullptr || kvalid[key] != 0);
                        if (!ok) f[i] = 0.f;
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) rowsum += f[i];
                uint8_t* atom_row = p_row + (c >> 1) * TILE_BYTES;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 u;
                    u.x = pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]);
                    u.y = pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]);
                    u.z = pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]);
                    u.w = pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]);
                    int chunk = ((c & 1) * 4 + g) ^ (r & 7);
                    *reinterpret_cast<uint4*>(atom_row + chunk * 16) = u;
                }
            }
            l_run = l_run * alpha + rowsum;
            m_run = m_new;
            fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core (async proxy)
            tc_fence_before();
            mbar_arrive(p_ready);
            // O_j from TMEM, rescale-and-accumulate in registers
            mbar_wait(o_full, j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha, __uint_as_float(v[i]));
            }
        }
        // ---- epilogue: normalise, bf16, stage into the (now idle) Q tile, TMA store
        const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
        uint8_t* o_row = smem + OFF_Q + r * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            uint4 u;
            u.x = pack_bf16x2(o_acc[g * 8 + 0] * inv_l, o_acc[g * 8 + 1] * inv_l);
            u.y = pack_bf16x2(o_acc[g * 8 + 2] * inv_l, o_acc[g * 8 + 3] * inv_l);
            u.z = pack_bf16x2(o_acc[g * 8 + 4] * inv_l, o_acc[g * 8 + 5] * inv_l);
            u.w = pack_bf16x2(o_acc[g * 8 + 6] * inv_l, o_acc[g * 8 + 7] * inv_l);
            *reinterpret_cast<uint4*>(o_row + ((g ^ (r & 7)) * 16)) = u;
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
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
