// Attention of the SAM ViT image encoder (segment_anything/modeling/image_encoder.py:196-247 Attention.forward and
// :325-361 add_decomposed_rel_pos): head_dim 80, softmax(scale * q k^T + rel_h[q, k_h] + rel_w[q, k_w]) v, for the two block kinds
//   MODE 0  windowed blocks: 14 x 14 windows of the zero-padded 70 x 70 map (196 tokens; the pad tokens ARE keys, as in the reference)
//   MODE 1  global blocks:   the whole 64 x 64 map (4096 tokens)
//
// One CTA = one 128-query tile of one head of one window / image; two CTAs per SM.  Warp 0 = TMA producer, warp 1 = tcgen05.mma
// issuer, warps 2-5 = one softmax thread per query row (TMEM lane == row).  head_dim 80 = 64 + 16: every operand is a 64-column tile
// under the 128-byte swizzle plus a 16-column tile under the 32-byte swizzle (the QKV GEMM writes its columns as
// [Q main | K main | V main | Q rem | K rem | V rem], so both tiles are plain 2-D TMA boxes); S = Q K^T takes 4 + 1 k-steps, O += P V
// is one MMA with N = 64 and one with N = 16 against the two V tiles (MN-major operands straight from their TMA tiles, P read from
// tensor memory).  Key blocks of 64 keys: tensor memory S [0,64) | P [64,96) | O [96,176) keeps two CTAs resident per SM.
//
// Decomposed relative position (the reference gathers [q_h, k_h, C] tables and runs two einsums over the UNSCALED query): per query tile
// a prologue MMA forms T = Q [R_h ; R_w]^T -- the query against EVERY relative offset (2g - 1 table rows per axis) -- in the still idle
// tensor memory; each softmax thread reads its row, and rel_h[k_h] = T[q_h - k_h + g - 1], rel_w[k_w] = T[q_w - k_w + g - 1] are a
// reversed contiguous slice of it.  MODE 1: a key block is exactly one map row, so rel_h is ONE scalar per (row, block) folded into the
// running-max shift and rel_w is 64 registers reused by all 64 blocks; MODE 0: 14 + 14 registers, indexed at compile time (the four
// key blocks are unrolled).  The tables go through the K/V stage buffers before the ring starts.
#include "host.h"
#include "ptx.cuh"
#include <cstdlib>

namespace wg {
namespace {

constexpr int SA_BQ = 128, SA_BKV = 64, SA_D = 80;
constexpr int SA_WS = 14, SA_WTOK = SA_WS * SA_WS;  // window side / tokens per window
constexpr int SA_G = 64;                            // map side
constexpr int SA_OFF_QM = 0;                        // Q main  [128 x 64] bf16, 128B swizzle            16 KB
constexpr int SA_OFF_KM = 16384;                    // K main  2 stages x [64 x 64]                     16 KB
constexpr int SA_OFF_VM = SA_OFF_KM + 16384;        // V main  2 stages                                 16 KB
constexpr int SA_OFF_QR = SA_OFF_VM + 16384;        // Q rem   [128 x 16] bf16, 32B swizzle              4 KB
constexpr int SA_OFF_KR = SA_OFF_QR + 4096;         // K rem   2 stages x [64 x 16]                      4 KB
constexpr int SA_OFF_VR = SA_OFF_KR + 4096;         // V rem   2 stages                                  4 KB
constexpr int SA_OFF_BAR = SA_OFF_VR + 4096;
constexpr int SA_NUM_BARS = 16;
constexpr int SA_SMEM_BYTES = SA_OFF_BAR + SA_NUM_BARS * 8 + 16;
constexpr int SA_TMEM_COLS = 256;
constexpr uint32_t SA_TM_S = 0, SA_TM_P = 64, SA_TM_O = 96, SA_TM_OR = 160;
constexpr float SA_LOG2E = 1.4426950408889634f;
constexpr float SA_RESCALE = 8.0f;
constexpr uint32_t LT_SW128 = 2, LT_SW32 = 6;

struct SamAttnParams {
    int dbg;  // bring-up switches (WG_SAM_DBG, debug builds of the host path only): 1 = no SW32 (rem) MMAs, 2 = no rem TMA loads either
    int heads;
    float scale_log2;  // head_dim^-0.5 * log2(e)
    __nv_bfloat16* out;  // [B * 64 * 64, heads * 80]
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct SmState {
    float m_ref, l_run;
};

// bring-up watchdog (WG_SAM_DBG & 4): a wait that gives up after ~2^26 polls, names the barrier and traps instead of hanging the GPU
__device__ int g_sa_dbg = 0;
__device__ __noinline__ void sa_wait_dbg(uint64_t* bar, uint32_t parity, int id) {
    const long long t0 = clock64();
    while (clock64() - t0 < 2000000000ll)
        if (mbar_try_wait(bar, parity)) return;
    printf("sam_attention watchdog: barrier %d parity %u never completed (block %d,%d,%d thread %d)\n", id, parity, blockIdx.x, blockIdx.y, blockIdx.z,
           threadIdx.x);
    __trap();
}
__device__ int* g_sa_marks = nullptr;  // bring-up: host-mapped array, one slot per thread = id of the barrier it is waiting on
__device__ __forceinline__ void sa_wait(uint64_t* bar, uint32_t parity, int id) {
    if (g_sa_dbg & 8) {
        int* m = g_sa_marks + ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 192 + threadIdx.x;
        *reinterpret_cast<volatile int*>(m) = id * 2 + (int)parity;
        mbar_wait(bar, parity);
        *reinterpret_cast<volatile int*>(m) = -id;
        return;
    }
    if (g_sa_dbg & 4) sa_wait_dbg(bar, parity, id);
    else mbar_wait(bar, parity);
}

// One key block (NC columns) for one query row.  MODE 0: J = compile-time block index (bias registers bh[14], bw[14], keys >= 196 are
// masked); MODE 1: bw[64] registers + the block's scalar bhj.  All logits are kept in log2 units.
template <int MODE, int J>
__device__ __forceinline__ void sam_softmax_block(SmState& st, int j, const float (&bw)[MODE == 0 ? SA_WS : SA_G], const float (&bh)[SA_WS], float bhj,
                                                  float sc, uint32_t tmem_base, uint32_t lane_off, uint64_t* s_free, uint64_t* p_free,
                                                  uint64_t* p_ready) {
    constexpr int NC = (MODE == 0 && J == 3) ? 16 : 64;
    uint32_t sv[NC];
    if constexpr (NC == 64) {
        uint32_t a[32], b[32];
        tmem_ld_32x32b_x32(tmem_base + SA_TM_S + lane_off, a);
        tmem_ld_32x32b_x32(tmem_base + SA_TM_S + lane_off + 32, b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) { sv[i] = a[i]; sv[32 + i] = b[i]; }
    } else {
        uint32_t a[16];
        tmem_ld_32x32b_x16(tmem_base + SA_TM_S + lane_off, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) sv[i] = a[i];
    }
    tc_fence_before();
    mbar_arrive(s_free);  // S_j is in registers: the tensor core may overwrite it with S_{j+1}
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        float t;
        if constexpr (MODE == 0) {
            constexpr int dummy = 0;
            (void)dummy;
            const int key = J * SA_BKV + i;  // compile-time after unrolling
            t = key < SA_WTOK ? fmaf(__uint_as_float(sv[i]), sc, bh[key / SA_WS]) + bw[key % SA_WS] : -INFINITY;
        } else {
            t = fmaf(__uint_as_float(sv[i]), sc, bw[i]);
        }
        sv[i] = __float_as_uint(t);
        mx4[i & 3] = fmaxf(mx4[i & 3], t);
    }
    const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) + bhj;  // bhj = 0 in MODE 0
    float alpha = 1.0f;
    bool refresh = false;
    if (j == 0) {
        st.m_ref = mx;
    } else if (mx - st.m_ref > SA_RESCALE) {
        alpha = ex2f(st.m_ref - mx);
        st.m_ref = mx;
        refresh = true;
    }
    const float c = bhj - st.m_ref;
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NC; i += 2) {
        const float e0 = ex2f(__uint_as_float(sv[i]) + c), e1 = ex2f(__uint_as_float(sv[i + 1]) + c);
        rs[(i >> 1) & 3] += e0 + e1;
        sv[i >> 1] = pack_bf16x2(e0, e1);
    }
    if (j > 0) {
        sa_wait(p_free, (j - 1) & 1, 1);  // PV_{j-1} finished: P buffer free, O quiescent
        if (__any_sync(0xffffffffu, refresh)) {
            tc_fence_after();
            uint32_t ov[32];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                tmem_ld_32x32b_x32(tmem_base + SA_TM_O + lane_off + cc * 32, ov);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                tmem_st_32x32b_x32(tmem_base + SA_TM_O + lane_off + cc * 32, ov);
            }
            uint32_t orr[16];
            tmem_ld_32x32b_x16(tmem_base + SA_TM_OR + lane_off, orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
            tmem_st_32x32b_x16(tmem_base + SA_TM_OR + lane_off, orr);
            tmem_st_wait();
        }
    }
#pragma unroll
    for (int q = 0; q < NC / 8; ++q)  // 8 keys = 4 packed 32-bit columns
        tmem_st_32x32b_x4(tmem_base + SA_TM_P + lane_off + q * 4, sv[q * 4 + 0], sv[q * 4 + 1], sv[q * 4 + 2], sv[q * 4 + 3]);
    st.l_run = st.l_run * alpha + ((rs[0] + rs[1]) + (rs[2] + rs[3]));
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(p_ready);
}

template <int MODE>
__global__ void __launch_bounds__(192, 2)
sam_attention_kernel(const __grid_constant__ CUtensorMap tmMain, const __grid_constant__ CUtensorMap tmRem, const __grid_constant__ CUtensorMap tmTabMain,
                     const __grid_constant__ CUtensorMap tmTabRem, const SamAttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SA_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* tab_full = bars + 1;
    uint64_t* t_full = bars + 2;   // prologue MMAs (T) complete: tables consumed, T readable
    uint64_t* t_cons = bars + 3;   // the 128 softmax threads have read T: tensor memory free for S / P / O
    uint64_t* k_full = bars + 4;   // [2]
    uint64_t* v_full = bars + 6;   // [2]
    uint64_t* k_empty = bars + 8;  // [2]
    uint64_t* v_empty = bars + 10; // [2]
    uint64_t* s_full = bars + 12;
    uint64_t* s_free = bars + 13;
    uint64_t* p_ready = bars + 14;
    uint64_t* p_free = bars + 15;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + SA_NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, unit = blockIdx.z;  // unit = window (MODE 0) or image (MODE 1)
    const int heads = p.heads;
    constexpr int NKEYS = MODE == 0 ? SA_WTOK : SA_G * SA_G;
    constexpr int NKB = (NKEYS + SA_BKV - 1) / SA_BKV;  // 4 / 64
    constexpr int N_LAST = MODE == 0 ? 16 : 64;         // keys (rounded up to 16) covered by the last block: 196 = 3 * 64 + 4
    const int row0 = unit * NKEYS;                      // first qkv row of this window / image
    // qkv column bases
    const int cQm = head * 64, cKm = heads * 64 + head * 64, cVm = 2 * heads * 64 + head * 64;
    const int cQr = 3 * heads * 64 + head * 16, cKr = cQr + heads * 16, cVr = cKr + heads * 16;
    constexpr int TAB_ROWS = MODE == 0 ? 64 : 256;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmMain);
        tma_prefetch_desc(&tmRem);
        tma_prefetch_desc(&tmTabMain);
        tma_prefetch_desc(&tmTabRem);
        mbar_init(q_full, 1);
        mbar_init(tab_full, 1);
        mbar_init(t_full, 1);
        mbar_init(t_cons, 128);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 128);
        mbar_init(p_ready, 128);
        mbar_init(p_free, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<SA_TMEM_COLS>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            const bool rem_tma = !(p.dbg & 2);
            const uint32_t RB = rem_tma ? 1u : 0u;
            mbar_arrive_expect_tx(q_full, 16384 + RB * 4096);
            tma_load_2d(smem + SA_OFF_QM, &tmMain, q_full, cQm, row0 + qt * SA_BQ);
            tma_load_2d(smem + SA_OFF_QM + 8192, &tmMain, q_full, cQm, row0 + qt * SA_BQ + 64);
            if (rem_tma) {
                tma_load_2d(smem + SA_OFF_QR, &tmRem, q_full, cQr, row0 + qt * SA_BQ);
                tma_load_2d(smem + SA_OFF_QR + 2048, &tmRem, q_full, cQr, row0 + qt * SA_BQ + 64);
            }
            // relative-position tables through the (still idle) stage buffers
            mbar_arrive_expect_tx(tab_full, TAB_ROWS * (128 + RB * 32));
            if (MODE == 0) {
                tma_load_2d(smem + SA_OFF_KM, &tmTabMain, tab_full, 0, 0);
                if (rem_tma) tma_load_2d(smem + SA_OFF_KR, &tmTabRem, tab_full, 64, 0);
            } else {
                for (int t = 0; t < 4; ++t) {  // rows 0..127 = R_h (K stage buffers), 128..255 = R_w (V stage buffers)
                    tma_load_2d(smem + SA_OFF_KM + t * 8192, &tmTabMain, tab_full, 0, t * 64);
                    if (rem_tma) tma_load_2d(smem + (t < 2 ? SA_OFF_KR : SA_OFF_VR) + (t & 1) * 2048, &tmTabRem, tab_full, 64, t * 64);
                }
            }
            sa_wait(t_full, 0, 2);  // tables consumed by the prologue MMAs: the ring may start
            mbar_arrive_expect_tx(&k_full[0], 8192 + RB * 2048);
            tma_load_2d(smem + SA_OFF_KM, &tmMain, &k_full[0], cKm, row0);
            if (rem_tma) tma_load_2d(smem + SA_OFF_KR, &tmRem, &k_full[0], cKr, row0);
            for (int j = 0; j < NKB; ++j) {
                if (j + 1 < NKB) {
                    const int s1 = (j + 1) & 1;
                    sa_wait(&k_empty[s1], (((j + 1) >> 1) & 1) ^ 1, 3);
                    mbar_arrive_expect_tx(&k_full[s1], 8192 + RB * 2048);
                    tma_load_2d(smem + SA_OFF_KM + s1 * 8192, &tmMain, &k_full[s1], cKm, row0 + (j + 1) * SA_BKV);
                    if (rem_tma) tma_load_2d(smem + SA_OFF_KR + s1 * 2048, &tmRem, &k_full[s1], cKr, row0 + (j + 1) * SA_BKV);
                }
                const int s = j & 1;
                sa_wait(&v_empty[s], ((j >> 1) & 1) ^ 1, 4);
                mbar_arrive_expect_tx(&v_full[s], 8192 + RB * 2048);
                tma_load_2d(smem + SA_OFF_VM + s * 8192, &tmMain, &v_full[s], cVm, row0 + j * SA_BKV);
                if (rem_tma) tma_load_2d(smem + SA_OFF_VR + s * 2048, &tmRem, &v_full[s], cVr, row0 + j * SA_BKV);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 64, false, false);
        constexpr uint32_t IDESC_S_LAST = umma_idesc_bf16(128, N_LAST, false, false);
        constexpr uint32_t IDESC_T128 = umma_idesc_bf16(128, 128, false, false);
        constexpr uint32_t IDESC_OM = umma_idesc_bf16(128, 64, false, true);   // P (TMEM) x V main (MN-major)
        constexpr uint32_t IDESC_OR = umma_idesc_bf16(128, 16, false, true);   // P (TMEM) x V rem  (MN-major)
        const uint64_t qm = umma_desc(smem_u32(smem + SA_OFF_QM), LT_SW128, 1024);
        const uint64_t qr = umma_desc(smem_u32(smem + SA_OFF_QR), LT_SW32, 256);
        const uint64_t km0 = umma_desc(smem_u32(smem + SA_OFF_KM), LT_SW128, 1024);
        const uint64_t kr0 = umma_desc(smem_u32(smem + SA_OFF_KR), LT_SW32, 256);
        const uint64_t vm0 = umma_desc(smem_u32(smem + SA_OFF_VM), LT_SW128, 1024);
        const uint64_t vr0 = umma_desc(smem_u32(smem + SA_OFF_VR), LT_SW32, 256);
        const bool rem_mma = !(p.dbg & 1);
        sa_wait(q_full, 0, 5);
        sa_wait(tab_full, 0, 6);
        tc_fence_after();
        if (elect_one()) {
            if (MODE == 0) {  // T [128 x 64] = Q [R_h (rows 0..26) ; R_w (rows 32..58)]^T
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, qm + k * 2, km0 + k * 2, IDESC_S, k != 0);
                if (rem_mma) umma_f16_ss(tmem_base, qr, kr0, IDESC_S, 1);
            } else {          // T_h [128 x 128] -> columns [0,128), T_w -> [128,256)
                const uint64_t vr_tab = umma_desc(smem_u32(smem + SA_OFF_VR), LT_SW32, 256);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, qm + k * 2, km0 + k * 2, IDESC_T128, k != 0);
                if (rem_mma) umma_f16_ss(tmem_base, qr, kr0, IDESC_T128, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + 128, qm + k * 2, vm0 + k * 2, IDESC_T128, k != 0);
                if (rem_mma) umma_f16_ss(tmem_base + 128, qr, vr_tab, IDESC_T128, 1);
            }
            umma_commit(t_full);
        }
        __syncwarp();
        sa_wait(t_cons, 0, 7);
        sa_wait(&k_full[0], 0, 8);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + SA_TM_S, qm + k * 2, km0 + k * 2, IDESC_S, k != 0);
            if (rem_mma) umma_f16_ss(tmem_base + SA_TM_S, qr, kr0, IDESC_S, 1);
            umma_commit(s_full);
            umma_commit(&k_empty[0]);
        }
        __syncwarp();
        for (int j = 0; j < NKB; ++j) {
            const int s = j & 1;
            if (j + 1 < NKB) {
                const int s1 = (j + 1) & 1;
                sa_wait(s_free, j & 1, 9);
                sa_wait(&k_full[s1], ((j + 1) >> 1) & 1, 10);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t km = km0 + s1 * (8192 >> 4), kr = kr0 + s1 * (2048 >> 4);
                    const uint32_t idesc = (j + 2 == NKB) ? IDESC_S_LAST : IDESC_S;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + SA_TM_S, qm + k * 2, km + k * 2, idesc, k != 0);
                    if (rem_mma) umma_f16_ss(tmem_base + SA_TM_S, qr, kr, idesc, 1);
                    umma_commit(s_full);
                    umma_commit(&k_empty[s1]);
                }
                __syncwarp();
            }
            sa_wait(p_ready, j & 1, 11);
            sa_wait(&v_full[s], (j >> 1) & 1, 12);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t vm = vm0 + s * (8192 >> 4), vr = vr0 + s * (2048 >> 4);
                const int steps = (j + 1 == NKB) ? N_LAST / 16 : 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < steps) {
                        // 16 keys per step: 8 packed P columns; V main rows step 16 x 128 B, V rem rows 16 x 32 B
                        umma_f16_ts(tmem_base + SA_TM_O, tmem_base + SA_TM_P + k * 8, vm + k * (2048 >> 4), IDESC_OM, (j | k) != 0);
                        if (rem_mma) umma_f16_ts(tmem_base + SA_TM_OR, tmem_base + SA_TM_P + k * 8, vr + k * (512 >> 4), IDESC_OR, (j | k) != 0);
                    }
                }
                umma_commit(&v_empty[s]);
                umma_commit(p_free);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax warps: one thread per query row =====================
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const int tok = qt * SA_BQ + r;  // token within the window / image
        const float sc = p.scale_log2;
        float bw[MODE == 0 ? SA_WS : SA_G];
        float bh[SA_WS];
        float th[MODE == 0 ? 1 : 128];  // MODE 1: this row's T_h (dynamic index per block -> local memory, L1 resident)
        int qh;
        sa_wait(t_full, 0, 13);
        tc_fence_after();
        if constexpr (MODE == 0) {
            float t[64];
            {
                uint32_t a[32];
                tmem_ld_32x32b_x32(tmem_base + lane_off, a);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = __uint_as_float(a[i]);
                tmem_ld_32x32b_x32(tmem_base + lane_off + 32, a);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) t[32 + i] = __uint_as_float(a[i]);
            }
            qh = tok / SA_WS;
            const int qw = tok - qh * SA_WS;
#pragma unroll
            for (int k = 0; k < SA_WS; ++k) {
                bh[k] = t[(qh + SA_WS - 1 - k) & 31] * SA_LOG2E;           // rows past the window (never stored) stay in range
                bw[k] = t[32 + qw + SA_WS - 1 - k] * SA_LOG2E;
            }
            th[0] = 0.f;
        } else {
            float tw[128];
#pragma unroll
            for (int cch = 0; cch < 4; ++cch) {
                uint32_t a[32];
                tmem_ld_32x32b_x32(tmem_base + lane_off + cch * 32, a);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) th[cch * 32 + i] = __uint_as_float(a[i]) * SA_LOG2E;
                tmem_ld_32x32b_x32(tmem_base + 128 + lane_off + cch * 32, a);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) tw[cch * 32 + i] = __uint_as_float(a[i]) * SA_LOG2E;
            }
            qh = tok >> 6;
            const int qw = tok & 63;
#pragma unroll
            for (int k = 0; k < SA_G; ++k) bw[k] = tw[qw + SA_G - 1 - k];
#pragma unroll
            for (int k = 0; k < SA_WS; ++k) bh[k] = 0.f;
        }
        tc_fence_before();
        mbar_arrive(t_cons);

        SmState st{0.f, 0.f};
        if constexpr (MODE == 0) {
#define SA_BLOCK(J)                                                                                                      \
    do {                                                                                                                 \
        sa_wait(s_full, (J) & 1, 14);                                                                                      \
        tc_fence_after();                                                                                                \
        sam_softmax_block<0, J>(st, J, bw, bh, 0.f, sc, tmem_base, lane_off, s_free, p_free, p_ready);                   \
    } while (0)
            SA_BLOCK(0);
            SA_BLOCK(1);
            SA_BLOCK(2);
            SA_BLOCK(3);
#undef SA_BLOCK
        } else {
            for (int j = 0; j < NKB; ++j) {
                sa_wait(s_full, j & 1, 15);
                tc_fence_after();
                const float bhj = th[qh + SA_G - 1 - j];
                sam_softmax_block<1, 0>(st, j, bw, bh, bhj, sc, tmem_base, lane_off, s_free, p_free, p_ready);
            }
        }
        // ---- epilogue: O / l -> bf16 -> this row's 160 contiguous bytes of the image-major output (pad tokens are dropped)
        sa_wait(p_free, (NKB - 1) & 1, 16);
        tc_fence_after();
        long long orow = -1;
        if constexpr (MODE == 0) {
            if (tok < SA_WTOK) {
                const int img = unit / 25, win = unit - img * 25;
                const int iy = (win / 5) * SA_WS + tok / SA_WS, ix = (win % 5) * SA_WS + tok % SA_WS;
                if (iy < SA_G && ix < SA_G) orow = (long long)img * (SA_G * SA_G) + iy * SA_G + ix;
            }
        } else {
            orow = (long long)unit * (SA_G * SA_G) + tok;
        }
        {
            // tcgen05.ld is warp-collective (.sync.aligned): EVERY lane runs the loads, only the global stores are predicated
            // (a window's pad rows and the rows past it have no output row)
            const bool live = orow >= 0;
            const float inv_l = st.l_run > 0.f ? 1.0f / st.l_run : 0.f;
            uint4* dst = reinterpret_cast<uint4*>(p.out + (live ? orow : 0) * ((long long)heads * SA_D) + head * SA_D);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t ov[32];
                tmem_ld_32x32b_x32(tmem_base + SA_TM_O + lane_off + cc * 32, ov);
                tmem_ld_wait();
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 0]) * inv_l, __uint_as_float(ov[g4 * 8 + 1]) * inv_l);
                    u.y = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 2]) * inv_l, __uint_as_float(ov[g4 * 8 + 3]) * inv_l);
                    u.z = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 4]) * inv_l, __uint_as_float(ov[g4 * 8 + 5]) * inv_l);
                    u.w = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 6]) * inv_l, __uint_as_float(ov[g4 * 8 + 7]) * inv_l);
                    if (live) dst[cc * 4 + g4] = u;
                }
            }
            uint32_t orr[16];
            tmem_ld_32x32b_x16(tmem_base + SA_TM_OR + lane_off, orr);
            tmem_ld_wait();
#pragma unroll
            for (int g4 = 0; g4 < 2; ++g4) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(orr[g4 * 8 + 0]) * inv_l, __uint_as_float(orr[g4 * 8 + 1]) * inv_l);
                u.y = pack_bf16x2(__uint_as_float(orr[g4 * 8 + 2]) * inv_l, __uint_as_float(orr[g4 * 8 + 3]) * inv_l);
                u.z = pack_bf16x2(__uint_as_float(orr[g4 * 8 + 4]) * inv_l, __uint_as_float(orr[g4 * 8 + 5]) * inv_l);
                u.w = pack_bf16x2(__uint_as_float(orr[g4 * 8 + 6]) * inv_l, __uint_as_float(orr[g4 * 8 + 7]) * inv_l);
                if (live) dst[8 + g4] = u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<SA_TMEM_COLS>(tmem_base);
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Windowed blocks, second kernel: a window has only 196 keys, so the whole score row fits tensor memory and nothing needs to be
// "flash": S [128 x 208] = Q K^T in ONE MMA pass (N = 208, 4 + 1 k-steps), a plain two-pass softmax over the row (pass 1: maximum,
// pass 2: exponentials; both re-read S from tensor memory in 32-column chunks, so a thread never holds more than a chunk), P written
// IN PLACE over the S columns it came from (P chunk c = columns [16c, 16c+16) only covers S columns that are already consumed), then
// O = P V (13 k-steps, N = 64 + 16) into columns [112, 192) of the same region, dead by then.  208 columns in all: two CTAs per SM.
// No ring, no online rescale, no per-block hand-offs: Q, K, V and the position tables are each loaded once, K / V while the prologue
// MMA (T = Q [R_h ; R_w]^T, see above) and the bias gather run.  First kernel (MODE 0 of sam_attention_kernel, four 64-key blocks):
// 13 us per CTA, 139 TFLOP/s at the ViT-H bench shape.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int SW_KROWS = 208;                              // 196 keys rounded up to 16
constexpr int SW_OFF_QM = 0;                               // [128 x 64] SW128                      16384
constexpr int SW_OFF_KM = 16384;                           // [208 x 64] SW128                      26624
constexpr int SW_OFF_VM = SW_OFF_KM + SW_KROWS * 128;      // 43008 (1024-aligned)                  26624
constexpr int SW_OFF_TM = SW_OFF_VM + SW_KROWS * 128;      // 69632: tables main [64 x 64]           8192
constexpr int SW_OFF_QR = SW_OFF_TM + 8192;                // 77824: [128 x 16] SW32                 4096
constexpr int SW_OFF_KR = SW_OFF_QR + 4096;                // 81920: [208 x 16] SW32                 6656
constexpr int SW_OFF_VR = SW_OFF_KR + SW_KROWS * 32;       // 88576 (256-aligned)                    6656
constexpr int SW_OFF_TR = SW_OFF_VR + SW_KROWS * 32;       // 95232: tables rem [64 x 16]            2048
constexpr int SW_OFF_BAR = SW_OFF_TR + 2048;               // 97280
constexpr int SW_NUM_BARS = 9;
constexpr int SW_SMEM_BYTES = SW_OFF_BAR + SW_NUM_BARS * 8 + 16;
constexpr uint32_t SW_TM_O = 112, SW_TM_OR = 176;

// t_i (log2 units) of the NCH columns of chunk CH held in a[]: scale * s + rel_h[k_h] + rel_w[k_w]; keys >= 196 give -inf
template <int CH, int NCH>
__device__ __forceinline__ void sw_chunk_logits(const uint32_t (&a)[NCH], float (&t)[NCH], float sc, const float (&bh)[SA_WS], const float (&bw)[SA_WS]) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int key = CH * 32 + i;  // compile-time after unrolling
        t[i] = key < SA_WTOK ? fmaf(__uint_as_float(a[i]), sc, bh[key / SA_WS]) + bw[key % SA_WS] : -INFINITY;
    }
}

// this row's relative-position bias registers from the prologue product T (columns [0,64) of the row's TMEM lane), log2 units
__device__ __forceinline__ void sw_read_bias(uint32_t trow, int tok, float (&bh)[SA_WS], float (&bw)[SA_WS]) {
    float t[64];  // dynamically indexed below: lives in local memory (L1)
    uint32_t a[32];
    tmem_ld_32x32b_x32(trow, a);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) t[i] = __uint_as_float(a[i]);
    tmem_ld_32x32b_x32(trow + 32, a);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) t[32 + i] = __uint_as_float(a[i]);
    const int qh = tok / SA_WS, qw = tok - qh * SA_WS;
#pragma unroll
    for (int k = 0; k < SA_WS; ++k) {
        bh[k] = t[(qh + SA_WS - 1 - k) & 31] * SA_LOG2E;  // rows past the window (never stored) stay in range
        bw[k] = t[32 + qw + SA_WS - 1 - k] * SA_LOG2E;
    }
}

// two-pass softmax of one score row held in tensor memory (208 columns at trow): P (bf16 pairs) replaces S in place; returns the row sum.
// Pass 1 only takes the maximum of the RAW scores: softmax is shift-invariant, so the shift may be any upper bound of the logits, here
// scale * max(s) + max(rel_h) + max(rel_w) (at most the spread of the position terms above the true maximum -- harmless in fp32 / bf16
// floating point), which spares the bias arithmetic of a whole pass.  Both passes load chunk c + 1 from tensor memory while chunk c is
// processed (tcgen05.wait::ld covers every outstanding load, so the prefetch is issued right after the wait for the current chunk).
__device__ __forceinline__ float sw_softmax_row(uint32_t trow, float sc, float (&bh)[SA_WS], const float (&bw)[SA_WS]) {
    uint32_t a[2][32];
    float smax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    tmem_ld_32x32b_x32(trow, a[0]);
    tmem_ld_wait();
#define SW_PASS1(CH)                                                                                  \
    do {                                                                                              \
        if ((CH) < 5) tmem_ld_32x32b_x32(trow + ((CH) + 1) * 32, a[((CH) + 1) & 1]);                  \
        else {                                                                                        \
            uint32_t(&nx)[16] = *reinterpret_cast<uint32_t(*)[16]>(&a[((CH) + 1) & 1][0]);            \
            tmem_ld_32x32b_x16(trow + 192, nx);                                                       \
        }                                                                                             \
        _Pragma("unroll") for (int i = 0; i < 32; ++i) smax[i & 3] = fmaxf(smax[i & 3], __uint_as_float(a[(CH) & 1][i])); \
        tmem_ld_wait();                                                                               \
    } while (0)
    SW_PASS1(0); SW_PASS1(1); SW_PASS1(2); SW_PASS1(3); SW_PASS1(4); SW_PASS1(5);
#undef SW_PASS1
#pragma unroll
    for (int i = 0; i < 4; ++i) smax[i & 3] = fmaxf(smax[i & 3], __uint_as_float(a[0][i]));  // chunk 6: keys 192..195 (the others are masked)
    tmem_ld_32x32b_x32(trow, a[0]);  // pass 2 starts: chunk 0 again, in flight while the shift is formed
    float bmax_h = bh[0], bmax_w = bw[0];
#pragma unroll
    for (int k = 1; k < SA_WS; ++k) {
        bmax_h = fmaxf(bmax_h, bh[k]);
        bmax_w = fmaxf(bmax_w, bw[k]);
    }
    const float mraw = fmaxf(fmaxf(smax[0], smax[1]), fmaxf(smax[2], smax[3]));
    const float shift = fmaf(mraw, sc, bmax_h + bmax_w);  // sc > 0
#pragma unroll
    for (int k = 0; k < SA_WS; ++k) bh[k] -= shift;       // fold the shift into the row term
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    tmem_ld_wait();
#define SW_PASS2(CH)                                                                                  \
    do {                                                                                              \
        if ((CH) < 5) tmem_ld_32x32b_x32(trow + ((CH) + 1) * 32, a[((CH) + 1) & 1]);                  \
        else {                                                                                        \
            uint32_t(&nx)[16] = *reinterpret_cast<uint32_t(*)[16]>(&a[((CH) + 1) & 1][0]);            \
            tmem_ld_32x32b_x16(trow + 192, nx);                                                       \
        }                                                                                             \
        float t[32];                                                                                  \
        sw_chunk_logits<CH, 32>(a[(CH) & 1], t, sc, bh, bw);                                          \
        uint32_t pk[16];                                                                              \
        _Pragma("unroll") for (int i = 0; i < 32; i += 2) {                                           \
            const float e0 = ex2f(t[i]), e1 = ex2f(t[i + 1]);                                         \
            rs[(i >> 1) & 3] += e0 + e1;                                                              \
            pk[i >> 1] = pack_bf16x2(e0, e1);                                                         \
        }                                                                                             \
        tmem_ld_wait();                                                                               \
        tmem_st_32x32b_x16(trow + (CH) * 16, pk);                                                     \
    } while (0)
    SW_PASS2(0); SW_PASS2(1); SW_PASS2(2); SW_PASS2(3); SW_PASS2(4); SW_PASS2(5);
#undef SW_PASS2
    {
        uint32_t(&a6)[16] = *reinterpret_cast<uint32_t(*)[16]>(&a[0][0]);
        float t[16];
        sw_chunk_logits<6, 16>(a6, t, sc, bh, bw);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const float e0 = ex2f(t[i]), e1 = ex2f(t[i + 1]);  // masked keys: 2^-inf = 0
            rs[(i >> 1) & 3] += e0 + e1;
            pk[i >> 1] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x32b_x4(trow + 96, pk[0], pk[1], pk[2], pk[3]);
        tmem_st_32x32b_x4(trow + 100, pk[4], pk[5], pk[6], pk[7]);
    }
    tmem_st_wait();
    return (rs[0] + rs[1]) + (rs[2] + rs[3]);
}

// O / l of one row -> bf16 -> its 160 contiguous output bytes.  tcgen05.ld is warp-collective: every lane loads, only the stores are
// predicated (a window's pad rows and the rows past it have no output row)
__device__ __forceinline__ void sw_store_row(uint32_t trow, float l_run, bool live, uint4* dst) {
    const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(trow + SW_TM_O + cc * 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 0]) * inv_l, __uint_as_float(ov[g4 * 8 + 1]) * inv_l);
            u.y = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 2]) * inv_l, __uint_as_float(ov[g4 * 8 + 3]) * inv_l);
            u.z = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 4]) * inv_l, __uint_as_float(ov[g4 * 8 + 5]) * inv_l);
            u.w = pack_bf16x2(__uint_as_float(ov[g4 * 8 + 6]) * inv_l, __uint_as_float(ov[g4 * 8 + 7]) * inv_l);
            if (live) dst[cc * 4 + g4] = u;
        }
    }
    uint32_t orr[16];
    tmem_ld_32x32b_x16(trow + SW_TM_OR, orr);
    tmem_ld_wait();
#pragma unroll
    for (int g4 = 0; g4 < 2; ++g4) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(orr[g4 * 8 + 0]) * inv_l, __uint_as_float(orr[g4 * 8 + 1]) * inv_l);
        u.y = pack_bf16x2(__uint_as_float(orr[g4 * 8 + 2]) * inv_l, __uint_as_float(orr[g4 * 8 + 3]) * inv_l);
        u.z = pack_bf16x2(__uint_as_float(orr[g4 * 8 + 4]) * inv_l, __uint_as_float(orr[g4 * 8 + 5]) * inv_l);
        u.w = pack_bf16x2(__uint_as_float(orr[g4 * 8 + 6]) * inv_l, __uint_as_float(orr[g4 * 8 + 7]) * inv_l);
        if (live) dst[8 + g4] = u;
    }
}

// image-major output row of token `tok` of window `unit` (-1: a pad token or a row past the window)
__device__ __forceinline__ long long sw_out_row(int unit, int tok) {
    if (tok >= SA_WTOK) return -1;
    const int img = unit / 25, win = unit - img * 25;
    const int iy = (win / 5) * SA_WS + tok / SA_WS, ix = (win % 5) * SA_WS + tok % SA_WS;
    return (iy < SA_G && ix < SA_G) ? (long long)img * (SA_G * SA_G) + iy * SA_G + ix : -1;
}

__global__ void __launch_bounds__(192, 2)
sam_window_attention_kernel(const __grid_constant__ CUtensorMap tmQm, const __grid_constant__ CUtensorMap tmQr, const __grid_constant__ CUtensorMap tmKVm,
                            const __grid_constant__ CUtensorMap tmKVr, const __grid_constant__ CUtensorMap tmTabMain,
                            const __grid_constant__ CUtensorMap tmTabRem, const SamAttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SW_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* tab_full = bars + 1;
    uint64_t* k_full = bars + 2;
    uint64_t* v_full = bars + 3;
    uint64_t* t_full = bars + 4;
    uint64_t* t_cons = bars + 5;
    uint64_t* s_full = bars + 6;
    uint64_t* p_ready = bars + 7;
    uint64_t* o_full = bars + 8;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + SW_NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, unit = blockIdx.z;
    const int heads = p.heads;
    const int row0 = unit * SA_WTOK;
    const int cQm = head * 64, cKm = heads * 64 + head * 64, cVm = 2 * heads * 64 + head * 64;
    const int cQr = 3 * heads * 64 + head * 16, cKr = cQr + heads * 16, cVr = cKr + heads * 16;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQm);
        tma_prefetch_desc(&tmKVm);
        mbar_init(q_full, 1);
        mbar_init(tab_full, 1);
        mbar_init(k_full, 1);
        mbar_init(v_full, 1);
        mbar_init(t_full, 1);
        mbar_init(t_cons, 128);
        mbar_init(s_full, 1);
        mbar_init(p_ready, 128);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<SA_TMEM_COLS>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer: everything is loaded exactly once
            mbar_arrive_expect_tx(q_full, 16384 + 4096);
            tma_load_2d(smem + SW_OFF_QM, &tmQm, q_full, cQm, row0 + qt * SA_BQ);
            tma_load_2d(smem + SW_OFF_QR, &tmQr, q_full, cQr, row0 + qt * SA_BQ);
            mbar_arrive_expect_tx(tab_full, 8192 + 2048);
            tma_load_2d(smem + SW_OFF_TM, &tmTabMain, tab_full, 0, 0);
            tma_load_2d(smem + SW_OFF_TR, &tmTabRem, tab_full, 64, 0);
            mbar_arrive_expect_tx(k_full, SW_KROWS * 160);
            tma_load_2d(smem + SW_OFF_KM, &tmKVm, k_full, cKm, row0);
            tma_load_2d(smem + SW_OFF_KR, &tmKVr, k_full, cKr, row0);
            mbar_arrive_expect_tx(v_full, SW_KROWS * 160);
            tma_load_2d(smem + SW_OFF_VM, &tmKVm, v_full, cVm, row0);
            tma_load_2d(smem + SW_OFF_VR, &tmKVr, v_full, cVr, row0);
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        constexpr uint32_t IDESC_T = umma_idesc_bf16(128, 64, false, false);
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, SW_KROWS, false, false);
        constexpr uint32_t IDESC_OM = umma_idesc_bf16(128, 64, false, true);
        constexpr uint32_t IDESC_OR = umma_idesc_bf16(128, 16, false, true);
        const uint64_t qm = umma_desc(smem_u32(smem + SW_OFF_QM), LT_SW128, 1024), qr = umma_desc(smem_u32(smem + SW_OFF_QR), LT_SW32, 256);
        const uint64_t km = umma_desc(smem_u32(smem + SW_OFF_KM), LT_SW128, 1024), kr = umma_desc(smem_u32(smem + SW_OFF_KR), LT_SW32, 256);
        const uint64_t vm = umma_desc(smem_u32(smem + SW_OFF_VM), LT_SW128, 1024), vr = umma_desc(smem_u32(smem + SW_OFF_VR), LT_SW32, 256);
        const uint64_t tm = umma_desc(smem_u32(smem + SW_OFF_TM), LT_SW128, 1024), tr = umma_desc(smem_u32(smem + SW_OFF_TR), LT_SW32, 256);
        mbar_wait(q_full, 0);
        mbar_wait(tab_full, 0);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, qm + k * 2, tm + k * 2, IDESC_T, k != 0);
            umma_f16_ss(tmem_base, qr, tr, IDESC_T, 1);
            umma_commit(t_full);
        }
        __syncwarp();
        mbar_wait(t_cons, 0);
        mbar_wait(k_full, 0);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, qm + k * 2, km + k * 2, IDESC_S, k != 0);
            umma_f16_ss(tmem_base, qr, kr, IDESC_S, 1);
            umma_commit(s_full);
        }
        __syncwarp();
        mbar_wait(p_ready, 0);
        mbar_wait(v_full, 0);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < SW_KROWS / 16; ++k) {
                umma_f16_ts(tmem_base + SW_TM_O, tmem_base + k * 8, vm + k * (2048 >> 4), IDESC_OM, k != 0);
                umma_f16_ts(tmem_base + SW_TM_OR, tmem_base + k * 8, vr + k * (512 >> 4), IDESC_OR, k != 0);
            }
            umma_commit(o_full);
        }
        __syncwarp();
    } else {
        // ---- softmax warps: one thread per query row
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t trow = tmem_base + lane_off;
        const int tok = qt * SA_BQ + r;
        const float sc = p.scale_log2;
        float bw[SA_WS], bh[SA_WS];
        mbar_wait(t_full, 0);
        tc_fence_after();
        sw_read_bias(trow, tok, bh, bw);
        tc_fence_before();
        mbar_arrive(t_cons);
        mbar_wait(s_full, 0);
        tc_fence_after();
        const float l_run = sw_softmax_row(trow, sc, bh, bw);
        tc_fence_before();
        mbar_arrive(p_ready);
        mbar_wait(o_full, 0);
        tc_fence_after();
        const long long orow = sw_out_row(unit, tok);
        sw_store_row(trow, l_run, orow >= 0, reinterpret_cast<uint4*>(p.out + (orow >= 0 ? orow : 0) * ((long long)heads * SA_D) + head * SA_D));
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<SA_TMEM_COLS>(tmem_base);
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Windowed blocks, persistent kernel: ONE CTA per SM walks the (window, head) items; it owns BOTH 128-query tiles of an item (two
// softmax warpgroups, one tensor-memory region of 208 columns each) so K / V are loaded once per item, and the producer loads the next
// item's Q / K / V (105 KB, second shared-memory stage) while the current one is computed: allocation, barrier set-up, the position
// tables and the TMA round trips -- what made the one-CTA-per-tile kernel 11 us per CTA -- are paid once or hidden.
//   warp 0 producer | warp 1 MMA issuer | warps 2-5 softmax of tile A (tokens 0..127) | warps 6-9 softmax of tile B (tokens 128..195)
// Per item:  T_A, T_B (prologue products for the position bias)  ->  S_A, S_B  ->  two-pass softmax, P in place  ->  O_A, O_B  -> store.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int SP_STAGE_BYTES = 32768 + 8192 + 2 * (SW_KROWS * 128) + 2 * (SW_KROWS * 32);  // 107520 per (window, head)
constexpr int SP_OFF_QM = 0, SP_OFF_KM = 32768, SP_OFF_VM = SP_OFF_KM + SW_KROWS * 128, SP_OFF_QR = SP_OFF_VM + SW_KROWS * 128;
constexpr int SP_OFF_KR = SP_OFF_QR + 8192, SP_OFF_VR = SP_OFF_KR + SW_KROWS * 32;
static_assert(SP_OFF_VR + SW_KROWS * 32 == SP_STAGE_BYTES && SP_STAGE_BYTES % 1024 == 0, "stage layout");
constexpr int SP_OFF_TM = 2 * SP_STAGE_BYTES;       // tables main [64 x 64]
constexpr int SP_OFF_TR = SP_OFF_TM + 8192;         // tables rem  [64 x 16]
constexpr int SP_OFF_BAR = SP_OFF_TR + 2048;
constexpr int SP_NUM_BARS = 24;
constexpr int SP_SMEM_BYTES = SP_OFF_BAR + SP_NUM_BARS * 8 + 16;
constexpr int SP_THREADS = 320;
constexpr uint32_t SP_TM_TILE = 256;                // tensor-memory region of tile B starts here

__global__ void __launch_bounds__(SP_THREADS, 1)
sam_window_persist_kernel(const __grid_constant__ CUtensorMap tmQm, const __grid_constant__ CUtensorMap tmQr, const __grid_constant__ CUtensorMap tmKVm,
                          const __grid_constant__ CUtensorMap tmKVr, const __grid_constant__ CUtensorMap tmTabMain,
                          const __grid_constant__ CUtensorMap tmTabRem, const SamAttnParams p, int n_items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SP_OFF_BAR);
    uint64_t* q_full = bars + 0;      // [2] per stage
    uint64_t* k_full = bars + 2;      // [2]
    uint64_t* v_full = bars + 4;      // [2]
    uint64_t* stage_free = bars + 6;  // [2] every MMA that reads the stage has completed
    uint64_t* tab_full = bars + 8;
    uint64_t* t_full = bars + 9;      // [2] per tile
    uint64_t* t_cons = bars + 11;     // [2]
    uint64_t* s_full = bars + 13;     // [2]
    uint64_t* p_ready = bars + 15;    // [2]
    uint64_t* o_full = bars + 17;     // [2]
    uint64_t* o_read = bars + 19;     // [2] the tile's O has been read: its tensor-memory region may take the next item
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + SP_NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int heads = p.heads;
    const int stride = gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQm);
        tma_prefetch_desc(&tmQr);
        tma_prefetch_desc(&tmKVm);
        tma_prefetch_desc(&tmKVr);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&stage_free[i], 1);
            mbar_init(&t_full[i], 1);
            mbar_init(&t_cons[i], 128);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_ready[i], 128);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_read[i], 128);
        }
        mbar_init(tab_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            mbar_arrive_expect_tx(tab_full, 8192 + 2048);
            tma_load_2d(smem + SP_OFF_TM, &tmTabMain, tab_full, 0, 0);
            tma_load_2d(smem + SP_OFF_TR, &tmTabRem, tab_full, 64, 0);
            int it = 0;
            for (int w = blockIdx.x; w < n_items; w += stride, ++it) {
                const int s = it & 1;
                const int head = w % heads, unit = w / heads;
                const int row0 = unit * SA_WTOK;
                const int cQm = head * 64, cKm = heads * 64 + head * 64, cVm = 2 * heads * 64 + head * 64;
                const int cQr = 3 * heads * 64 + head * 16, cKr = cQr + heads * 16, cVr = cKr + heads * 16;
                uint8_t* st = smem + s * SP_STAGE_BYTES;
                if (it >= 2) mbar_wait_relaxed(&stage_free[s], ((it >> 1) - 1) & 1);
                mbar_arrive_expect_tx(&q_full[s], 32768 + 8192);
                tma_load_2d(st + SP_OFF_QM, &tmQm, &q_full[s], cQm, row0);
                tma_load_2d(st + SP_OFF_QM + 16384, &tmQm, &q_full[s], cQm, row0 + 128);
                tma_load_2d(st + SP_OFF_QR, &tmQr, &q_full[s], cQr, row0);
                tma_load_2d(st + SP_OFF_QR + 4096, &tmQr, &q_full[s], cQr, row0 + 128);
                mbar_arrive_expect_tx(&k_full[s], SW_KROWS * 160);
                tma_load_2d(st + SP_OFF_KM, &tmKVm, &k_full[s], cKm, row0);
                tma_load_2d(st + SP_OFF_KR, &tmKVr, &k_full[s], cKr, row0);
                mbar_arrive_expect_tx(&v_full[s], SW_KROWS * 160);
                tma_load_2d(st + SP_OFF_VM, &tmKVm, &v_full[s], cVm, row0);
                tma_load_2d(st + SP_OFF_VR, &tmKVr, &v_full[s], cVr, row0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC_T = umma_idesc_bf16(128, 64, false, false);
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, SW_KROWS, false, false);
        constexpr uint32_t IDESC_OM = umma_idesc_bf16(128, 64, false, true);
        constexpr uint32_t IDESC_OR = umma_idesc_bf16(128, 16, false, true);
        const uint64_t tm = umma_desc(smem_u32(smem + SP_OFF_TM), LT_SW128, 1024), tr = umma_desc(smem_u32(smem + SP_OFF_TR), LT_SW32, 256);
        mbar_wait(tab_full, 0);
        int it = 0;
        for (int w = blockIdx.x; w < n_items; w += stride, ++it) {
            const int s = it & 1, ph = it & 1, sp = (it >> 1) & 1;
            const uint32_t sb = smem_u32(smem + s * SP_STAGE_BYTES);
            const uint64_t km = umma_desc(sb + SP_OFF_KM, LT_SW128, 1024), kr = umma_desc(sb + SP_OFF_KR, LT_SW32, 256);
            const uint64_t vm = umma_desc(sb + SP_OFF_VM, LT_SW128, 1024), vr = umma_desc(sb + SP_OFF_VR, LT_SW32, 256);
            mbar_wait(&q_full[s], sp);
#pragma unroll
            for (int x = 0; x < 2; ++x) {  // prologue products T_x = Q_x [R_h ; R_w]^T
                const uint64_t qm = umma_desc(sb + SP_OFF_QM + x * 16384, LT_SW128, 1024), qr = umma_desc(sb + SP_OFF_QR + x * 4096, LT_SW32, 256);
                if (it > 0) mbar_wait(&o_read[x], (it - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + x * SP_TM_TILE, qm + k * 2, tm + k * 2, IDESC_T, k != 0);
                    umma_f16_ss(tmem_base + x * SP_TM_TILE, qr, tr, IDESC_T, 1);
                    umma_commit(&t_full[x]);
                }
                __syncwarp();
            }
            mbar_wait(&k_full[s], sp);
#pragma unroll
            for (int x = 0; x < 2; ++x) {  // S_x = Q_x K^T, all 208 key columns at once
                const uint64_t qm = umma_desc(sb + SP_OFF_QM + x * 16384, LT_SW128, 1024), qr = umma_desc(sb + SP_OFF_QR + x * 4096, LT_SW32, 256);
                mbar_wait(&t_cons[x], ph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + x * SP_TM_TILE, qm + k * 2, km + k * 2, IDESC_S, k != 0);
                    umma_f16_ss(tmem_base + x * SP_TM_TILE, qr, kr, IDESC_S, 1);
                    umma_commit(&s_full[x]);
                }
                __syncwarp();
            }
            mbar_wait(&v_full[s], sp);
#pragma unroll
            for (int x = 0; x < 2; ++x) {  // O_x = P_x V
                mbar_wait(&p_ready[x], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t tb = tmem_base + x * SP_TM_TILE;
#pragma unroll
                    for (int k = 0; k < SW_KROWS / 16; ++k) {
                        umma_f16_ts(tb + SW_TM_O, tb + k * 8, vm + k * (2048 >> 4), IDESC_OM, k != 0);
                        umma_f16_ts(tb + SW_TM_OR, tb + k * 8, vr + k * (512 >> 4), IDESC_OR, k != 0);
                    }
                    umma_commit(&o_full[x]);
                    if (x == 1) umma_commit(&stage_free[s]);  // every MMA reading this stage has been issued before this commit
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== softmax warpgroups: tile x, one thread per query row =====================
        const int x = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t trow = tmem_base + x * SP_TM_TILE + (static_cast<uint32_t>(quarter * 32) << 16);
        const int tok = x * SA_BQ + r;
        const float sc = p.scale_log2;
        int it = 0;
        for (int w = blockIdx.x; w < n_items; w += stride, ++it) {
            const int ph = it & 1;
            const int head = w % heads, unit = w / heads;
            float bw[SA_WS], bh[SA_WS];
            mbar_wait(&t_full[x], ph);
            tc_fence_after();
            sw_read_bias(trow, tok, bh, bw);
            tc_fence_before();
            mbar_arrive(&t_cons[x]);
            mbar_wait(&s_full[x], ph);
            tc_fence_after();
            const float l_run = sw_softmax_row(trow, sc, bh, bw);
            tc_fence_before();
            mbar_arrive(&p_ready[x]);
            mbar_wait(&o_full[x], ph);
            tc_fence_after();
            const long long orow = sw_out_row(unit, tok);
            sw_store_row(trow, l_run, orow >= 0, reinterpret_cast<uint4*>(p.out + (orow >= 0 ? orow : 0) * ((long long)heads * SA_D) + head * SA_D));
            tc_fence_before();
            mbar_arrive(&o_read[x]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace
}  // namespace wg

extern "C" int wg_sam_attention(const void* qkv, const void* rel_table, void* out, int n_images, int mode, int heads, void* stream_) {
    using namespace wg;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(qkv && rel_table && out, "wg_sam_attention: null pointer");
    WG_REQUIRE(n_images > 0 && heads > 0 && heads <= 65535 && (mode == 0 || mode == 1), "wg_sam_attention: bad arguments (images %d heads %d mode %d)",
               n_images, heads, mode);
    WG_REQUIRE((long long)n_images * 25 <= 65535, "wg_sam_attention: at most 2621 images per call");
    if (!device_is_sm100()) {
        set_error("wg_sam_attention: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const uint64_t ncols = (uint64_t)heads * 3 * SA_D;
    const uint64_t rows = (uint64_t)n_images * (mode == 0 ? 25 * SA_WTOK : SA_G * SA_G);
    CUtensorMap tmMain, tmRem, tmTabMain, tmTabRem;
    {
        uint64_t dims[2] = {ncols, rows};
        uint64_t strides[1] = {ncols * 2};
        uint32_t box[2] = {64, 64};
        WG_TRY(make_tensor_map_sw(&tmMain, qkv, 2, 2, dims, strides, box, 128));
        uint32_t boxr[2] = {16, 64};
        WG_TRY(make_tensor_map_sw(&tmRem, qkv, 2, 2, dims, strides, boxr, 32));
    }
    {
        uint64_t dims[2] = {SA_D, (uint64_t)(mode == 0 ? 64 : 256)};
        uint64_t strides[1] = {SA_D * 2};
        uint32_t box[2] = {64, 64};
        WG_TRY(make_tensor_map_sw(&tmTabMain, rel_table, 2, 2, dims, strides, box, 128));
        uint32_t boxr[2] = {16, 64};
        WG_TRY(make_tensor_map_sw(&tmTabRem, rel_table, 2, 2, dims, strides, boxr, 32));
    }
    SamAttnParams p;
    static const int dbg = [] { const char* e = getenv("WG_SAM_DBG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg;
    if (dbg & 12) WG_CHECK_CUDA(cudaMemcpyToSymbol(g_sa_dbg, &dbg, sizeof(int)));
    p.heads = heads;
    p.scale_log2 = 0.11180339887498948f * SA_LOG2E;  // 80^-0.5
    p.out = static_cast<__nv_bfloat16*>(out);
    const double keys = mode == 0 ? SA_WTOK : SA_G * SA_G;
    const double units = (double)n_images * (mode == 0 ? 25 : 1);
    Prof prof(mode == 0 ? "sam_attention_window" : "sam_attention_global", stream, 4.0 * units * heads * keys * keys * SA_D,
              2.0 * 4.0 * units * keys * heads * SA_D);
    // windowed blocks: 0 = persistent two-tile kernel (default), 1 = one CTA per query tile, 2 = the first version (four 64-key flash blocks)
    static const int window_kernel = [] { const char* e = getenv("WG_SAM_WINDOW_KERNEL"); return e ? atoi(e) : 0; }();
    if (mode == 0 && window_kernel != 2) {
        // one-pass window kernel: Q tile / whole-window K, V boxes
        CUtensorMap tmQm, tmQr, tmKVm, tmKVr;
        uint64_t dims[2] = {ncols, rows};
        uint64_t strides[1] = {ncols * 2};
        uint32_t bq[2] = {64, 128}, bqr[2] = {16, 128}, bk[2] = {64, SW_KROWS}, bkr[2] = {16, SW_KROWS};
        WG_TRY(make_tensor_map_sw(&tmQm, qkv, 2, 2, dims, strides, bq, 128));
        WG_TRY(make_tensor_map_sw(&tmQr, qkv, 2, 2, dims, strides, bqr, 32));
        WG_TRY(make_tensor_map_sw(&tmKVm, qkv, 2, 2, dims, strides, bk, 128));
        WG_TRY(make_tensor_map_sw(&tmKVr, qkv, 2, 2, dims, strides, bkr, 32));
        if (window_kernel == 0) {
            const int n_items = n_images * 25 * heads;
            const int sms = device_sm_count();
            WG_SMEM_OPT_IN(sam_window_persist_kernel, SP_SMEM_BYTES);
            sam_window_persist_kernel<<<n_items < sms ? n_items : sms, SP_THREADS, SP_SMEM_BYTES, stream>>>(tmQm, tmQr, tmKVm, tmKVr, tmTabMain, tmTabRem, p,
                                                                                                         n_items);
        } else {
            WG_SMEM_OPT_IN(sam_window_attention_kernel, SW_SMEM_BYTES);
            sam_window_attention_kernel<<<dim3(2, heads, n_images * 25), 192, SW_SMEM_BYTES, stream>>>(tmQm, tmQr, tmKVm, tmKVr, tmTabMain, tmTabRem, p);
        }
    } else if (mode == 0) {
        WG_SMEM_OPT_IN(sam_attention_kernel<0>, SA_SMEM_BYTES);
        sam_attention_kernel<0><<<dim3(2, heads, n_images * 25), 192, SA_SMEM_BYTES, stream>>>(tmMain, tmRem, tmTabMain, tmTabRem, p);
    } else {
        WG_SMEM_OPT_IN(sam_attention_kernel<1>, SA_SMEM_BYTES);
        sam_attention_kernel<1><<<dim3(SA_G * SA_G / SA_BQ, heads, n_images), 192, SA_SMEM_BYTES, stream>>>(tmMain, tmRem, tmTabMain, tmTabRem, p);
    }
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

// bring-up only (not part of the ABI): host-mapped progress array for WG_SAM_DBG & 8
extern "C" __attribute__((visibility("default"))) int wg_debug_sam_marks(void* p) {
    return (int)cudaMemcpyToSymbol(wg::g_sa_marks, &p, sizeof(void*));
}
