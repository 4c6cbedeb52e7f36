// CTA-pair (cta_group::2) variant of the persistent tcgen05 GEMM:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
// Two CTAs on the two SMs of a TPC (cluster of 2) compute one 256 x 256 tile together.  Each CTA stages only ITS half of
// both operands -- A rows [m0 + 128*rank, +128) and W rows [n0 + 128*rank, +128) -- and one thread of the leader CTA
// issues tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16): the tensor cores of both SMs read the two W halves from both
// CTAs' shared memory.  Against the 1-CTA kernel (128 x 256 tile, 16 KB A + 32 KB W per k-block per SM) a k-block costs each
// SM 16 KB + 16 KB of TMA fill and tensor-core smem reads: a third less shared-memory traffic, six pipeline stages instead of
// four, and half the L2 -> smem W traffic.  Accumulators: each CTA's TMEM holds its own 128 rows x 256 columns, double
// buffered (512 columns), and its own 8 epilogue warps drain them (shared code: epilogue_tile in gemm_common.cuh).
//
// Barriers (all mbarriers; "L" = lives in / is waited on by the leader only):
//   full[s]  (L, count 1)  : the leader's producer arrives with expect_tx(64 KB = both CTAs); both CTAs' TMA loads complete_tx on it
//   empty[s] (per CTA)     : tcgen05.commit multicast from the leader after the MMAs that read stage s
//   tmem_full[a] (per CTA) : tcgen05.commit multicast after the last k-block of a tile
//   tmem_empty[a] (L, 2)   : one arrival per CTA, forwarded by its warp 3 once the CTA's 8 epilogue warps arrived on epi_done[a] (local):
//                            the remote arrive costs ~1100-1800 cycles, which the epilogue warps used to pay once per tile
#include "gemm_common.cuh"
#include "internal.h"

namespace wg {

namespace {

using namespace gemm_detail;

constexpr int BN2 = 256;                      // pair tile: 256 (M, two CTAs) x 256 (N)
constexpr int B_HALF_BYTES = 128 * BK * 2;    // this CTA's half of the W tile

#ifdef GEMM2_TRACE
__device__ long long g_gemm2_trace[3][64][4];  // role (0 MMA, 1 producer rank 0, 2 producer rank 1) x k-block x stamp
#define TR2(role, i, slot) do { if (trace_on && (i) < 64) g_gemm2_trace[role][i][slot] = clock64(); } while (0)
__device__ long long g_gemm2_epi[16][8];  // per tile of pair 5: MMA warp {0 wait tmem_empty, 1 got it, 2 committed}, epilogue warp 4 of
                                          // rank 0 {3 wait tmem_full, 4 got it, 5 tile drained}
#define TRE(i, slot) do { if (trace_pair && (i) < 16) g_gemm2_epi[i][slot] = clock64(); } while (0)
#else
#define TR2(role, i, slot) do { } while (0)
#define TRE(i, slot) do { } while (0)
#endif

template <int STAGES, int EPI = WG_OUT_BF16>
struct SmemLayout2 {
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = STAGES * A_STAGE_BYTES;
    static constexpr int OFF_C = OFF_B + STAGES * B_HALF_BYTES;
    // WG_OUT_F32_TMA: two boxes (ping-pong) for each of the 8 epilogue warps, and one mbarrier per box for the residual loads
    static constexpr int C_BYTES = EPI == WG_OUT_F32_TMA ? 8 * F32_NBOX * F32_WARP_BUF_BYTES : 2 * C_BUF_BYTES;
    static constexpr int OFF_BAR = OFF_C + C_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 6 + (EPI == WG_OUT_F32_TMA ? 8 * F32_NBOX : 0);
    static constexpr int TOTAL = OFF_BAR + NUM_BARS * 8 + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024B alignment
};

template <int STAGES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
    using L = SmemLayout2<STAGES, EPI>;
    constexpr bool F32_EPI = EPI == WG_OUT_F32 || EPI == WG_OUT_F32_LN;  // per-thread global stores (no TMA store to drain)
    constexpr int TMEM_COLS = 2 * BN2;
    constexpr uint32_t IDESC = umma_idesc_bf16(2 * BM, BN2, false, false);
    constexpr uint32_t STAGE_TX_BYTES = A_STAGE_BYTES + B_HALF_BYTES;  // per CTA

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint64_t* epi_done = bars + 2 * STAGES + 4;  // [2] local: this CTA's 8 epilogue warps have drained accumulator a
    uint64_t* resid_bar = bars + 2 * STAGES + 6;  // WG_OUT_F32_TMA: [8 warps][F32_NBOX boxes] residual box has landed
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
#ifdef GEMM2_TRACE
    const bool trace_pair = pair == 5;
#endif

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (!F32_EPI) tma_prefetch_desc(&tmC);
        if (EPI == WG_OUT_F32_TMA) tma_prefetch_desc(&tmR);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 2);   // one (forwarded) arrival per CTA of the pair
            mbar_init(&epi_done[a], 8);     // one arrival per epilogue warp of this CTA
        }
        if (EPI == WG_OUT_F32_TMA) {
            for (int i = 0; i < 8 * F32_NBOX; ++i) mbar_init(&resid_bar[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_cg2<TMEM_COLS>(tmem_ptr_smem);
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer (both CTAs: own A rows, own half of the W tile) =====================
        int stage = 0;
        uint32_t phase = 0;
        int titer = 0;
        const uint32_t full_leader0 = mapa_shared(smem_u32(&full_bar[0]), 0);
        for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++titer) {
            const int m0 = (tile / p.num_n_tiles) * (2 * BM) + rank * BM;
            const int n0 = (tile % p.num_n_tiles) * BN2 + rank * 128;
#ifdef GEMM2_TRACE
            const bool trace_on = trace_pair && (titer == 2 || titer == 3);
            const int tb = (titer - 2) * p.num_k_blocks;
#endif
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                TR2(1 + rank, tb + kb, 0);
                mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
                TR2(1 + rank, tb + kb, 1);
                // only the leader arms the barrier, with the bytes of BOTH CTAs; the peer's loads may complete_tx before
                // that (the count goes negative transiently -- the phase cannot complete before the leader's arrival)
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_TX_BYTES);
                int ka = kb * BK;
                if (p.a_k_wrap > 0 && ka >= p.a_k_wrap) ka -= p.a_k_wrap;  // K' <= 2 * wrap by construction
                const uint32_t full_leader = full_leader0 + stage * 8;
                tma_load_2d_cg2(smem + L::OFF_A + stage * A_STAGE_BYTES, &tmA, full_leader, ka, m0);
                TR2(1 + rank, tb + kb, 3);
                tma_load_2d_cg2(smem + L::OFF_B + stage * B_HALF_BYTES, &tmB, full_leader, kb * BK, n0);
                TR2(1 + rank, tb + kb, 2);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ===================== MMA issuer (leader CTA only) =====================
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            TRE(iter, 0);
            mbar_wait_relaxed(&tmem_empty[acc], acc_phase ^ 1);
            TRE(iter, 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN2;
#ifdef GEMM2_TRACE
            const bool trace_on = trace_pair && (iter == 2 || iter == 3);
            const int tb = (iter - 2) * p.num_k_blocks;
#endif
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                TR2(0, tb + kb, 0);
                mbar_wait_relaxed(&full_bar[stage], phase);
                TR2(0, tb + kb, 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + L::OFF_A + stage * A_STAGE_BYTES);
                const uint32_t b_addr = smem_u32(smem + L::OFF_B + stage * B_HALF_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    umma_f16_ss_cg2(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), IDESC,
                                    (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit_cg2(&empty_bar[stage], 0b11);  // frees the smem slot in both CTAs
                TR2(0, tb + kb, 2);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit_cg2(&tmem_full[acc], 0b11);  // accumulators of both CTAs complete -> both epilogues
            TRE(iter, 2);
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 2 groups x 4 warps per CTA, on this CTA's 128 rows =====================
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int epi_tid = threadIdx.x - (4 + 4 * grp) * 32;  // index inside the group
        uint8_t* cbufs = smem + L::OFF_C;
        int iter = 0;
        // WG_OUT_F32_TMA: this warp's two boxes, their barriers, its chunk counter; the first residual box of the first tile is requested here
        uint8_t* wbuf = cbufs + (grp * 4 + q) * (F32_NBOX * F32_WARP_BUF_BYTES);
        uint64_t* rbar = resid_bar + (grp * 4 + q) * F32_NBOX;
        uint32_t cc = 0;
        if (EPI == WG_OUT_F32_TMA && p.resid_f32 != nullptr && pair < p.num_tiles && lane == 0) {
            const F32TmaNext none = {0, 0, false};
            for (int i = 0; i < F32_NBOX - 1; ++i)
                f32_tma_request<BN2>(tmR, wbuf, rbar, i, i, (pair / p.num_n_tiles) * (2 * BM) + (int)rank * BM, (pair % p.num_n_tiles) * BN2, q, grp, none);
        }
        for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int m0 = (tile / p.num_n_tiles) * (2 * BM) + rank * BM;
            const int n0 = (tile % p.num_n_tiles) * BN2;
#ifdef GEMM2_TRACE
            const bool tre = trace_pair && rank == 0 && warp == 4 && lane == 0;
            if (tre) TRE(iter, 3);
#endif
            mbar_wait_relaxed(&tmem_full[acc], acc_phase);
#ifdef GEMM2_TRACE
            if (tre) TRE(iter, 4);
#endif
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN2;

            LnRowStats ln;
            if constexpr (EPI == WG_OUT_F32_TMA) {
                const int nt_ = tile + num_pairs;
                const F32TmaNext nx = {(nt_ / p.num_n_tiles) * (2 * BM) + (int)rank * BM, (nt_ % p.num_n_tiles) * BN2, nt_ < p.num_tiles};
                epilogue_tile_f32_tma<BN2>(p, tmC, tmR, wbuf, rbar, cc, taddr, m0, n0, q, grp, lane, nx);
            } else {
                epilogue_tile<BN2, EPI>(p, tmC, cbufs, taddr, m0, n0, q, grp, lane, epi_tid, &ln);
            }
            tc_fence_before();
            __syncwarp();
#ifdef GEMM2_TRACE
            if (tre) TRE(iter, 5);
#endif
            if (lane == 0) mbar_arrive(&epi_done[acc]);
            // fused LayerNorm: the accumulator is already back with the MMA warp; the wait for the row block's other tiles and the
            // normalising pass run beside the next tile's main loop
            if constexpr (EPI == WG_OUT_F32_LN) ln_pass2<BN2>(p, m0, n0, q, grp, lane, epi_tid, ln);
        }
        if (!F32_EPI && EPI != WG_OUT_F32_TMA && epi_tid == 0) tma_store_wait_all<0>();
        if (EPI == WG_OUT_F32_TMA && lane == 0) tma_store_wait_all<0>();  // every warp's elected lane drains its own stores
    } else if (warp == 3 && lane == 0) {
        // ===================== forwards "accumulator drained" to the leader's MMA warp =====================
        int iter = 0;
        for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++iter) {
            const int acc = iter & 1;
            mbar_wait_relaxed(&epi_done[acc], (iter >> 1) & 1);
            mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
        }
    }

    // neither CTA may leave (or free its TMEM) while the peer can still signal its barriers or read its shared memory
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
    }
}

template <int STAGES, int EPI>
int launch2(const wg_gemm_args* a, cudaStream_t stream, const GemmLnFuse* fuse = nullptr) {
    using L = SmemLayout2<STAGES, EPI>;
    constexpr bool F32_EPI = EPI == WG_OUT_F32 || EPI == WG_OUT_F32_LN || EPI == WG_OUT_F32_TMA;
    CUtensorMap tmA, tmB, tmC, tmR;
    WG_TRY(make_tmap_2d_bf16(&tmA, a->A, a->M, a->a_k_wrap > 0 ? a->a_k_wrap : a->K, a->lda, BM, BK));
    WG_TRY(make_tmap_2d_bf16(&tmB, a->W, a->N, a->K, a->ldw, 128, BK));
    if (!F32_EPI) {
        WG_TRY(make_tmap_2d_bf16(&tmC, a->out, a->M, a->split_out ? 2 * a->N : a->N, a->ldo, BM, 64));
    } else {
        tmC = tmA;
    }
    tmR = tmA;
    if (EPI == WG_OUT_F32_TMA) {  // fp32 [M, N] as boxes of 32 rows x 32 columns (128-byte rows, 128B swizzle)
        uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->M};
        uint64_t strides[1] = {(uint64_t)a->ldo * 4};
        uint32_t box[2] = {32, 32};
        WG_TRY(make_tensor_map(&tmC, a->out, 4, 2, dims, strides, box));
        if (a->resid) WG_TRY(make_tensor_map(&tmR, a->resid, 4, 2, dims, strides, box));
    }
    GemmParams p;
    p.M = a->M;
    p.N = a->N;
    p.K = a->K;
    p.num_n_tiles = (a->N + BN2 - 1) / BN2;
    p.num_tiles = ((a->M + 2 * BM - 1) / (2 * BM)) * p.num_n_tiles;
    p.num_k_blocks = (a->K + BK - 1) / BK;
    p.bias = a->bias;
    p.bias_period = a->bias_period;
    p.act = a->act;
    p.out_f32 = F32_EPI ? static_cast<float*>(a->out) : nullptr;
    p.resid_f32 = F32_EPI ? static_cast<const float*>(a->resid) : nullptr;
    p.resid_bf16 = (EPI == WG_OUT_BF16_LN) ? static_cast<const __nv_bfloat16*>(a->resid) : nullptr;
    p.ldo = a->ldo;
    p.ln_gamma = a->ln_gamma;
    p.ln_beta = a->ln_beta;
    p.ln_eps = a->ln_eps;
    p.a_k_wrap = a->a_k_wrap;
    p.split_out = !F32_EPI ? a->split_out : 0;
    p.ln_out = nullptr; p.ln_stats = nullptr; p.ln_flags = nullptr; p.ln_epoch = 0;
    if (EPI == WG_OUT_F32_LN) {
        p.ln_out = static_cast<__nv_bfloat16*>(fuse->ln_out);
        p.ln_stats = fuse->stats;
        p.ln_flags = fuse->flags;
        p.ln_epoch = fuse->epoch;
        p.ln_gamma = fuse->gamma;
        p.ln_beta = fuse->beta;
        p.ln_eps = fuse->eps;
    }

    auto kern = gemm2_bf16_kernel<STAGES, EPI>;
    WG_SMEM_OPT_IN(kern, L::DYN_BYTES);  // per instantiation and device
    int max_pairs = device_sm_count() / 2;
    // fused LayerNorm: the n-tiles of a row block must run in the SAME wave (a tile that waits for the next wave's tiles stalls its
    // pair's epilogue for a whole main loop, and the delay spreads through the groups): a multiple of the n-tile count of pairs
    if (EPI == WG_OUT_F32_LN) max_pairs = (max_pairs / p.num_n_tiles) * p.num_n_tiles;
    const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
    static const char* kname = EPI == WG_OUT_BF16 ? "gemm2_bf16" : (EPI == WG_OUT_F32 || EPI == WG_OUT_F32_TMA) ? "gemm2_f32" : EPI == WG_OUT_F32_LN ? "gemm2_f32_ln" : "gemm2_bf16ln";
    const double out_bytes = (double)a->M * a->N * (F32_EPI ? (a->resid ? 8.0 : 4.0) + (EPI == WG_OUT_F32_LN ? 2.0 : 0.0) : (a->resid ? 4.0 : 2.0));
    // ALGORITHMIC flops: a split-bf16 operand ([hi | lo | hi] against [W_hi | W_hi | W_lo], K = 2C or 3C executed) stands for ONE
    // fp32-accurate product over C = a_k_wrap / 2 columns
    const double k_alg = a->a_k_wrap > 0 ? 0.5 * a->a_k_wrap : (double)a->K;
    Prof prof(kname, stream, 2.0 * a->M * a->N * k_alg, 2.0 * ((double)a->M * a->K + (double)a->N * a->K) + out_bytes);
    kern<<<2 * pairs, NUM_THREADS, L::DYN_BYTES, stream>>>(tmA, tmB, tmC, tmR, p);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

}  // namespace

// fp32 output + fp32 residual (in place) + bf16 LayerNorm of the result over the full row, N % 256 == 0 with 2..8 n-tiles.  The n-tiles
// of a row block wait for each other inside the kernel, which is safe because the grid is persistent (every pair resident) and pairs
// take tiles in increasing order with at least as many pairs as n-tiles (no cyclic wait; see ln_pass2).  Used at every M (also below
// one wave of tiles) so that a row's arithmetic does not depend on the batch it is computed in.
int launch_gemm_pair_ln(const wg_gemm_args* a, const GemmLnFuse* f, cudaStream_t stream) {
    const int nt = a->N / BN2;
    if (a->out_mode != WG_OUT_F32 || a->N % BN2 != 0 || nt < 2 || nt > 8 || device_sm_count() / 2 < 2 * nt ||
        !f || !f->ln_out || !f->stats || !f->flags || f->epoch == 0 || !f->gamma || !f->beta || a->conv_grid != 0) {
        set_error("launch_gemm_pair_ln: unsupported problem (M=%d N=%d)", a->M, a->N);
        return WG_ERR_INVALID;
    }
    return launch2<6, WG_OUT_F32_LN>(a, stream, f);
}

// Used by wg_gemm for N % 256 == 0 problems with at least one full wave of 256 x 256 tiles.
int launch_gemm_pair(const wg_gemm_args* a, cudaStream_t stream) {
    switch (a->out_mode) {
        case WG_OUT_BF16: return launch2<6, WG_OUT_BF16>(a, stream);
        case WG_OUT_F32: {
            // TMA epilogue (5 stages + 64 KB of output boxes) unless WG_GEMM_F32_TMA=0 or the row pitch is not a multiple of 16 bytes
            static const bool tma_epi = [] { const char* e = getenv("WG_GEMM_F32_TMA"); return e == nullptr || atoi(e) != 0; }();
            // short K only: there the epilogue is the critical path (K = 1024 out-proj 0.175 -> 0.143 ms); the K = 4096 fc2 GEMM hides its
            // epilogue behind the main loop and prefers the sixth operand stage (0.386 vs 0.408 ms)
            const bool ok = tma_epi && a->K <= 2048 && a->ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(a->resid) & 15) == 0;
            return ok ? launch2<5, WG_OUT_F32_TMA>(a, stream) : launch2<6, WG_OUT_F32>(a, stream);
        }
        default: set_error("launch_gemm_pair: unsupported out_mode %d", a->out_mode); return WG_ERR_INVALID;
    }
}

}  // namespace wg

#ifdef GEMM2_TRACE
extern "C" __attribute__((visibility("default"))) int wg_debug_gemm2_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, wg::g_gemm2_trace, sizeof(long long) * 3 * 64 * 4);
}
extern "C" __attribute__((visibility("default"))) int wg_debug_gemm2_epi(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, wg::g_gemm2_epi, sizeof(long long) * 16 * 8);
}
#endif
