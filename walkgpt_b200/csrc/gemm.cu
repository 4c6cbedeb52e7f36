// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//   * operands staged by TMA (128B swizzle) through a multi-stage mbarrier ring
//   * tcgen05.mma (UMMA 128 x BN x 16, cta_group::1) issued by one thread, fp32 accumulators in TMEM,
//     double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1
//   * epilogue warps read TMEM with tcgen05.ld, fuse bias / activation / residual / LayerNorm and
//     write bf16 through swizzled smem + TMA store, or fp32 straight from registers.
// This one kernel serves the ViT QKV/out/MLP GEMMs (SURVEY K2,K4,K5), patch embedding (K1), MSQP
// projections (K6,K9,K10), the out_mm_projector MLP (K11), the neck convolutions (K12), CTP (K13) and
// the mask decoder's image-side projections and ConvTranspose (K14,K15).
#include "gemm_common.cuh"
#include <cstdlib>

namespace wg {

namespace {

using namespace gemm_detail;

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
    using L = SmemLayout<BN, STAGES, EPI>;
    static_assert(WG_OUT_F32_TMA == 17 && F32_NBOX == 2 && F32_WARP_BUF_BYTES == 4096, "SmemLayout's WG_OUT_F32_TMA sizes");
    constexpr int B_STAGE_BYTES = L::B_STAGE_BYTES;
    constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages (power of two: 256 or 512)
    constexpr uint32_t IDESC = umma_idesc_bf16(BM, BN, false, false);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint64_t* resid_bar = bars + 2 * STAGES + 4;  // WG_OUT_F32_TMA: [8 warps][F32_NBOX] residual box has landed
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (EPI != WG_OUT_F32) tma_prefetch_desc(&tmC);
        if (EPI == WG_OUT_F32_TMA) tma_prefetch_desc(&tmR);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);  // one arrival per epilogue warp
        }
        if (EPI == WG_OUT_F32_TMA) {
            for (int i = 0; i < 8 * F32_NBOX; ++i) mbar_init(&resid_bar[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int m0 = (tile / p.num_n_tiles) * BM;
            const int n0 = (tile % p.num_n_tiles) * BN;
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
                int ka = kb * BK;
                if (p.a_k_wrap > 0 && ka >= p.a_k_wrap) ka -= p.a_k_wrap;  // K' <= 2 * wrap by construction
                if (p.conv_g > 0) {
                    // implicit 3x3 convolution: k-block = (term, tap, 64-channel slice); the 128 output positions of this tile are
                    // 128 / g full rows of one image, their (kx-1, ky-1)-shifted neighbours one 4-D box (zero-filled outside the map)
                    const int per_term = 9 * p.conv_c;
                    const int term = ka / per_term, r = ka - term * per_term;
                    const int tap = r / p.conv_c, c = r - tap * p.conv_c;
                    const int gg = p.conv_g * p.conv_g;
                    const int img = m0 / gg, y0 = (m0 - img * gg) / p.conv_g;
                    tma_load_4d(smem + L::OFF_A + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], term * p.conv_c + c, tap % 3 - 1, y0 + tap / 3 - 1, img);
                } else
                tma_load_2d(smem + L::OFF_A + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], ka, m0);
                tma_load_2d(smem + L::OFF_B + stage * B_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, n0);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================== MMA issuer =====================
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            mbar_wait_relaxed(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                mbar_wait_relaxed(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + L::OFF_A + stage * A_STAGE_BYTES);
                const uint32_t b_addr = smem_u32(smem + L::OFF_B + stage * B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    umma_f16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), IDESC,
                                (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 2 groups x 4 warps (a warp may only touch TMEM lanes 32*(warp%4)..+31) ===============
        // The groups take alternate column chunks of every tile, so two warps per scheduler overlap each other's
        // TMEM-load / SFU / shared-store latencies: with K = 1024 a 128x256 tile leaves only ~8k cycles for its epilogue.
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int epi_tid = threadIdx.x - (4 + 4 * grp) * 32;  // index inside the group
        uint8_t* cbufs = smem + L::OFF_C;
        int iter = 0;
        // WG_OUT_F32_TMA (epilogue_tile_f32_tma): this warp's boxes, their barriers, its chunk counter, the first residual request
        uint8_t* wbuf = cbufs + (grp * 4 + q) * (F32_NBOX * F32_WARP_BUF_BYTES);
        uint64_t* rbar = resid_bar + (grp * 4 + q) * F32_NBOX;
        uint32_t cc = 0;
        if (EPI == WG_OUT_F32_TMA && p.resid_f32 != nullptr && (int)blockIdx.x < p.num_tiles && lane == 0) {
            const F32TmaNext none = {0, 0, false};
            for (int i = 0; i < F32_NBOX - 1; ++i)
                f32_tma_request<BN>(tmR, wbuf, rbar, i, i, ((int)blockIdx.x / p.num_n_tiles) * BM, ((int)blockIdx.x % p.num_n_tiles) * BN, q, grp, none);
        }
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int m0 = (tile / p.num_n_tiles) * BM;
            const int n0 = (tile % p.num_n_tiles) * BN;
            mbar_wait_relaxed(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

            if constexpr (EPI == WG_OUT_F32_TMA) {
                const int nt_ = tile + (int)gridDim.x;
                const F32TmaNext nx = {(nt_ / p.num_n_tiles) * BM, (nt_ % p.num_n_tiles) * BN, nt_ < p.num_tiles};
                epilogue_tile_f32_tma<BN>(p, tmC, tmR, wbuf, rbar, cc, taddr, m0, n0, q, grp, lane, nx);
            } else {
                epilogue_tile<BN, EPI>(p, tmC, cbufs, taddr, m0, n0, q, grp, lane, epi_tid);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (EPI != WG_OUT_F32 && EPI != WG_OUT_F32_TMA && epi_tid == 0) tma_store_wait_all<0>();
        if (EPI == WG_OUT_F32_TMA && lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

template <int BN, int STAGES, int EPI>
int launch_gemm(const wg_gemm_args* a, cudaStream_t stream) {
    using L = SmemLayout<BN, STAGES, EPI>;
    constexpr bool F32_ANY = EPI == WG_OUT_F32 || EPI == WG_OUT_F32_TMA;
    CUtensorMap tmA, tmB, tmC, tmR;
    if (a->conv_grid > 0) {
        const uint64_t g = (uint64_t)a->conv_grid, terms = a->a_k_wrap > 0 ? 2 : (uint64_t)a->K / (9 * (uint64_t)a->conv_channels);
        uint64_t dims[4] = {terms * (uint64_t)a->conv_channels, g, g, (uint64_t)a->M / (g * g)};
        uint64_t strides[3] = {(uint64_t)a->lda * 2, g * (uint64_t)a->lda * 2, g * g * (uint64_t)a->lda * 2};
        uint32_t box[4] = {(uint32_t)BK, (uint32_t)g, (uint32_t)(BM / a->conv_grid), 1};
        WG_TRY(make_tensor_map(&tmA, a->A, 2, 4, dims, strides, box));
    } else {
        WG_TRY(make_tmap_2d_bf16(&tmA, a->A, a->M, a->a_k_wrap > 0 ? a->a_k_wrap : a->K, a->lda, BM, BK));
    }
    WG_TRY(make_tmap_2d_bf16(&tmB, a->W, a->N, a->K, a->ldw, BN, BK));
    if (!F32_ANY) {
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
    p.num_n_tiles = (a->N + BN - 1) / BN;
    p.num_tiles = ((a->M + BM - 1) / BM) * p.num_n_tiles;
    p.num_k_blocks = (a->K + BK - 1) / BK;
    p.bias = a->bias;
    p.bias_period = a->bias_period;
    p.act = a->act;
    p.out_f32 = F32_ANY ? static_cast<float*>(a->out) : nullptr;
    p.resid_f32 = F32_ANY ? static_cast<const float*>(a->resid) : nullptr;
    p.resid_bf16 = (EPI == WG_OUT_BF16_LN) ? static_cast<const __nv_bfloat16*>(a->resid) : nullptr;
    p.ldo = a->ldo;
    p.ln_gamma = a->ln_gamma;
    p.ln_beta = a->ln_beta;
    p.ln_eps = a->ln_eps;
    p.a_k_wrap = a->a_k_wrap;
    p.conv_g = a->conv_grid;
    p.conv_c = a->conv_channels;
    p.split_out = !F32_ANY ? a->split_out : 0;
    p.ln_out = nullptr; p.ln_stats = nullptr; p.ln_flags = nullptr; p.ln_epoch = 0;

    auto kern = gemm_bf16_kernel<BN, STAGES, EPI>;
    WG_SMEM_OPT_IN(kern, L::DYN_BYTES);  // per instantiation and device
    int sms = device_sm_count();
    int grid = p.num_tiles < sms ? p.num_tiles : sms;
    static const char* kname = EPI == WG_OUT_BF16 ? (BN == 256 ? "gemm_bf16_bn256" : "gemm_bf16_bn128")
                               : F32_ANY ? (BN == 256 ? "gemm_f32_bn256" : "gemm_f32_bn128") : "gemm_bf16ln_bn256";
    const double out_bytes = (double)a->M * a->N * (F32_ANY ? (a->resid ? 8.0 : 4.0) : (a->resid ? 4.0 : 2.0));
    // ALGORITHMIC flops: a split-bf16 operand ([hi | lo | hi] against [W_hi | W_hi | W_lo], K = 2C or 3C executed) stands for ONE
    // fp32-accurate product over C = a_k_wrap / 2 columns
    const double k_alg = a->a_k_wrap > 0 ? 0.5 * a->a_k_wrap : (double)a->K;
    Prof prof(kname, stream, 2.0 * a->M * a->N * k_alg, 2.0 * ((double)a->M * a->K + (double)a->N * a->K) + out_bytes);
    kern<<<grid, NUM_THREADS, L::DYN_BYTES, stream>>>(tmA, tmB, tmC, tmR, p);
    WG_CHECK_CUDA(cudaGetLastError());
    return WG_OK;
}

}  // namespace
}  // namespace wg

extern "C" int wg_gemm(const wg_gemm_args* a, void* stream_) {
    using namespace wg;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WG_REQUIRE(a != nullptr, "wg_gemm: null args");
    WG_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "wg_gemm: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
    WG_REQUIRE(a->K % 8 == 0 && a->lda % 8 == 0 && a->ldw % 8 == 0, "wg_gemm: K/lda/ldw must be multiples of 8 (K=%d lda=%lld ldw=%lld)",
               a->K, (long long)a->lda, (long long)a->ldw);
    WG_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
               "wg_gemm: A/W/out must be 16-byte aligned");
    WG_REQUIRE(a->N % 8 == 0, "wg_gemm: N must be a multiple of 8 (N=%d)", a->N);
    WG_REQUIRE(a->ldo % 4 == 0, "wg_gemm: ldo must be a multiple of 4");
    WG_REQUIRE(a->a_k_wrap == 0 || (a->a_k_wrap % 64 == 0 && a->K <= 2 * a->a_k_wrap && a->K > a->a_k_wrap),
               "wg_gemm: a_k_wrap=%d must be a multiple of 64 with a_k_wrap < K <= 2*a_k_wrap (K=%d)", a->a_k_wrap, a->K);
    WG_REQUIRE(!a->split_out || (a->out_mode != WG_OUT_F32 && a->N % 64 == 0 && a->ldo >= 2 * (long long)a->N),
               "wg_gemm: split_out needs a bf16 output mode, N %% 64 == 0 and ldo >= 2N");
    if (a->conv_grid > 0) {
        const long long g = a->conv_grid, cc = a->conv_channels;
        WG_REQUIRE(128 % g == 0 && (g * g) % 128 == 0 && a->M % (g * g) == 0, "wg_gemm: conv_grid=%d needs 128 %% g == 0, g^2 %% 128 == 0 and M a multiple of g^2", a->conv_grid);
        WG_REQUIRE(cc > 0 && cc % 64 == 0 && a->K % (9 * cc) == 0, "wg_gemm: conv_channels=%d must be a multiple of 64 and divide K / 9", a->conv_channels);
        WG_REQUIRE(a->a_k_wrap == 0 || a->a_k_wrap == 18 * cc, "wg_gemm: with conv_grid, a_k_wrap must be 2 * 9 * conv_channels");
        WG_REQUIRE(a->lda >= (a->a_k_wrap > 0 ? 2 : a->K / (9 * cc)) * cc, "wg_gemm: conv map row pitch lda too small");
    }
    if (!device_is_sm100()) {
        set_error("wg_gemm: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const long long tiles256 = (long long)((a->M + 127) / 128) * ((a->N + 255) / 256);
    // CTA-pair kernel (256 x 256 tiles over two SMs) once there is at least one full wave of pair tiles; WG_GEMM_PAIR=0 disables
    static const int pair_enabled = [] {
        const char* e = getenv("WG_GEMM_PAIR");
        return (e == nullptr || atoi(e) != 0) ? 1 : 0;
    }();
    const long long tiles_pair = (long long)((a->M + 255) / 256) * ((a->N + 255) / 256);
    const bool use_pair = pair_enabled && a->conv_grid == 0 && (a->N % 256 == 0) && tiles_pair >= device_sm_count() / 2;
    switch (a->out_mode) {
        case WG_OUT_BF16: {
            WG_REQUIRE(a->ldo % 8 == 0, "wg_gemm: bf16 output needs ldo %% 8 == 0");
            if (use_pair) return launch_gemm_pair(a, stream);
            bool use256 = (a->N % 256 == 0) && tiles256 >= device_sm_count();
            return use256 ? launch_gemm<256, 4, WG_OUT_BF16>(a, stream) : launch_gemm<128, 6, WG_OUT_BF16>(a, stream);
        }
        case WG_OUT_F32: {
            if (use_pair) return launch_gemm_pair(a, stream);
            bool use256 = (a->N % 256 == 0) && tiles256 >= device_sm_count();
            if (use256) return launch_gemm<256, 4, WG_OUT_F32>(a, stream);
            // 128-wide tiles: TMA epilogue (epilogue_tile_f32_tma) for short K, where the epilogue is the critical path
            static const bool tma_epi = [] { const char* e = getenv("WG_GEMM_F32_TMA"); return e == nullptr || atoi(e) != 0; }();
            if (tma_epi && a->K <= 2048 && a->N % 32 == 0 && (reinterpret_cast<uintptr_t>(a->resid) & 15) == 0) return launch_gemm<128, 5, WG_OUT_F32_TMA>(a, stream);
            return launch_gemm<128, 6, WG_OUT_F32>(a, stream);
        }
        case WG_OUT_BF16_LN:
            WG_REQUIRE(a->N == 256, "wg_gemm: WG_OUT_BF16_LN requires N == 256 (got %d)", a->N);
            WG_REQUIRE(a->ln_gamma && a->ln_beta, "wg_gemm: WG_OUT_BF16_LN needs gamma/beta");
            WG_REQUIRE(a->ldo % 8 == 0, "wg_gemm: bf16 output needs ldo %% 8 == 0");
            return launch_gemm<256, 4, WG_OUT_BF16_LN>(a, stream);
        default:
            set_error("wg_gemm: unknown out_mode %d", a->out_mode);
            return WG_ERR_INVALID;
    }
}
