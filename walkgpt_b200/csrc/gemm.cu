// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//   * operands staged by TMA (128B swizzle) through a multi-stage mbarrier ring
//   * tcgen05.mma (UMMA 128 x BN x 16, cta_group::1) issued by one thread, fp32 accumulators in TMEM,
//     double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1
//   * epilogue warps read TMEM with tcgen05.ld, fuse bias / activation / residual / LayerNorm and
//     write bf16 through swizzled smem + TMA store, or fp32 straight from registers.
// This one kernel serves the ViT QKV/out/MLP GEMMs (SURVEY K2,K4,K5), patch embedding (K1), MSQP
// projections (K6,K9,K10), the out_mm_projector MLP (K11), the neck convolutions (K12), CTP (K13) and
// the mask decoder's image-side projections and ConvTranspose (K14,K15).
#include "host.h"
#include "ptx.cuh"

namespace wg {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int C_BUF_BYTES = BM * 64 * 2;  // one TMA-store box: 128 rows x 64 bf16
constexpr int NUM_THREADS = 384;   // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warp 3 idle, warps 4-11 epilogue
constexpr int EPI_GROUP_THREADS = 128;  // two epilogue groups of 4 warps (one warp per TMEM lane quarter) take alternate column chunks

struct GemmParams {
    int M, N, K;
    int num_n_tiles, num_tiles, num_k_blocks;
    const float* bias;
    int bias_period;
    int act;
    float* out_f32;
    const float* resid_f32;
    const __nv_bfloat16* resid_bf16;
    long long ldo;
    const float* ln_gamma;
    const float* ln_beta;
    float ln_eps;
    int a_k_wrap;   // A's k coordinate wraps at this many columns (0 = off): A' = [hi | lo | hi] of a split-bf16 operand
    int split_out;  // bf16 outputs are written as split-bf16: hi at column c, lo = bf16(x - hi) at column N + c
};

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = STAGES * A_STAGE_BYTES;
    static constexpr int OFF_C = OFF_B + STAGES * B_STAGE_BYTES;
    static constexpr int OFF_BAR = OFF_C + 2 * C_BUF_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 4;
    static constexpr int TOTAL = OFF_BAR + NUM_BARS * 8 + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024B alignment
};

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case WG_ACT_QUICK_GELU: return quick_gelu(x);
        case WG_ACT_GELU_ERF: return gelu_erf(x);
        case WG_ACT_RELU: return fmaxf(x, 0.0f);
        default: return x;
    }
}

// v[0..31] (raw accumulator bits) -> f[0..31] = acc + bias, for row `row`, columns col0..col0+31
__device__ __forceinline__ void add_bias32(const uint32_t (&v)[32], float (&f)[32], const GemmParams& p, int row, int col0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (p.bias != nullptr) {
        const float* b = p.bias + (p.bias_period > 1 ? (size_t)(row % p.bias_period) * p.N : 0);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            int col = col0 + g * 4;
            if (col < p.N) {
                float4 bb = __ldg(reinterpret_cast<const float4*>(b + col));
                f[g * 4 + 0] += bb.x;
                f[g * 4 + 1] += bb.y;
                f[g * 4 + 2] += bb.z;
                f[g * 4 + 3] += bb.w;
            }
        }
    }
}

__device__ __forceinline__ void add_resid_bf16_32(float (&f)[32], const __nv_bfloat16* r, bool row_ok, int col0, int N) {
    if (!row_ok) return;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        int col = col0 + g * 8;
        if (col < N) {
            uint4 u = __ldg(reinterpret_cast<const uint4*>(r + col));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 t = __bfloat1622float2(h[e]);
                f[g * 8 + e * 2] += t.x;
                f[g * 8 + e * 2 + 1] += t.y;
            }
        }
    }
}

// resid stored as split-bf16 [.., 2N]: value = hi + lo
__device__ __forceinline__ void add_resid_split_32(float (&f)[32], const __nv_bfloat16* r_hi, int N, bool row_ok, int ncols_left) {
    add_resid_bf16_32(f, r_hi, row_ok, 0, ncols_left);
    add_resid_bf16_32(f, r_hi + N, row_ok, 0, ncols_left);
}

// write 32 fp32 values as bf16 into a 128B-swizzled [128 x 64] staging tile (row r, column half `half`)
__device__ __forceinline__ void stage_bf16_32(uint8_t* cbuf, int r, int half, const float (&f)[32]) {
    uint8_t* rowp = cbuf + r * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16x2(f[j * 8 + 0], f[j * 8 + 1]);
        u.y = pack_bf16x2(f[j * 8 + 2], f[j * 8 + 3]);
        u.z = pack_bf16x2(f[j * 8 + 4], f[j * 8 + 5]);
        u.w = pack_bf16x2(f[j * 8 + 6], f[j * 8 + 7]);
        int chunk = (half * 4 + j) ^ (r & 7);
        *reinterpret_cast<uint4*>(rowp + chunk * 16) = u;
    }
}

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    using L = SmemLayout<BN, STAGES>;
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
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (EPI != WG_OUT_F32) tma_prefetch_desc(&tmC);
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
        const int r = q * 32 + lane;  // row inside the tile == TMEM lane
        const int epi_tid = threadIdx.x - (4 + 4 * grp) * 32;  // index inside the group
        uint8_t* cbufs = smem + L::OFF_C;
        int iter = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int m0 = (tile / p.num_n_tiles) * BM;
            const int n0 = (tile % p.num_n_tiles) * BN;
            const int row = m0 + r;
            const bool row_ok = row < p.M;
            mbar_wait_relaxed(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

            if constexpr (EPI == WG_OUT_F32) {
                // fp32 output (+ fp32 residual, in place on the ViT residual stream).  Each thread owns one accumulator row, which
                // would make global accesses 16 B per thread at a 4 KB stride; instead every warp transposes its 32x32 chunk
                // through a padded smem tile so that each quarter-warp touches one full 128-byte line (coalesced LDG/STG.128).
                // per-warp 32x32 fp32 tile, 16-byte chunks XOR-swizzled by row (conflict-free float4 writes by row and reads by line)
                float* stg = reinterpret_cast<float*>(cbufs) + (grp * 4 + q) * (32 * 32);
                const int rr0 = lane >> 3, cj = lane & 7, cc = cj * 4;
#pragma unroll 1
                for (int c = grp; c < BN / 32; c += 2) {
                    uint32_t v[32];
                    float f[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    const int col0 = n0 + c * 32;
                    if (col0 >= p.N) continue;
                    add_bias32(v, f, p, row, col0);
                    if (p.act != WG_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
                    }
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        *reinterpret_cast<float4*>(stg + lane * 32 + ((g ^ (lane & 7)) * 4)) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
                    __syncwarp();
                    const int gcol = col0 + cc;
                    float4 o[8], rsd[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int grow = m0 + q * 32 + it * 4 + rr0;
                        o[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + rr0) * 32 + ((cj ^ ((it * 4 + rr0) & 7)) * 4));
                        rsd[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.resid_f32 && grow < p.M && gcol < p.N)
                            rsd[it] = *reinterpret_cast<const float4*>(p.resid_f32 + (size_t)grow * p.ldo + gcol);
                    }
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int grow = m0 + q * 32 + it * 4 + rr0;
                        if (grow < p.M && gcol < p.N) {
                            o[it].x += rsd[it].x; o[it].y += rsd[it].y; o[it].z += rsd[it].z; o[it].w += rsd[it].w;
                            *reinterpret_cast<float4*>(p.out_f32 + (size_t)grow * p.ldo + gcol) = o[it];
                        }
                    }
                    __syncwarp();
                }
            } else {
                float mean = 0.f, rstd = 1.f;
                const __nv_bfloat16* rrow = p.resid_bf16 ? p.resid_bf16 + (size_t)row * p.ldo : nullptr;
                if constexpr (EPI == WG_OUT_BF16_LN) {
                    // pass 1: row statistics over the full N == BN columns (this thread owns the row)
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
                    for (int c = 0; c < BN / 32; ++c) {
                        uint32_t v[32];
                        float f[32];
                        tmem_ld_32x32b_x32(taddr + c * 32, v);
                        tmem_ld_wait();
                        add_bias32(v, f, p, row, n0 + c * 32);
                        if (rrow) {
                            if (p.split_out) add_resid_split_32(f, rrow + n0 + c * 32, p.N, row_ok, BN - c * 32);
                            else add_resid_bf16_32(f, rrow + n0 + c * 32, row_ok, 0, BN - c * 32);
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            s1 += f[j];
                            s2 = fmaf(f[j], f[j], s2);
                        }
                    }
                    mean = s1 * (1.0f / BN);
                    float var = fmaxf(s2 * (1.0f / BN) - mean * mean, 0.f);
                    rstd = rsqrtf(var + p.ln_eps);
                }
                const int nparts = p.split_out ? 2 : 1;
#pragma unroll 1
                for (int cp = grp; cp < (BN / 64) * nparts; cp += 2) {
                    const int c = cp / nparts, part = cp - c * nparts;
                    const int colc = n0 + c * 64;
                    if (colc >= p.N) break;  // uniform across the group
                    uint8_t* cbuf = cbufs + grp * C_BUF_BYTES;  // one staging buffer per group
                    if (epi_tid == 0) tma_store_wait_read<0>();  // this group's previous store is done reading the buffer
                    named_bar_sync(1 + grp, EPI_GROUP_THREADS);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t v[32];
                        float f[32];
                        tmem_ld_32x32b_x32(taddr + c * 64 + half * 32, v);
                        tmem_ld_wait();
                        const int col0 = colc + half * 32;
                        add_bias32(v, f, p, row, col0);
                        if constexpr (EPI == WG_OUT_BF16_LN) {
                            if (rrow) {
                                if (p.split_out) add_resid_split_32(f, rrow + col0, p.N, row_ok, p.N - col0);
                                else add_resid_bf16_32(f, rrow + col0, row_ok, 0, p.N - col0);
                            }
#pragma unroll
                            for (int g = 0; g < 8; ++g) {
                                float4 ga = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col0 + g * 4));
                                float4 be = __ldg(reinterpret_cast<const float4*>(p.ln_beta + col0 + g * 4));
                                f[g * 4 + 0] = (f[g * 4 + 0] - mean) * rstd * ga.x + be.x;
                                f[g * 4 + 1] = (f[g * 4 + 1] - mean) * rstd * ga.y + be.y;
                                f[g * 4 + 2] = (f[g * 4 + 2] - mean) * rstd * ga.z + be.z;
                                f[g * 4 + 3] = (f[g * 4 + 3] - mean) * rstd * ga.w + be.w;
                            }
                        } else {
                            if (p.act == WG_ACT_QUICK_GELU) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] = quick_gelu(f[j]);
                            } else if (p.act == WG_ACT_GELU_ERF) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
                            } else if (p.act == WG_ACT_RELU) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                            }
                        }
                        if (part == 1) {  // low half of the split representation
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] -= __bfloat162float(__float2bfloat16(f[j]));
                        }
                        stage_bf16_32(cbuf, r, half, f);
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1 + grp, EPI_GROUP_THREADS);
                    if (epi_tid == 0) {
                        tma_store_2d(&tmC, cbuf, part * p.N + colc, m0);
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (EPI != WG_OUT_F32 && epi_tid == 0) tma_store_wait_all<0>();
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
    using L = SmemLayout<BN, STAGES>;
    CUtensorMap tmA, tmB, tmC;
    WG_TRY(make_tmap_2d_bf16(&tmA, a->A, a->M, a->a_k_wrap > 0 ? a->a_k_wrap : a->K, a->lda, BM, BK));
    WG_TRY(make_tmap_2d_bf16(&tmB, a->W, a->N, a->K, a->ldw, BN, BK));
    if (EPI != WG_OUT_F32) {
        WG_TRY(make_tmap_2d_bf16(&tmC, a->out, a->M, a->split_out ? 2 * a->N : a->N, a->ldo, BM, 64));
    } else {
        tmC = tmA;
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
    p.out_f32 = (EPI == WG_OUT_F32) ? static_cast<float*>(a->out) : nullptr;
    p.resid_f32 = (EPI == WG_OUT_F32) ? static_cast<const float*>(a->resid) : nullptr;
    p.resid_bf16 = (EPI == WG_OUT_BF16_LN) ? static_cast<const __nv_bfloat16*>(a->resid) : nullptr;
    p.ldo = a->ldo;
    p.ln_gamma = a->ln_gamma;
    p.ln_beta = a->ln_beta;
    p.ln_eps = a->ln_eps;
    p.a_k_wrap = a->a_k_wrap;
    p.split_out = (EPI != WG_OUT_F32) ? a->split_out : 0;

    auto kern = gemm_bf16_kernel<BN, STAGES, EPI>;
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        WG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
        attr_set = true;
    }
    int sms = device_sm_count();
    int grid = p.num_tiles < sms ? p.num_tiles : sms;
    static const char* kname = EPI == WG_OUT_BF16 ? (BN == 256 ? "gemm_bf16_bn256" : "gemm_bf16_bn128")
                               : EPI == WG_OUT_F32 ? (BN == 256 ? "gemm_f32_bn256" : "gemm_f32_bn128") : "gemm_bf16ln_bn256";
    const double out_bytes = (double)a->M * a->N * (EPI == WG_OUT_F32 ? (a->resid ? 8.0 : 4.0) : (a->resid ? 4.0 : 2.0));
    Prof prof(kname, stream, 2.0 * a->M * a->N * a->K, 2.0 * ((double)a->M * a->K + (double)a->N * a->K) + out_bytes);
    kern<<<grid, NUM_THREADS, L::DYN_BYTES, stream>>>(tmA, tmB, tmC, p);
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
    if (!device_is_sm100()) {
        set_error("wg_gemm: this library only runs on sm_100 (B200) devices; there is no fallback");
        return WG_ERR_UNSUPPORTED;
    }
    const long long tiles256 = (long long)((a->M + 127) / 128) * ((a->N + 255) / 256);
    switch (a->out_mode) {
        case WG_OUT_BF16: {
            WG_REQUIRE(a->ldo % 8 == 0, "wg_gemm: bf16 output needs ldo %% 8 == 0");
            bool use256 = (a->N % 256 == 0) && tiles256 >= device_sm_count();
            return use256 ? launch_gemm<256, 4, WG_OUT_BF16>(a, stream) : launch_gemm<128, 6, WG_OUT_BF16>(a, stream);
        }
        case WG_OUT_F32: {
            bool use256 = (a->N % 256 == 0) && tiles256 >= device_sm_count();
            return use256 ? launch_gemm<256, 4, WG_OUT_F32>(a, stream) : launch_gemm<128, 6, WG_OUT_F32>(a, stream);
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
