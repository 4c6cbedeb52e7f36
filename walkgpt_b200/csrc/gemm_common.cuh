// Shared pieces of the tcgen05 GEMM kernels (1-CTA `gemm.cu`, 2-CTA pair `gemm2.cu`): parameters, epilogue math and the
// per-tile epilogue (TMEM -> registers -> fused bias / activation / residual / LayerNorm -> smem -> TMA store, or fp32 global).
#pragma once
#include "host.h"
#include "ptx.cuh"

namespace wg {

// gemm2.cu: CTA-pair (cta_group::2) kernel, 256 x 256 tiles; out_mode WG_OUT_BF16 or WG_OUT_F32, N % 256 == 0
int launch_gemm_pair(const wg_gemm_args* a, cudaStream_t stream);

namespace gemm_detail {

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
    int conv_g, conv_c;  // implicit 3x3 convolution over a channels-last g x g map with conv_c channels per term (conv_g = 0: off)
    int a_k_wrap;   // A's k coordinate wraps at this many columns (0 = off): A' = [hi | lo | hi] of a split-bf16 operand
    int split_out;  // bf16 outputs are written as split-bf16: hi at column c, lo = bf16(x - hi) at column N + c
    // WG_OUT_F32_LN (pair kernel only): the fp32 output row block is ALSO emitted as bf16 LayerNorm_N(out) * gamma + beta.  A row
    // spans N / 256 tiles that different CTA pairs compute at the same time: every CTA publishes its rows' (sum, sum of squares) over
    // its 256 columns in ln_stats, raises ln_flags[row block][n tile] to ln_epoch, and waits for the other tiles of its row block.
    __nv_bfloat16* ln_out;   // bf16 [M, N], ld = N
    float* ln_stats;         // [2 * ceil(M / 256)][N / 256][128][2]
    unsigned* ln_flags;      // [2 * ceil(M / 256)][N / 256], zeroed before the first launch of a sequence
    unsigned ln_epoch;       // unique (> 0) per launch since the flags were zeroed
};

constexpr int WG_OUT_F32_TMA = 17;  // internal out mode of the pair kernel: WG_OUT_F32 whose residual tiles arrive and whose output tiles leave as TMA
                                    // bulk transfers through per-warp shared-memory buffers (epilogue_tile_f32_tma in gemm2.cu)
constexpr int WG_OUT_F32_LN = 16;  // internal out mode: WG_OUT_F32 + residual + fused LayerNorm over the full row (launch_gemm_pair_ln)

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int BN, int STAGES, int EPI = 0>
struct SmemLayout {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = STAGES * A_STAGE_BYTES;
    static constexpr int OFF_C = OFF_B + STAGES * B_STAGE_BYTES;
    // WG_OUT_F32_TMA (= 17): F32_NBOX (= 2) boxes of 4 KB per epilogue warp + one mbarrier per box
    static constexpr int OFF_BAR = OFF_C + (EPI == 17 ? 8 * 2 * 4096 : 2 * C_BUF_BYTES);
    static constexpr int NUM_BARS = 2 * STAGES + 4 + (EPI == 17 ? 16 : 0);
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

// Epilogue of one 128 x BN accumulator tile held in this CTA's TMEM at `taddr` (lane quarter q already applied).
// Called by all 8 epilogue warps (grp = 0/1 take alternate column chunks); m0/n0 = global origin of this CTA's tile.
// WG_OUT_F32_LN: per-thread row statistics handed from the tile's first pass (epilogue_tile) to ln_pass2 -- a thread of the transposed
// store layout touches rows q*32 + it*4 + (lane >> 3), it = 0..7, of its warp's 32 rows
struct LnRowStats {
    float s1[8], s2[8];
};

template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const CUtensorMap& tmC, uint8_t* cbufs, uint32_t taddr, int m0, int n0, int q,
                                              int grp, int lane, int epi_tid, LnRowStats* ln = nullptr) {
    const int r = q * 32 + lane;
    const int row = m0 + r;
    const bool row_ok = row < p.M;
    if constexpr (EPI == WG_OUT_F32 || EPI == WG_OUT_F32_LN) {
        constexpr bool FUSE_LN = EPI == WG_OUT_F32_LN;
        if constexpr (FUSE_LN) {
#pragma unroll
            for (int it = 0; it < 8; ++it) ln->s1[it] = ln->s2[it] = 0.f;
        }
        // fp32 output (+ fp32 residual, in place on the ViT residual stream).  Each thread owns one accumulator row, which
        // would make global accesses 16 B per thread at a 4 KB stride; instead every warp transposes its 32x32 chunk
        // through a padded smem tile so that each quarter-warp touches one full 128-byte line (coalesced LDG/STG.128).
        // per-warp 32x32 fp32 tile, 16-byte chunks XOR-swizzled by row (conflict-free float4 writes by row and reads by line)
        float* stg = reinterpret_cast<float*>(cbufs) + (grp * 4 + q) * (32 * 32);
        const int rr0 = lane >> 3, cj = lane & 7, cc = cj * 4;
        // the residual rows of a chunk do not depend on the accumulator: they are requested one chunk ahead (the first chunk's
        // before the TMEM load), so that their HBM latency overlaps the TMEM load, the transpose and the previous chunk's stores
        // instead of being paid once per chunk (the epilogue of the K = 1024 out-proj GEMM is longer than its main loop)
        float4 rsd[8];
        auto load_resid = [&](int c, float4 (&dst)[8]) {
            const int gcol = n0 + c * 32 + cc;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int grow = m0 + q * 32 + it * 4 + rr0;
                dst[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (grow < p.M && gcol < p.N) dst[it] = *reinterpret_cast<const float4*>(p.resid_f32 + (size_t)grow * p.ldo + gcol);
            }
        };
        if (p.resid_f32) load_resid(grp, rsd);
#pragma unroll 1
        for (int c = grp; c < BN / 32; c += 2) {
            uint32_t v[32];
            float f[32];
            tmem_ld_32x32b_x32(taddr + c * 32, v);
            tmem_ld_wait();
            const int col0 = n0 + c * 32;
            if (col0 >= p.N) continue;
#ifdef GEMM_DBG_NOEPI  // debug build (wrong results): the fp32 epilogue only drains tensor memory -- what the main loop alone costs
            if (v[0] != 0x7fc12345u) continue;
#endif
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
            float4 o[8], nxt[8];
#pragma unroll
            for (int it = 0; it < 8; ++it)
                o[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + rr0) * 32 + ((cj ^ ((it * 4 + rr0) & 7)) * 4));
            if (p.resid_f32 && c + 2 < BN / 32) load_resid(c + 2, nxt);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int grow = m0 + q * 32 + it * 4 + rr0;
                if (grow < p.M && gcol < p.N) {
                    if (p.resid_f32) { o[it].x += rsd[it].x; o[it].y += rsd[it].y; o[it].z += rsd[it].z; o[it].w += rsd[it].w; }
                    *reinterpret_cast<float4*>(p.out_f32 + (size_t)grow * p.ldo + gcol) = o[it];
                    if constexpr (FUSE_LN) {
                        ln->s1[it] += (o[it].x + o[it].y) + (o[it].z + o[it].w);
                        ln->s2[it] = fmaf(o[it].x, o[it].x, fmaf(o[it].y, o[it].y, fmaf(o[it].z, o[it].z, fmaf(o[it].w, o[it].w, ln->s2[it]))));
                    }
                }
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) rsd[it] = nxt[it];
            __syncwarp();
        }
        if constexpr (FUSE_LN) {
            // (sum, sum of squares) of this CTA's 256 columns per row: reduce over the 8 lanes that share a row, add the other column
            // group's half through the (now idle) staging tiles, publish for the other n-tiles of the row block
#pragma unroll
            for (int it = 0; it < 8; ++it) {
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
                    ln->s1[it] += __shfl_xor_sync(0xffffffffu, ln->s1[it], d);
                    ln->s2[it] += __shfl_xor_sync(0xffffffffu, ln->s2[it], d);
                }
            }
            float2* mine = reinterpret_cast<float2*>(stg);                                                               // [32 rows]
            const float2* other = reinterpret_cast<const float2*>(reinterpret_cast<float*>(cbufs) + ((grp ^ 1) * 4 + q) * (32 * 32));
            if (cj == 0) {
#pragma unroll
                for (int it = 0; it < 8; ++it) mine[it * 4 + rr0] = make_float2(ln->s1[it], ln->s2[it]);
            }
            named_bar_sync(3, 2 * EPI_GROUP_THREADS);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const float2 o2 = other[it * 4 + rr0];
                ln->s1[it] += o2.x;
                ln->s2[it] += o2.y;
            }
            const int nt = p.num_n_tiles;
            const size_t slot = (size_t)(m0 >> 7) * nt + (n0 / BN);
            if (grp == 0 && cj == 0) {
                float2* dst = reinterpret_cast<float2*>(p.ln_stats) + slot * 128 + q * 32;
#pragma unroll
                for (int it = 0; it < 8; ++it) dst[it * 4 + rr0] = make_float2(ln->s1[it], ln->s2[it]);
            }
            named_bar_sync(3, 2 * EPI_GROUP_THREADS);  // every staging read and every statistics store of this CTA is done
            if (epi_tid == 0 && grp == 0) {
                __threadfence();
                st_release_gpu(p.ln_flags + slot, p.ln_epoch);
            }
        }
    } else {
        float mean = 0.f, rstd = 1.f;
        const __nv_bfloat16* rrow = p.resid_bf16 ? p.resid_bf16 + (size_t)row * p.ldo : nullptr;
        if constexpr (EPI == WG_OUT_BF16_LN) {
            // pass 1: acc + bias + residual for the full N == BN columns of this thread's row, split between the two epilogue
            // groups (alternate 32-column chunks).  The sums are written BACK to tensor memory, so pass 2 neither re-reads the
            // residual rows (the expensive, uncoalesced part of this epilogue) nor the bias; the two groups exchange their
            // partial (sum, sum of squares) through the head of the staging area.
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c = grp; c < BN / 32; c += 2) {
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
                    v[j] = __float_as_uint(f[j]);
                }
                tmem_st_32x32b_x32(taddr + c * 32, v);
            }
            tmem_st_wait();
            tc_fence_before();
            float2* xch = reinterpret_cast<float2*>(cbufs);  // [2 groups][128 rows]
            if (grp == 0 && epi_tid == 0) tma_store_wait_read<0>();  // the previous tile's store no longer reads the staging area
            named_bar_sync(3, 2 * EPI_GROUP_THREADS);
            xch[grp * 128 + r] = make_float2(s1, s2);
            named_bar_sync(3, 2 * EPI_GROUP_THREADS);
            const float2 other = xch[(grp ^ 1) * 128 + r];
            named_bar_sync(3, 2 * EPI_GROUP_THREADS);  // both groups have read before pass 2 reuses the staging area
            tc_fence_after();
            s1 += other.x;
            s2 += other.y;
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
                if constexpr (EPI == WG_OUT_BF16_LN) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);  // pass 1 left acc + bias + residual in TMEM
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
                    add_bias32(v, f, p, row, col0);
                    if (p.act == WG_ACT_QUICK_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = quick_gelu(f[j]);
                    } else if (p.act == WG_ACT_GELU_ERF) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = gelu_erf_fast(f[j]);
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
}

// WG_OUT_F32_LN, second pass of a tile (after its accumulator has been handed back to the MMA warp): wait for the other n-tiles of
// the row block, complete the row statistics, re-read this CTA's fp32 output tile (just written: L2) in the coalesced store layout,
// normalise, write bf16.  All 8 epilogue warps call it.
template <int BN>
__device__ __forceinline__ void ln_pass2(const GemmParams& p, int m0, int n0, int q, int grp, int lane, int epi_tid, LnRowStats& ln) {
    const int rr0 = lane >> 3, cj = lane & 7, cc = cj * 4;
    const int nt = p.num_n_tiles;
    const int my_n = n0 / BN;
    const size_t slot0 = (size_t)(m0 >> 7) * nt;
    if (epi_tid == 0 && grp == 0) {
        for (int t = 0; t < nt; ++t) {
            if (t == my_n) continue;
            while (ld_acquire_gpu(p.ln_flags + slot0 + t) != p.ln_epoch) __nanosleep(64);
        }
    }
    named_bar_sync(3, 2 * EPI_GROUP_THREADS);
    float a[8], b[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        float s1 = 0.f, s2 = 0.f;  // summed in n-tile order by every tile of the row block: all of them normalise with the same bits
        for (int t = 0; t < nt; ++t) {
            float2 v = make_float2(ln.s1[it], ln.s2[it]);
            if (t != my_n) v = __ldcg(reinterpret_cast<const float2*>(p.ln_stats) + (slot0 + t) * 128 + q * 32 + it * 4 + rr0);
            s1 += v.x;
            s2 += v.y;
        }
        const float inv_n = 1.0f / (float)p.N;
        const float mean = s1 * inv_n;
        const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
        a[it] = rsqrtf(var + p.ln_eps);
        b[it] = -mean * a[it];
    }
#pragma unroll 1
    for (int c = grp; c < BN / 32; c += 2) {
        const int gcol = n0 + c * 32 + cc;
        const float4 ga = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + gcol));
        const float4 be = __ldg(reinterpret_cast<const float4*>(p.ln_beta + gcol));
        float4 x[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int grow = m0 + q * 32 + it * 4 + rr0;
            x[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow < p.M) x[it] = __ldcg(reinterpret_cast<const float4*>(p.out_f32 + (size_t)grow * p.ldo + gcol));
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int grow = m0 + q * 32 + it * 4 + rr0;
            if (grow < p.M) {
                uint2 u;
                u.x = pack_bf16x2(fmaf(fmaf(x[it].x, a[it], b[it]), ga.x, be.x), fmaf(fmaf(x[it].y, a[it], b[it]), ga.y, be.y));
                u.y = pack_bf16x2(fmaf(fmaf(x[it].z, a[it], b[it]), ga.z, be.z), fmaf(fmaf(x[it].w, a[it], b[it]), ga.w, be.w));
                *reinterpret_cast<uint2*>(p.ln_out + (size_t)grow * p.N + gcol) = u;
            }
        }
    }
}

constexpr int F32_WARP_BUF_BYTES = 32 * 32 * 4;  // WG_OUT_F32_TMA: one [32 rows x 32 columns] fp32 box, 128-byte rows under the 128B swizzle
constexpr int F32_NBOX = 2;                      // boxes per epilogue warp: residual boxes are requested F32_NBOX - 1 chunks ahead (3 boxes with 4
                                                 // operand stages measured the same 0.143 ms on the out-proj GEMM: the depth is not the limit)


// fp32 epilogue (+ fp32 residual, possibly in place) without per-thread global accesses.  A warp owns 32 rows x 32-column chunks of its
// CTA's accumulator tile (the two epilogue groups take alternate chunks): the chunk's residual box is fetched by TMA into one of the
// warp's two shared-memory boxes one chunk AHEAD (across tile boundaries too), the row-owning thread adds its accumulator row (+ bias,
// activation) in place with conflict-free 16-byte accesses (TMA's 128-byte swizzle = the XOR pattern the row owners use), and the box
// leaves as a TMA store; rows past M are zero-filled on the way in and clipped on the way out.  The round-1 epilogue transposed every
// chunk through shared memory to reach coalesced LDG / STG.128 and was the critical path of the K = 1024 out-proj GEMM (16 000 cycles per
// tile against a 10 000-cycle main loop).  `cc` = chunks this warp has processed (box = cc & 1, barrier phase = (cc >> 1) & 1).
struct F32TmaNext {
    int m0, n0;   // the CTA's next tile (prefetch of its first residual boxes), valid if ok
    bool ok;
};
// box / barrier of the warp's chunk number x (running count over all tiles)
__device__ __forceinline__ uint32_t f32_box(uint32_t x) { return x % F32_NBOX; }
// request the residual box of the warp's chunk `i` of tile (m0, n0) -- or, past the tile's last chunk, of the next tile -- as chunk number x
template <int BN>
__device__ __forceinline__ void f32_tma_request(const CUtensorMap& tmR, uint8_t* wbuf, uint64_t* rbar, uint32_t x, int i, int m0, int n0, int q, int grp,
                                                const F32TmaNext& nx) {
    constexpr int NCH = BN / 64;
    const uint32_t b = f32_box(x);
    int col, rowc;
    if (i < NCH) {
        col = n0 + (grp + 2 * i) * 32;
        rowc = m0 + q * 32;
    } else {
        if (!nx.ok) return;
        col = nx.n0 + (grp + 2 * (i - NCH)) * 32;
        rowc = nx.m0 + q * 32;
    }
    mbar_arrive_expect_tx(&rbar[b], F32_WARP_BUF_BYTES);
    tma_load_2d(wbuf + b * F32_WARP_BUF_BYTES, &tmR, &rbar[b], col, rowc);
}
template <int BN>
__device__ __forceinline__ void epilogue_tile_f32_tma(const GemmParams& p, const CUtensorMap& tmC, const CUtensorMap& tmR, uint8_t* wbuf, uint64_t* rbar,
                                                      uint32_t& cc, uint32_t taddr, int m0, int n0, int q, int grp, int lane, const F32TmaNext& nx) {
    const int row = m0 + q * 32 + lane;
    const int wrow0 = m0 + q * 32;
    const bool has_res = p.resid_f32 != nullptr;
    constexpr int NCH = BN / 64;  // chunks per group
    static_assert(F32_NBOX - 1 <= NCH, "prefetch distance within one tile ahead");
#pragma unroll 1
    for (int i = 0; i < NCH; ++i) {
        const int c = grp + 2 * i;
        const uint32_t b = f32_box(cc);
        uint32_t v[32];
        float f[32];
        tmem_ld_32x32b_x32(taddr + c * 32, v);
        tmem_ld_wait();
        add_bias32(v, f, p, row, n0 + c * 32);
        if (p.act != WG_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
        }
        uint8_t* box = wbuf + b * F32_WARP_BUF_BYTES;
        if (has_res) {
            // the residual box F32_NBOX - 1 chunks ahead goes into the box whose store (chunk cc - 1) was issued last: it has had the
            // whole TMEM load + bias phase to finish reading
            if (lane == 0) {
                tma_store_wait_read<0>();
                f32_tma_request<BN>(tmR, wbuf, rbar, cc + F32_NBOX - 1, i + F32_NBOX - 1, m0, n0, q, grp, nx);
            }
            mbar_wait(&rbar[b], (cc / F32_NBOX) & 1);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float4* slot = reinterpret_cast<float4*>(box + lane * 128 + ((g ^ (lane & 7)) << 4));
                float4 r4 = *slot;
                r4.x += f[g * 4]; r4.y += f[g * 4 + 1]; r4.z += f[g * 4 + 2]; r4.w += f[g * 4 + 3];
                *slot = r4;
            }
        } else {
            if (lane == 0) tma_store_wait_read<F32_NBOX - 1>();  // the store of chunk cc - F32_NBOX (this box) has read it
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g)
                *reinterpret_cast<float4*>(box + lane * 128 + ((g ^ (lane & 7)) << 4)) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(&tmC, box, n0 + c * 32, wrow0);
            tma_store_commit();
        }
        ++cc;
    }
}

}  // namespace gemm_detail
}  // namespace wg
