// Fused flash-style multi-head attention for the CLIP ViT-L/14@448 tower: head_dim 64,
// T = 1025 tokens, optional key-padding mask (custom_clip.py:27-38 semantics: masked KEYS never receive weight).
//
// DEFAULT: attention_d64_q4_kernel (bottom of the file) -- FOUR CTAs per SM, 64-key blocks, P written in place over S, 128 TMEM columns
// and 5 warps per CTA (B200, B=64, T=1025: 0.395 ms = 697 TFLOP/s; T=1024: 0.346 ms = 795 TFLOP/s).  The two-CTAs-per-SM kernels
// described next are what it replaced (0.51 ms); they stay reachable with WG_ATTN_Q4=0 as the A/B baseline and are still covered by the
// attention tests when that variable is set.
//
// Two CTAs co-reside per SM so that one CTA's softmax overlaps the other's tensor work.  From 4 key blocks per tile on they
// are persistent and walk a list of (image, head, 128-query tile) items (attention_d64_persist_kernel);
// shorter sequences get one CTA per item (attention_d64_kernel).  Per CTA:
//   warp 0  : TMA producer  (Q once, K/V blocks of 128 keys through 2-stage rings with separate K / V release;
//             3-D tensor map over [image, token, 3*heads*64] so rows past T are zero-filled by the hardware)
//   warp 1  : TMEM allocation + single-thread tcgen05.mma issue:  S = Q K^T (128x128x64, both operands from smem)
//             into TMEM;  O += P_j V_j (128x64x128) with P read FROM TENSOR MEMORY (A operand) and V consumed as an
//             MN-major smem operand straight from its TMA tile -- P never touches shared memory
//   warps 2-5: softmax warpgroup, ONE THREAD PER QUERY ROW (TMEM lane == row, so row max / row sum need no
//             shuffles): tcgen05.ld S -> online softmax in fp32 (exp2 with folded scale, packed FFMA2/FADD2, 3 of 8
//             exponentials on the FMA pipe) -> P as bf16 pairs back into TMEM (tcgen05.st); O stays in TMEM across
//             all key blocks (lazy rescale) -> final 1/l, bf16, TMA store.
//
// Measured on B200 (B=64, T=1025, 16 heads): 0.71 ms (P through smem, serial softmax phases, branchy key masking)
// -> 0.51 ms (T=1024: 0.43 ms).  What moved it: P in TMEM + packed arithmetic + software-pipelined exponentials, branch-free
// key masking only where keys can be invalid, narrow last key block, polynomial exp2 share.
#include "host.h"
#include "ptx.cuh"
#include <cstdlib>

namespace wg {
namespace {

constexpr int ATT_BQ = 128;      // queries per CTA
constexpr int ATT_BKV = 128;     // keys per block
constexpr int ATT_D = 64;        // head dim
// threads: warp 0 TMA, warp 1 MMA, then 4 * SPLIT softmax warps (SPLIT = threads per query row: 1, or 2 with the key block's
// columns divided between them -- four softmax warps per scheduler at two CTAs per SM instead of two)
__host__ __device__ constexpr int att_threads(int split) { return 64 + 128 * split; }
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: a [128 x 64] bf16 tile
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;          // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;      // 2 stages
constexpr int OFF_BAR = OFF_V + 2 * TILE_BYTES;
constexpr int ATT_NUM_BARS = 14;
constexpr int OFF_KMASK = OFF_BAR + ATT_NUM_BARS * 8 + 16;  // key-validity bit words (32 keys each), built once per CTA
constexpr int ATT_MAX_T = 8192;
constexpr int OFF_XCH = OFF_KMASK + (ATT_MAX_T / 32) * 4;  // SPLIT = 2: row max [2 parities][2 halves][128] + row sum [2][128] exchange
constexpr int ATT_SMEM_BYTES = OFF_XCH + (4 + 2) * 128 * 4;
constexpr int ATT_TMEM_COLS = 256;  // S: [0,128)  O: [128,192)  P (bf16 pairs): [192,256)
constexpr int ATT_SPLIT_DEFAULT = 1;
constexpr int ATT_Q4_DEFAULT = 1;
constexpr int ATT_Q4_POLY_DEFAULT = 0;  // the q4 kernel is no longer helped by polynomial exponentials (B200: 0 -> 0.399, 2 -> 0.399, 3 -> 0.411 ms)  // 1: attention_d64_q4_kernel (four CTAs per SM, 64-key blocks)
constexpr int ATT_POLY_DEFAULT = 3;  // of every 8 exponentials, evaluated on the FMA pipe (B200: 0 -> 0.556 ms, 3 -> 0.545 ms)
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units: the running scale is refreshed only when the row max grew by > 2^8

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): two lanes per issue slot
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// 2^x for a PAIR of arguments on the FMA/ALU pipes (no MUFU): x = j + f with j = round(x), f in [-0.5, 0.5]; a degree-3
// minimax polynomial gives 2^f (max rel. error 1.0e-4, below the bf16 rounding of P) and j goes into the exponent field.
// The MUFU unit evaluates one warp-wide ex2 per 8 cycles per scheduler, which bounds d=64 attention (ncu: XU pipe 70 % of
// peak over the whole kernel with every exponential on it), so a fixed fraction of the exponentials is evaluated here.
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& e0, float& e1) {
    x0 = fmaxf(x0, -126.0f);
    x1 = fmaxf(x1, -126.0f);
    const uint64_t X = pk2(x0, x1);
    const uint64_t MAGIC = pk2(12582912.0f, 12582912.0f);        // 1.5 * 2^23: the low mantissa bits of x + MAGIC hold round(x)
    const uint64_t NMAGIC = pk2(-12582912.0f, -12582912.0f);
    const uint64_t T = fadd2(X, MAGIC);
    const uint64_t F = ffma2(fadd2(T, NMAGIC), pk2(-1.0f, -1.0f), X);  // x - round(x)
    uint64_t P = ffma2(pk2(0.05500893f, 0.05500893f), F, pk2(0.24221096f, 0.24221096f));
    P = ffma2(P, F, pk2(0.69328293f, 0.69328293f));
    P = ffma2(P, F, pk2(1.0f, 1.0f));
    float p0, p1, t0, t1;
    upk2(P, p0, p1);
    upk2(T, t0, t1);
    e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
// which of the 8 column PAIRS of a 16-column window use the polynomial: POLY of 8 (0..4), spread between the MUFU ones
__device__ __forceinline__ constexpr bool use_poly(int pair, int poly) {
    return poly == 1 ? (pair == 3) : poly == 2 ? (pair == 1 || pair == 5) : poly == 3 ? (pair == 1 || pair == 4 || pair == 6)
         : poly == 4 ? ((pair & 1) == 1) : false;
}

// Debug-only pipeline trace (compile with -DATT_TRACE): clock64() stamps of three CTAs, read back by wg_debug_attn_trace.
#ifdef ATT_TRACE
__device__ long long g_att_trace[3][3][16][8];
#define TR(role, j, slot) do { if (tr >= 0 && (j) < 16) g_att_trace[tr][role][j][slot] = clock64(); } while (0)
#else
#define TR(role, j, slot) do { } while (0)
#endif

struct AttnParams {
    int T, heads, num_kv_blocks;
    int n_last;  // keys covered by the last key block, rounded up to 16 (T = 1025: 16 -- the block holds only the 1025th token)
    float scale_log2;
    const uint8_t* key_valid;  // [B, T] or null
    int n_items, n_full_tiles, batch;  // persistent kernel: work items = (image, head, query tile), full tiles first
    // T = 128 k + 1 (the CLIP tower: 1024 patches + CLS) in the persistent kernel: the key blocks cover the first 128 k keys and the
    // LAST key is scored once per tile on CUDA cores by the row's own softmax thread (extra_key = 1); it enters the initial maximum
    // and is folded into O and l in the epilogue, so the tile runs k instead of k + 1 softmax steps
    int extra_key;
    const __nv_bfloat16* qkv;
};

struct SoftmaxState {
    float m_ref;  // the maximum the current scale refers to
    float l_run;  // running row sum in that scale
    float x_acc;  // persistent kernel, extra_key: the last key's score, accumulated two 16-byte chunks per key block (blocks 0..3)
};

// One key block for one query row (= one thread): S row (NCH x 32 columns) from TMEM -> registers, hand S back, online softmax
// with the lazily refreshed scale, P row (bf16) -> smem.
//
// The exponential phase is software-pipelined in groups of 8 columns: the MUFU unit accepts one warp-wide ex2 every 8 cycles,
// so the row sum (FADD), the bf16 packing (F2FP) and the shared-memory stores of group g-1 are issued in the shadow of group
// g's exponentials instead of as serial phases after them (ncu showed the softmax warps spending 3/4 of their time outside
// the MUFU-paced stretch).  The wait for the previous PV MMA (P buffer free) sits in the middle of the phase: the first 64
// columns are packed into registers before it, so that the tensor core's latency is covered by exponentials, not by a stall.
template <int NCH, int POLY, int SPLIT>
__device__ __forceinline__ void softmax_block(SoftmaxState& st, int j, int gblk, const AttnParams& p, const uint32_t* kmask, uint32_t tmem_S, uint32_t tmem_O,
                                              uint32_t tmem_P, uint32_t lane_off, uint64_t* s_free, uint64_t* p_free, uint64_t* p_ready, int tr, int half, int r, float* xch,
                                              const uint8_t* xq = nullptr, const uint8_t* xk = nullptr) {
    // j = key block within the current query tile; gblk = key blocks this CTA has processed before it (barrier phases; == j in
    // the one-tile-per-CTA kernel, a running count in the persistent one)
    // NCH = number of 32-column chunks THIS thread handles (SPLIT = 2: the thread owns columns [64 half, 64 half + 32 NCH))
    constexpr int NG = NCH * 4;                  // groups of 8 columns
    constexpr int HALF_NG = 4 * (4 / SPLIT) / 2; // the P-buffer wait happens once this many groups are packed (half a full row share)
    constexpr int WAIT_AT = NG > HALF_NG ? HALF_NG : NG;
    uint32_t sv2[NCH > 0 ? NCH : 1][32];
#define sv(i) sv2[(i) >> 5][(i) & 31]
#pragma unroll
    for (int c = 0; c < NCH; ++c) tmem_ld_32x32b_x32(tmem_S + lane_off + half * 64 + c * 32, sv2[c]);
    if (xq != nullptr && j < 4) {
        // extra_key: a quarter of the dot product q_row . k_last in the shadow of the tensor-memory loads (xq: this row of the swizzled
        // Q tile, xk: the K row, both in shared memory; chunk `pos` of the Q row holds the row's 16-byte chunk pos ^ (r & 7))
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            const int pos = j * 2 + pp;
            const uint4 qa = *reinterpret_cast<const uint4*>(xq + pos * 16);
            const uint4 ka = *reinterpret_cast<const uint4*>(xk + ((pos ^ (r & 7)) * 16));
            const uint32_t qw[4] = {qa.x, qa.y, qa.z, qa.w}, kw[4] = {ka.x, ka.y, ka.z, ka.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a0 = fmaf(__uint_as_float(qw[e] << 16), __uint_as_float(kw[e] << 16), a0);
                a1 = fmaf(__uint_as_float(qw[e] & 0xffff0000u), __uint_as_float(kw[e] & 0xffff0000u), a1);
            }
        }
        st.x_acc += a0 + a1;
    }
    if (NCH > 0) tmem_ld_wait();
    TR(0, j, 2);
    tc_fence_before();
    mbar_arrive(s_free);  // S_j is in registers: the tensor core may overwrite it with S_{j+1}
    if (kmask != nullptr) {  // keys past T and masked keys: one validity word per 32 columns, branch-free selects
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const uint32_t word = kmask[j * (ATT_BKV / 32) + half * 2 + c];
            if (word != 0xffffffffu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) sv(c * 32 + i) = ((word >> i) & 1u) ? sv(c * 32 + i) : 0xff800000u;  // -inf
            }
        }
    }
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains (ILP)
#ifndef ATT_DBG_NOMAX  // debug builds (wrong results): knock one phase out to measure what it costs
#pragma unroll
    for (int i = 0; i < NCH * 32; i += 8) {  // 3-input maxima (FMNMX3): 4 chains, two new values per instruction
#pragma unroll
        for (int c = 0; c < 4; ++c) mx4[c] = fmaxf(fmaxf(mx4[c], __uint_as_float(sv(i + 2 * c))), __uint_as_float(sv(i + 2 * c + 1)));
    }
#else
    mx4[0] = 0.f;
#endif
    float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    if constexpr (SPLIT == 2) {
        // the two threads of a row agree on the block maximum through smem (double-buffered by block parity: one 64-thread
        // named barrier per block is enough); every later decision (reference max, rescale) is then identical in both
        float* slot = xch + (j & 1) * 256;
        slot[half * 128 + r] = mx;
        named_bar_sync(2 + (r >> 5), 64);
        mx = fmaxf(mx, slot[(half ^ 1) * 128 + r]);
    }
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
    const float scale = p.scale_log2;
    const float neg_m = -st.m_ref * scale;
    const uint64_t SC2 = pk2(scale, scale), NM2 = pk2(neg_m, neg_m);
    uint64_t rsA = pk2(0.f, 0.f), rsB = pk2(0.f, 0.f);  // row sum, two packed accumulators
#pragma unroll
    for (int g = 0; g <= NG; ++g) {
        if (g < NG) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // column pairs: packed scale-and-shift, then MUFU or polynomial
                const int i = g * 8 + k * 2;
                float x0, x1, e0, e1;
#ifndef ATT_DBG_NOSCALE
                upk2(ffma2(pk2(__uint_as_float(sv(i)), __uint_as_float(sv(i + 1))), SC2, NM2), x0, x1);  // -inf for masked keys -> 0
#else
                x0 = __uint_as_float(sv(i)); x1 = __uint_as_float(sv(i + 1));
#endif
#ifdef ATT_DBG_NOEXP  // debug build: no exponentials at all (wrong results) -- measures what the rest of the block costs
                if (true) { e0 = x0; e1 = x1; } else
#endif
                if (use_poly((g & 1) * 4 + k, POLY)) {
                    exp2_poly2(x0, x1, e0, e1);
                } else {
                    e0 = ex2_approx(x0);
                    e1 = ex2_approx(x1);
                }
                sv(i) = __float_as_uint(e0);
                sv(i + 1) = __float_as_uint(e1);
            }
        }
        if (g >= 1) {  // row sum + bf16 packing of the previous group, in place (4 packed words at its first 4 registers)
            const int b = (g - 1) * 8;
#ifndef ATT_DBG_NOSUM
            rsA = fadd2(rsA, pk2(__uint_as_float(sv(b + 0)), __uint_as_float(sv(b + 1))));
            rsB = fadd2(rsB, pk2(__uint_as_float(sv(b + 2)), __uint_as_float(sv(b + 3))));
            rsA = fadd2(rsA, pk2(__uint_as_float(sv(b + 4)), __uint_as_float(sv(b + 5))));
            rsB = fadd2(rsB, pk2(__uint_as_float(sv(b + 6)), __uint_as_float(sv(b + 7))));
#else
            if (b == 0) rsA = fadd2(rsA, pk2(__uint_as_float(sv(b + 0)), __uint_as_float(sv(b + 1))));
#endif
            const uint32_t w0 = pack_bf16x2(__uint_as_float(sv(b + 0)), __uint_as_float(sv(b + 1)));
            const uint32_t w1 = pack_bf16x2(__uint_as_float(sv(b + 2)), __uint_as_float(sv(b + 3)));
            const uint32_t w2 = pack_bf16x2(__uint_as_float(sv(b + 4)), __uint_as_float(sv(b + 5)));
            const uint32_t w3 = pack_bf16x2(__uint_as_float(sv(b + 6)), __uint_as_float(sv(b + 7)));
            sv(b + 0) = w0; sv(b + 1) = w1; sv(b + 2) = w2; sv(b + 3) = w3;
        }
        if (g == WAIT_AT) {
            if (j > 0) {
                TR(0, j, 3);
                mbar_wait(p_free, (gblk - 1) & 1);  // PV_{j-1} finished: P buffer free, O quiescent
                TR(0, j, 4);
                if (__any_sync(0xffffffffu, refresh)) {
                    tc_fence_after();
#pragma unroll
                    for (int cc = 0; cc < 2 / SPLIT; ++cc) {  // SPLIT = 2: each thread of the row rescales its 32 O columns
                        const int c = SPLIT == 2 ? half : cc;
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
        }
        // group q of 8 keys = 4 packed 32-bit TMEM columns of this thread's lane
        if (g == WAIT_AT) {
#pragma unroll
            for (int q = 0; q < WAIT_AT; ++q)
                tmem_st_32x32b_x4(tmem_P + lane_off + half * 32 + q * 4, sv(q * 8 + 0), sv(q * 8 + 1), sv(q * 8 + 2), sv(q * 8 + 3));
        } else if (g > WAIT_AT) {
            const int q = g - 1;
            tmem_st_32x32b_x4(tmem_P + lane_off + half * 32 + q * 4, sv(q * 8 + 0), sv(q * 8 + 1), sv(q * 8 + 2), sv(q * 8 + 3));
        }
    }
    float rs0, rs1;
    upk2(fadd2(rsA, rsB), rs0, rs1);
    st.l_run = st.l_run * alpha + (rs0 + rs1);
    if (NCH > 0) tmem_st_wait();  // P is in tensor memory
    tc_fence_before();
    mbar_arrive(p_ready);
    TR(0, j, 5);
#undef sv
}

// Pipeline (per CTA = one 128-query tile of one head of one image; two CTAs share an SM):
//   S_j = Q K_j^T (TMEM)  ->  softmax threads pull their S row into registers and immediately hand the S buffer back
//   (s_free) so the tensor core computes S_{j+1} while the exponentials of block j are evaluated  ->  P_j (bf16 pairs, TMEM)
//   ->  O += P_j V_j accumulated IN TMEM across all key blocks.  The softmax uses a lazily refreshed scale: the reference
//   maximum is only moved (and O / l rescaled through tcgen05.ld/st) when a row's maximum grew by more than 2^8, which is
//   exact arithmetic (any common scale cancels in O / l) and takes the O round trip off the per-block critical path.
template <int POLY, int SPLIT>
__global__ void __launch_bounds__(att_threads(SPLIT), 2)
attention_d64_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzled tiles need 1024B alignment

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* v_full = bars + 3;    // [2]
    uint64_t* k_empty = bars + 5;   // [2]  K stage free: committed right behind the S MMA that read it
    uint64_t* v_empty = bars + 7;   // [2]  V stage free: committed behind the PV MMA
    uint64_t* s_full = bars + 9;
    uint64_t* s_free = bars + 10;
    uint64_t* p_ready = bars + 11;
    uint64_t* p_free = bars + 12;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + ATT_NUM_BARS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qt = blockIdx.x;
    const int head = blockIdx.y;
    const int img = blockIdx.z;
    const int HD = p.heads * ATT_D;
    const int nkb = p.num_kv_blocks;
#ifdef ATT_TRACE
    int tr = -1;
    if (qt == 3 && head == 5 && (img == 0 || img == 20 || img == 40) && (threadIdx.x == 32 || threadIdx.x == 64)) tr = img / 20;
#else
    const int tr = -1;
#endif

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmO);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 128 * SPLIT);
        mbar_init(p_ready, 128 * SPLIT);
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
    const uint32_t tmem_P = tmem_base + 192;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
            tma_load_3d(smem + OFF_Q, &tmQKV, q_full, head * ATT_D, qt * ATT_BQ, img);
            // K and V stages are released separately (K right after its S MMA, V after its PV MMA) and the loads are issued
            // in the order K_{j+1}, V_j: every load goes out about two key blocks before the tensor core needs it.  (With one
            // combined release after PV_j, K_{j+2} could only be requested when S_{j+2} was already due -- the TMA round trip
            // was exposed once per key block and set the kernel's pace.)
            mbar_arrive_expect_tx(&k_full[0], TILE_BYTES);
            tma_load_3d(smem + OFF_K, &tmQKV, &k_full[0], HD + head * ATT_D, 0, img);
            for (int j = 0; j < nkb; ++j) {
                if (j + 1 < nkb) {
                    const int s1 = (j + 1) & 1;
                    mbar_wait_relaxed(&k_empty[s1], (((j + 1) >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&k_full[s1], TILE_BYTES);
                    tma_load_3d(smem + OFF_K + s1 * TILE_BYTES, &tmQKV, &k_full[s1], HD + head * ATT_D, (j + 1) * ATT_BKV, img);
                }
                const int s = j & 1;
                mbar_wait_relaxed(&v_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
                tma_load_3d(smem + OFF_V + s * TILE_BYTES, &tmQKV, &v_full[s], 2 * HD + head * ATT_D, j * ATT_BKV, img);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: the whole warp runs the protocol, one elected lane issues =====================
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, false, false);  // Q (K-major) x K (K-major)
        constexpr uint32_t IDESC_O = umma_idesc_bf16(128, 64, false, true);    // P (K-major) x V (MN-major)
        // the last key block only spans n_last (multiple of 16) keys: narrower S MMA, fewer PV k-steps
        const uint32_t idesc_s_last = umma_idesc_bf16(128, p.n_last, false, false);
        const int pv_steps_last = p.n_last / 16;
        // descriptors differ only in the 14-bit start-address field (bytes >> 4): build each once, then add
        const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + OFF_Q));
        const uint64_t k_desc0 = umma_desc_sw128(smem_u32(smem + OFF_K));
        const uint64_t v_desc0 = umma_desc_sw128(smem_u32(smem + OFF_V));
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t idesc = nkb == 1 ? idesc_s_last : IDESC_S;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, q_desc + k * 2, k_desc0 + k * 2, idesc, k != 0);
            umma_commit(s_full);
            umma_commit(&k_empty[0]);
        }
        __syncwarp();
        for (int j = 0; j < nkb; ++j) {
            const int s = j & 1;
            if (j + 1 < nkb) {
                // S_{j+1} as soon as the softmax threads have pulled S_j out of TMEM: overlaps their exponentials
                const int s1 = (j + 1) & 1;
                TR(1, j, 0);
                mbar_wait(s_free, j & 1);
                TR(1, j, 1);
                mbar_wait(&k_full[s1], ((j + 1) >> 1) & 1);
                TR(1, j, 2);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t k_desc = k_desc0 + s1 * (TILE_BYTES >> 4);
                    const uint32_t idesc = (j + 2 == nkb) ? idesc_s_last : IDESC_S;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, q_desc + k * 2, k_desc + k * 2, idesc, k != 0);
                    umma_commit(s_full);
                    umma_commit(&k_empty[s1]);
                }
                __syncwarp();
            }
            TR(1, j, 3);
            mbar_wait(p_ready, j & 1);
            TR(1, j, 4);
            mbar_wait(&v_full[s], (j >> 1) & 1);
            TR(1, j, 5);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t v_desc = v_desc0 + s * (TILE_BYTES >> 4);
                const int pv_steps = (j + 1 == nkb) ? pv_steps_last : 8;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // A: P from tensor memory, 16 keys = 8 packed columns per step;  B: V rows (keys) step 16 rows = +2048 B
                    if (k < pv_steps) umma_f16_ts(tmem_O, tmem_P + k * 8, v_desc + k * (2048 >> 4), IDESC_O, (j | k) != 0);
                }
                umma_commit(&v_empty[s]);
                umma_commit(p_free);  // P buffer reusable, O consistent
            }
            __syncwarp();
            TR(1, j, 6);
        }
    } else {
        // ===================== softmax warps: SPLIT threads per query row =====================
        // SPLIT = 2 puts four softmax warps on every scheduler (two CTAs per SM); the two threads of a row exchange only the
        // block maximum and the final row sum.  Measured on B200 (ncu, B=64 T=1025): issue slots 52 % -> 64 % busy, eligible
        // warps 0.7 -> 1.4 per cycle, but 23 % more instructions and the same 0.47 ms -- kept as a tuning knob, default 1.
        const int quarter = warp & 3;            // TMEM lane quarter this warp may touch
        const int half = (warp - 2) >> 2;        // which 64 columns of a key block this thread owns (always 0 for SPLIT = 1)
        const int r = quarter * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        float* xch = reinterpret_cast<float*>(smem + OFF_XCH);
        // key-validity words (keys past T, masked keys), built once: the per-block softmax only does selects on them
        const uint32_t* kmask = nullptr;
        if (p.key_valid != nullptr || (p.T % ATT_BKV) != 0) {
            uint32_t* words = reinterpret_cast<uint32_t*>(smem + OFF_KMASK);
            const uint8_t* kvalid = p.key_valid ? p.key_valid + (size_t)img * p.T : nullptr;
            for (int w = warp - 2; w < nkb * (ATT_BKV / 32); w += 4 * SPLIT) {
                const int key = w * 32 + lane;
                const bool ok = key < p.T && (kvalid == nullptr || kvalid[key] != 0);
                const uint32_t word = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) words[w] = word;
            }
            named_bar_sync(1, 128 * SPLIT);
            kmask = words;
        }
        // warps whose 32 query rows all lie past T (last query tile: only the 1025th token is real) skip the arithmetic and only
        // keep the barrier protocol alive; their P / O rows are never stored (the TMA store clips at T).
        const bool warp_active = qt * ATT_BQ + quarter * 32 < p.T;
        SoftmaxState st{0.f, 0.f, 0.f};
        for (int j = 0; j < nkb; ++j) {
            TR(0, j, 0);
            mbar_wait(s_full, j & 1);
            TR(0, j, 1);
            tc_fence_after();
            if (!warp_active) {
                tc_fence_before();
                mbar_arrive(s_free);
                if (j > 0) mbar_wait(p_free, (j - 1) & 1);
                mbar_arrive(p_ready);
                continue;
            }
            const int ncols = (j + 1 == nkb) ? p.n_last : ATT_BKV;
            const uint32_t* km = (p.key_valid != nullptr || j + 1 == nkb) ? kmask : nullptr;  // only blocks that can hold invalid keys
            const int nch_tot = (ncols + 31) >> 5;                       // 32-column chunks that hold keys in this block
            int nch = nch_tot - half * 2;                                // ... of which this thread owns
            nch = nch < 0 ? 0 : (nch > 4 / SPLIT ? 4 / SPLIT : nch);
#define WG_SM_ARGS st, j, j, p, km, tmem_S, tmem_O, tmem_P, lane_off, s_free, p_free, p_ready, tr, half, r, xch
            if constexpr (SPLIT == 1) {
                if (nch == 4) softmax_block<4, POLY, 1>(WG_SM_ARGS);
                else if (nch == 3) softmax_block<3, POLY, 1>(WG_SM_ARGS);
                else if (nch == 2) softmax_block<2, POLY, 1>(WG_SM_ARGS);
                else softmax_block<1, POLY, 1>(WG_SM_ARGS);
            } else {
                if (nch == 2) softmax_block<2, POLY, 2>(WG_SM_ARGS);
                else if (nch == 1) softmax_block<1, POLY, 2>(WG_SM_ARGS);
                else softmax_block<0, POLY, 2>(WG_SM_ARGS);
            }
#undef WG_SM_ARGS
        }
        float l_run = st.l_run;
        if constexpr (SPLIT == 2) {  // row sum = the two threads' partial sums (same scale: the reference max is common)
            float* ls = xch + 512;
            ls[half * 128 + r] = l_run;
            named_bar_sync(2 + quarter, 64);
            l_run += ls[(half ^ 1) * 128 + r];
        }
        // ---- epilogue: O / l, bf16, stage into the (now idle) Q tile, TMA store
        mbar_wait(p_free, (nkb - 1) & 1);
        tc_fence_after();
        const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
        uint8_t* o_row = smem + OFF_Q + r * 128;
#pragma unroll
        for (int cc = 0; cc < 2 / SPLIT; ++cc) {
            if (!warp_active) break;
            const int c = SPLIT == 2 ? half : cc;  // SPLIT = 2: each thread of the row normalises its 32 O columns
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
        named_bar_sync(1, 128 * SPLIT);
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

// ---------------------------------------------------------------------------------------------------------------------------
// Persistent variant: 2 CTAs per SM, each walking a strided list of (image, head, query tile) items with barriers, tensor memory
// and the K/V rings kept alive across items.  A one-tile CTA costs ~2.1 us of launch, TMEM allocation, first Q/K load latency and
// O drain next to 1.6 us per key block (14 % of the kernel at the CLIP length); here the next item's Q and first K block are
// requested while the current item's last key blocks are processed, S_0 of the next item is issued as soon as the last S has been
// pulled into registers, and the O tile leaves through the (finished) Q buffer while the next item's first key block is computed.
// Two Q buffers: Q_it lives in buffer it & 1, which then stages O_it; it is handed back to the producer (q_free) by the storing
// thread once the bulk store has read it, one item later.  Needs >= 4 key blocks per tile (the hand-backs assume it).
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int PQ_OFF_Q = 0;                              // 2 buffers
constexpr int PQ_OFF_K = PQ_OFF_Q + 2 * TILE_BYTES;      // 2 stages
constexpr int PQ_OFF_V = PQ_OFF_K + 2 * TILE_BYTES;      // 2 stages
constexpr int PQ_OFF_BAR = PQ_OFF_V + 2 * TILE_BYTES;
constexpr int PQ_NUM_BARS = 16;
constexpr int PQ_OFF_KMASK = PQ_OFF_BAR + PQ_NUM_BARS * 8 + 16;
constexpr int PQ_OFF_XKV = PQ_OFF_KMASK + (ATT_MAX_T / 32) * 4;  // extra_key: [2 Q buffers][K row | V row] of the last key, 128 B each
constexpr int PQ_SMEM_BYTES = PQ_OFF_XKV + 2 * 256;

struct AttnItem {
    int qt, head, img;
};
__device__ __forceinline__ AttnItem attn_item(const AttnParams& p, int w) {
    // Items of one (image, head) are CONSECUTIVE -- its full tiles, then its partial last tile -- so that the ~300 CTAs running at
    // any moment work on a few dozen (image, head) pairs whose K / V (262 KB each at the CLIP length) stay in L2.  (Round 1 put every
    // partial tile at the end of the schedule: each of them re-read its K / V from DRAM long after its siblings had finished --
    // 1.5x the algorithmic DRAM traffic in the ncu capture.)
    AttnItem it;
    const int per = p.n_full_tiles + ((p.T % ATT_BQ) != 0 ? 1 : 0);
    it.qt = w % per;
    const int r = w / per;
    it.head = r % p.heads;
    it.img = r / p.heads;
    return it;
}

template <int POLY>
__global__ void __launch_bounds__(192, 2)
attention_d64_persist_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PQ_OFF_BAR);
    uint64_t* q_full = bars + 0;    // [2]
    uint64_t* q_free = bars + 2;    // [2]  O tile staged in this Q buffer has been read by its bulk store
    uint64_t* k_full = bars + 4;    // [2]
    uint64_t* v_full = bars + 6;    // [2]
    uint64_t* k_empty = bars + 8;   // [2]
    uint64_t* v_empty = bars + 10;  // [2]
    uint64_t* s_full = bars + 12;
    uint64_t* s_free = bars + 13;
    uint64_t* p_ready = bars + 14;
    uint64_t* p_free = bars + 15;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + PQ_NUM_BARS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int HD = p.heads * ATT_D;
    const int nkb = p.num_kv_blocks;
    const int stride = gridDim.x;
    const int tr = -1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmO);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&q_full[s], 1);
            mbar_init(&q_free[s], 1);
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
    if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_S = tmem_base;
    const uint32_t tmem_O = tmem_base + 128;
    const uint32_t tmem_P = tmem_base + 192;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int g = 0, it = 0;
            for (int w = blockIdx.x; w < p.n_items; w += stride, ++it) {
                const AttnItem cur = attn_item(p, w);
                const bool has_next = w + stride < p.n_items;
                const AttnItem nxt = has_next ? attn_item(p, w + stride) : cur;
                if (it == 0) {
                    mbar_arrive_expect_tx(&q_full[0], TILE_BYTES + (p.extra_key ? 256 : 0));
                    tma_load_3d(smem + PQ_OFF_Q, &tmQKV, &q_full[0], cur.head * ATT_D, cur.qt * ATT_BQ, cur.img);
                    if (p.extra_key) {
                        const __nv_bfloat16* xrow = p.qkv + ((size_t)cur.img * p.T + (p.T - 1)) * (3 * HD) + cur.head * ATT_D;
                        bulk_load_1d(smem + PQ_OFF_XKV, xrow + HD, 128, &q_full[0]);
                        bulk_load_1d(smem + PQ_OFF_XKV + 128, xrow + 2 * HD, 128, &q_full[0]);
                    }
                    mbar_arrive_expect_tx(&k_full[0], TILE_BYTES);
                    tma_load_3d(smem + PQ_OFF_K, &tmQKV, &k_full[0], HD + cur.head * ATT_D, 0, cur.img);
                }
                for (int j = 0; j < nkb; ++j, ++g) {
                    // K of the next key block of this CTA's stream: this tile's block j + 1, or block 0 of the next item
                    const bool last = j + 1 == nkb;
                    if (!last || has_next) {
                        const int s1 = (g + 1) & 1;
                        mbar_wait_relaxed(&k_empty[s1], (((g + 1) >> 1) & 1) ^ 1);
                        mbar_arrive_expect_tx(&k_full[s1], TILE_BYTES);
                        if (!last) tma_load_3d(smem + PQ_OFF_K + s1 * TILE_BYTES, &tmQKV, &k_full[s1], HD + cur.head * ATT_D, (j + 1) * ATT_BKV, cur.img);
                        else tma_load_3d(smem + PQ_OFF_K + s1 * TILE_BYTES, &tmQKV, &k_full[s1], HD + nxt.head * ATT_D, 0, nxt.img);
                    }
                    const int s = g & 1;
                    mbar_wait_relaxed(&v_empty[s], ((g >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
                    tma_load_3d(smem + PQ_OFF_V + s * TILE_BYTES, &tmQKV, &v_full[s], 2 * HD + cur.head * ATT_D, j * ATT_BKV, cur.img);
                    if (last && has_next) {
                        // Q of the next item into the other buffer; from the third item on that buffer last staged an O tile
                        const int qb = (it + 1) & 1;
                        if (it + 1 >= 2) mbar_wait_relaxed(&q_free[qb], (((it + 1) >> 1) - 1) & 1);
                        mbar_arrive_expect_tx(&q_full[qb], TILE_BYTES + (p.extra_key ? 256 : 0));
                        tma_load_3d(smem + PQ_OFF_Q + qb * TILE_BYTES, &tmQKV, &q_full[qb], nxt.head * ATT_D, nxt.qt * ATT_BQ, nxt.img);
                        if (p.extra_key) {  // the last key's K and V rows travel with the Q tile
                            const __nv_bfloat16* xrow = p.qkv + ((size_t)nxt.img * p.T + (p.T - 1)) * (3 * HD) + nxt.head * ATT_D;
                            bulk_load_1d(smem + PQ_OFF_XKV + qb * 256, xrow + HD, 128, &q_full[qb]);
                            bulk_load_1d(smem + PQ_OFF_XKV + qb * 256 + 128, xrow + 2 * HD, 128, &q_full[qb]);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform protocol, one elected lane issues) =====================
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, false, false);
        constexpr uint32_t IDESC_O = umma_idesc_bf16(128, 64, false, true);
        const uint32_t idesc_s_last = umma_idesc_bf16(128, p.n_last, false, false);
        const int pv_steps_last = p.n_last / 16;
        const uint64_t q_desc0 = umma_desc_sw128(smem_u32(smem + PQ_OFF_Q));
        const uint64_t k_desc0 = umma_desc_sw128(smem_u32(smem + PQ_OFF_K));
        const uint64_t v_desc0 = umma_desc_sw128(smem_u32(smem + PQ_OFF_V));
        int g = 0, it = 0;
        for (int w = blockIdx.x; w < p.n_items; w += stride, ++it) {
            const bool has_next = w + stride < p.n_items;
            if (it == 0) {
                mbar_wait(&q_full[0], 0);
                mbar_wait(&k_full[0], 0);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, q_desc0 + k * 2, k_desc0 + k * 2, IDESC_S, k != 0);
                    umma_commit(s_full);
                    umma_commit(&k_empty[0]);
                }
                __syncwarp();
            }
            for (int j = 0; j < nkb; ++j, ++g) {
                const bool last = j + 1 == nkb;
                if (!last || has_next) {
                    // S of the next key block of the stream as soon as the softmax threads have pulled the current S out of TMEM
                    const int s1 = (g + 1) & 1;
                    mbar_wait(s_free, g & 1);
                    const int qb = last ? (it + 1) & 1 : it & 1;
                    if (last) mbar_wait(&q_full[qb], ((it + 1) >> 1) & 1);
                    mbar_wait(&k_full[s1], ((g + 1) >> 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t q_desc = q_desc0 + qb * (TILE_BYTES >> 4);
                        const uint64_t k_desc = k_desc0 + s1 * (TILE_BYTES >> 4);
                        const uint32_t idesc = (!last && j + 2 == nkb) ? idesc_s_last : IDESC_S;  // nkb >= 4: block 0 of a tile is full
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, q_desc + k * 2, k_desc + k * 2, idesc, k != 0);
                        umma_commit(s_full);
                        umma_commit(&k_empty[s1]);
                    }
                    __syncwarp();
                }
                const int s = g & 1;
                mbar_wait(p_ready, g & 1);  // (for j == 0 this also says: the softmax threads have drained the previous tile's O)
                mbar_wait(&v_full[s], (g >> 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t v_desc = v_desc0 + s * (TILE_BYTES >> 4);
                    const int pv_steps = last ? pv_steps_last : 8;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (k < pv_steps) umma_f16_ts(tmem_O, tmem_P + k * 8, v_desc + k * (2048 >> 4), IDESC_O, (j | k) != 0);
                    }
                    umma_commit(&v_empty[s]);
                    umma_commit(p_free);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== softmax warps: one thread per query row =====================
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        uint32_t* words = reinterpret_cast<uint32_t*>(smem + PQ_OFF_KMASK);
        const bool need_mask = p.key_valid != nullptr || ((p.T % ATT_BKV) != 0 && !p.extra_key);
        int g = 0, it = 0, mask_img = -1;
        for (int w = blockIdx.x; w < p.n_items; w += stride, ++it) {
            const AttnItem cur = attn_item(p, w);
            const uint32_t* kmask = nullptr;
            if (need_mask) {
                // validity words: per image when there is a key mask, once otherwise (keys past T)
                if (mask_img < 0 || (p.key_valid != nullptr && mask_img != cur.img)) {
                    if (mask_img >= 0) named_bar_sync(1, 128);  // nobody still reads the previous image's words
                    const uint8_t* kvalid = p.key_valid ? p.key_valid + (size_t)cur.img * p.T : nullptr;
                    for (int x = warp - 2; x < nkb * (ATT_BKV / 32); x += 4) {
                        const int key = x * 32 + lane;
                        const bool ok = key < p.T && (kvalid == nullptr || kvalid[key] != 0);
                        const uint32_t word = __ballot_sync(0xffffffffu, ok);
                        if (lane == 0) words[x] = word;
                    }
                    named_bar_sync(1, 128);
                    mask_img = cur.img;
                }
                kmask = words;
            }
            const bool warp_active = cur.qt * ATT_BQ + quarter * 32 < p.T;
            SoftmaxState st{0.f, 0.f, 0.f};
            // extra_key: the last key's K row | V row arrived with the Q tile; its score is accumulated inside blocks 0..3
            const uint8_t* xkv = smem + PQ_OFF_XKV + (it & 1) * 256;
            const uint8_t* xq = nullptr;
            bool x_valid = false;
            if (p.extra_key && warp_active) {
                x_valid = p.key_valid == nullptr || p.key_valid[(size_t)cur.img * p.T + (p.T - 1)] != 0;
                if (x_valid) {
                    mbar_wait(&q_full[it & 1], (it >> 1) & 1);  // long complete (S_0 of this item is already being computed)
                    xq = smem + PQ_OFF_Q + (it & 1) * TILE_BYTES + r * 128;
                }
            }
            for (int j = 0; j < nkb; ++j, ++g) {
                mbar_wait(s_full, g & 1);
                tc_fence_after();
                if (!warp_active) {
                    tc_fence_before();
                    mbar_arrive(s_free);
                    if (j > 0) mbar_wait(p_free, (g - 1) & 1);
                    mbar_arrive(p_ready);
                } else {
                    const int ncols = (j + 1 == nkb) ? p.n_last : ATT_BKV;
                    const uint32_t* km = (p.key_valid != nullptr || j + 1 == nkb) ? kmask : nullptr;
                    const int nch = (ncols + 31) >> 5;
#define WG_SM_ARGS st, j, g, p, km, tmem_S, tmem_O, tmem_P, lane_off, s_free, p_free, p_ready, tr, 0, r, nullptr, xq, xkv
                    if (nch == 4) softmax_block<4, POLY, 1>(WG_SM_ARGS);
                    else if (nch == 3) softmax_block<3, POLY, 1>(WG_SM_ARGS);
                    else if (nch == 2) softmax_block<2, POLY, 1>(WG_SM_ARGS);
                    else softmax_block<1, POLY, 1>(WG_SM_ARGS);
#undef WG_SM_ARGS
                }
                if (j == 1 && it > 0 && threadIdx.x == 64) {
                    // the previous item's O tile has long left its staging buffer: hand that Q buffer back to the producer
                    tma_store_wait_read<0>();
                    mbar_arrive(&q_free[(it - 1) & 1]);
                }
            }
            // ---- epilogue: O / l, bf16, staged in this item's (finished) Q buffer, bulk store left in flight
            mbar_wait(p_free, (g - 1) & 1);
            tc_fence_after();
            // the extra key joins here: common maximum m_fin, O and l brought to it (a_x = 1 unless the last key's score exceeds the
            // running reference), weight p_x <= 1
            float a_x = 1.f, p_x = 0.f;
            if (x_valid) {
                const float d = (st.x_acc - st.m_ref) * p.scale_log2;
                if (st.l_run > 0.f) {
                    a_x = d > 0.f ? ex2_approx(-d) : 1.f;
                    p_x = d > 0.f ? 1.f : ex2_approx(d);
                } else {  // every key of the blocks is masked: the extra key is the whole row
                    a_x = 0.f;
                    p_x = 1.f;
                }
            }
            const float l_fin = st.l_run * a_x + p_x;
            const float inv0 = l_fin > 0.f ? 1.0f / l_fin : 0.f;
            const float inv_l = inv0 * a_x, w_x = inv0 * p_x;
            uint8_t* o_tile = smem + PQ_OFF_Q + (it & 1) * TILE_BYTES;
            uint8_t* o_row = o_tile + r * 128;
            if (warp_active) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t ov[32];
                    tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, ov);
                    tmem_ld_wait();
                    const float fin = x_valid ? 1.f : inv_l;
                    if (x_valid) {  // O * inv_l + w_x * V[last key] here; the packing below then scales by 1
                        const uint4* v16 = reinterpret_cast<const uint4*>(xkv + 128) + c * 4;
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const uint4 va = v16[q4];
                            const uint32_t vw[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                ov[q4 * 8 + 2 * e] = __float_as_uint(fmaf(w_x, __uint_as_float(vw[e] << 16), __uint_as_float(ov[q4 * 8 + 2 * e]) * inv_l));
                                ov[q4 * 8 + 2 * e + 1] = __float_as_uint(fmaf(w_x, __uint_as_float(vw[e] & 0xffff0000u), __uint_as_float(ov[q4 * 8 + 2 * e + 1]) * inv_l));
                            }
                        }
                    }
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 0]) * fin, __uint_as_float(ov[q4 * 8 + 1]) * fin);
                        u.y = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 2]) * fin, __uint_as_float(ov[q4 * 8 + 3]) * fin);
                        u.z = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 4]) * fin, __uint_as_float(ov[q4 * 8 + 5]) * fin);
                        u.w = pack_bf16x2(__uint_as_float(ov[q4 * 8 + 6]) * fin, __uint_as_float(ov[q4 * 8 + 7]) * fin);
                        *reinterpret_cast<uint4*>(o_row + (((c * 4 + q4) ^ (r & 7)) * 16)) = u;
                    }
                }
            }
            tc_fence_before();  // the next tile's first PV overwrites O: ordered behind these loads through p_ready
            fence_proxy_async_smem();
            named_bar_sync(2, 128);
            if (threadIdx.x == 64) {
                tma_store_3d(&tmO, o_tile, cur.head * ATT_D, cur.qt * ATT_BQ, cur.img);
                tma_store_commit();
            }
        }
        if (threadIdx.x == 64) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Four CTAs per SM (round 2, fourth session).  The kernels above are bound by the per-key-block dependency chain
// (S ready -> tcgen05.ld -> max -> exp -> P -> PV), not by a pipe: two co-resident CTAs give every scheduler two softmax warps,
// and a lone CTA needs ~2800 cycles per 128-key block of which the MUFU unit is busy for < 1000.  This variant buys more
// independent chains per scheduler instead of shortening one: 64-key blocks need only 64 fp32 S columns, P (bf16 pairs, 32
// columns) is written IN PLACE over the S columns its thread has already pulled into registers, O takes 64 columns -> 128 TMEM
// columns per CTA, 64 live S registers per softmax thread, 50 KB of shared memory -> FOUR CTAs per SM (four softmax warps per
// scheduler).  Because P aliases S, the tensor pipe's in-order execution is the only ordering needed between PV_j and S_{j+1}
// (issued back to back by the same thread), and one commit behind S_{j+1} tells the softmax threads both "S_{j+1} is there" and
// "PV_j is done, O is quiescent": no s_free / p_free barriers.  The same warp issues the TMA loads: K stage s is free once S_j has
// completed and V stage s once PV_j has -- both are implied by p_ready(j) / p_ready(j+1) -- so there are no empty barriers either,
// and the CTA is 5 warps (160 threads, 102 registers per thread at four CTAs per SM).
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int Q4_BKV = 64;
constexpr int Q4_KV_BYTES = Q4_BKV * ATT_D * 2;  // 8 KB: a [64 x 64] bf16 tile
constexpr int Q4_OFF_Q = 0;
constexpr int Q4_OFF_K = Q4_OFF_Q + TILE_BYTES;        // 2 stages
constexpr int Q4_OFF_V = Q4_OFF_K + 2 * Q4_KV_BYTES;   // 2 stages
constexpr int Q4_OFF_BAR = Q4_OFF_V + 2 * Q4_KV_BYTES;
constexpr int Q4_NUM_BARS = 8;
constexpr int Q4_OFF_KMASK = Q4_OFF_BAR + Q4_NUM_BARS * 8 + 16;
constexpr int Q4_SMEM_BYTES = Q4_OFF_KMASK + (ATT_MAX_T / 32) * 4;
constexpr int Q4_TMEM_COLS = 128;  // S: [0,64) (P bf16 pairs over [0,32))  O: [64,128)
constexpr int Q4_THREADS = 160;    // warp 0: TMA + MMA issue, warps 1..4: softmax (TMEM lane quarter = warp & 3)

// One 64-key block for one query row: S row -> registers, online softmax with the lazily refreshed scale, P row (bf16 pairs) back
// into the first columns of the S row.  PV_{j-1} has completed before S_j became visible, so O may be rescaled without a wait.
template <int NCH, int POLY>
__device__ __forceinline__ void softmax_block_q4(SoftmaxState& st, int j, const AttnParams& p, const uint32_t* kmask, uint32_t tS, uint32_t tO,
                                                 uint64_t* p_ready) {
    constexpr int NG = NCH * 4;  // groups of 8 columns
    uint32_t sv2[NCH][32];
#define sv(i) sv2[(i) >> 5][(i) & 31]
#pragma unroll
    for (int c = 0; c < NCH; ++c) tmem_ld_32x32b_x32(tS + c * 32, sv2[c]);
    tmem_ld_wait();
    if (kmask != nullptr) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const uint32_t word = kmask[j * (Q4_BKV / 32) + c];
            if (word != 0xffffffffu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) sv(c * 32 + i) = ((word >> i) & 1u) ? sv(c * 32 + i) : 0xff800000u;  // -inf
            }
        }
    }
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < NCH * 32; i += 8) {
#pragma unroll
        for (int c = 0; c < 4; ++c) mx4[c] = fmaxf(fmaxf(mx4[c], __uint_as_float(sv(i + 2 * c))), __uint_as_float(sv(i + 2 * c + 1)));
    }
    const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    float alpha = 1.0f;
    bool refresh = false;
    if (j == 0) {
        st.m_ref = (mx == -INFINITY) ? 0.f : mx;
    } else if ((mx - st.m_ref) * p.scale_log2 > RESCALE_THRESHOLD) {
        alpha = ex2_approx((st.m_ref - mx) * p.scale_log2);
        st.m_ref = mx;
        refresh = true;
    }
    if (j > 0 && __any_sync(0xffffffffu, refresh)) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {  // 16 columns at a time: the S row occupies 64 registers
            uint32_t ov[16];
            tmem_ld_32x32b_x16(tO + c * 16, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x32b_x16(tO + c * 16, ov);
        }
    }
    const float scale = p.scale_log2;
    const float neg_m = -st.m_ref * scale;
    const uint64_t SC2 = pk2(scale, scale), NM2 = pk2(neg_m, neg_m);
    uint64_t rsA = pk2(0.f, 0.f), rsB = pk2(0.f, 0.f);
#pragma unroll
    for (int g = 0; g <= NG; ++g) {
        if (g < NG) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = g * 8 + k * 2;
                float x0, x1, e0, e1;
                upk2(ffma2(pk2(__uint_as_float(sv(i)), __uint_as_float(sv(i + 1))), SC2, NM2), x0, x1);
                if (use_poly((g & 1) * 4 + k, POLY)) {
                    exp2_poly2(x0, x1, e0, e1);
                } else {
                    e0 = ex2_approx(x0);
                    e1 = ex2_approx(x1);
                }
                sv(i) = __float_as_uint(e0);
                sv(i + 1) = __float_as_uint(e1);
            }
        }
        if (g >= 1) {  // row sum + bf16 packing + TMEM store of the previous group, in the shadow of this group's exponentials
            const int b = (g - 1) * 8;
            rsA = fadd2(rsA, pk2(__uint_as_float(sv(b + 0)), __uint_as_float(sv(b + 1))));
            rsB = fadd2(rsB, pk2(__uint_as_float(sv(b + 2)), __uint_as_float(sv(b + 3))));
            rsA = fadd2(rsA, pk2(__uint_as_float(sv(b + 4)), __uint_as_float(sv(b + 5))));
            rsB = fadd2(rsB, pk2(__uint_as_float(sv(b + 6)), __uint_as_float(sv(b + 7))));
            const uint32_t w0 = pack_bf16x2(__uint_as_float(sv(b + 0)), __uint_as_float(sv(b + 1)));
            const uint32_t w1 = pack_bf16x2(__uint_as_float(sv(b + 2)), __uint_as_float(sv(b + 3)));
            const uint32_t w2 = pack_bf16x2(__uint_as_float(sv(b + 4)), __uint_as_float(sv(b + 5)));
            const uint32_t w3 = pack_bf16x2(__uint_as_float(sv(b + 6)), __uint_as_float(sv(b + 7)));
            tmem_st_32x32b_x4(tS + (g - 1) * 4, w0, w1, w2, w3);
        }
    }
    float rs0, rs1;
    upk2(fadd2(rsA, rsB), rs0, rs1);
    st.l_run = st.l_run * alpha + (rs0 + rs1);
    tmem_st_wait();  // P (and a rescaled O) are in tensor memory
    tc_fence_before();
    mbar_arrive(p_ready);
#undef sv
}

template <int POLY>
__global__ void __launch_bounds__(Q4_THREADS, 4)
attention_d64_q4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmO,
                        const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q4_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* v_full = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;   // S_j visible (and PV_{j-1} complete); its last phase: the final O is complete
    uint64_t* p_ready = bars + 6;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Q4_NUM_BARS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
    const int HD = p.heads * ATT_D;
    const int nkb = p.num_kv_blocks;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmKV);
            tma_prefetch_desc(&tmO);
            mbar_init(q_full, 1);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&k_full[s], 1);
                mbar_init(&v_full[s], 1);
            }
            mbar_init(s_full, 1);
            mbar_init(p_ready, 128);
            fence_mbar_init();
            // the first loads go out before the tensor-memory allocation and the CTA-wide barrier: their round trip overlaps the set-up
            const int kcol0 = HD + head * ATT_D, vcol0 = 2 * HD + head * ATT_D;
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
            tma_load_3d(smem + Q4_OFF_Q, &tmQ, q_full, head * ATT_D, qt * ATT_BQ, img);
            mbar_arrive_expect_tx(&k_full[0], Q4_KV_BYTES);
            tma_load_3d(smem + Q4_OFF_K, &tmKV, &k_full[0], kcol0, 0, img);
            mbar_arrive_expect_tx(&v_full[0], Q4_KV_BYTES);
            tma_load_3d(smem + Q4_OFF_V, &tmKV, &v_full[0], vcol0, 0, img);
            if (nkb > 1) {
                mbar_arrive_expect_tx(&k_full[1], Q4_KV_BYTES);
                tma_load_3d(smem + Q4_OFF_K + Q4_KV_BYTES, &tmKV, &k_full[1], kcol0, Q4_BKV, img);
            }
        }
        __syncwarp();
        tmem_alloc<Q4_TMEM_COLS>(tmem_ptr_smem);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_S = tmem_base;       // P over its first 32 columns
    const uint32_t tmem_O = tmem_base + 64;

    if (warp == 0) {
        // ===================== TMA loads + MMA issue (warp-uniform protocol, one elected lane issues) =====================
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, Q4_BKV, false, false);  // Q (K-major) x K (K-major)
        constexpr uint32_t IDESC_O = umma_idesc_bf16(128, 64, false, true);       // P (TMEM, K-major) x V (MN-major)
        const uint32_t idesc_s_last = umma_idesc_bf16(128, p.n_last, false, false);
        const int pv_steps_last = p.n_last / 16;
        const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + Q4_OFF_Q));
        const uint64_t k_desc0 = umma_desc_sw128(smem_u32(smem + Q4_OFF_K));
        const uint64_t v_desc0 = umma_desc_sw128(smem_u32(smem + Q4_OFF_V));
        const int kcol = HD + head * ATT_D, vcol = 2 * HD + head * ATT_D;
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t idesc = nkb == 1 ? idesc_s_last : IDESC_S;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, q_desc + k * 2, k_desc0 + k * 2, idesc, k != 0);
            umma_commit(s_full);
        }
        __syncwarp();
        for (int j = 0; j < nkb; ++j) {
            const int s = j & 1, s1 = s ^ 1;
            const bool last = j + 1 == nkb;
            // S_j visible => S_j and PV_{j-1} are complete: K stage s and V stage s1 are free.  Their next loads (V_{j+1}, K_{j+2}) go out
            // here, a whole softmax block before p_ready(j), i.e. about two block periods before the tensor core needs them.
            mbar_wait(s_full, j & 1);
            if (elect_one()) {
                if (!last) {
                    mbar_arrive_expect_tx(&v_full[s1], Q4_KV_BYTES);
                    tma_load_3d(smem + Q4_OFF_V + s1 * Q4_KV_BYTES, &tmKV, &v_full[s1], vcol, (j + 1) * Q4_BKV, img);
                }
                if (j + 2 < nkb) {
                    mbar_arrive_expect_tx(&k_full[s], Q4_KV_BYTES);
                    tma_load_3d(smem + Q4_OFF_K + s * Q4_KV_BYTES, &tmKV, &k_full[s], kcol, (j + 2) * Q4_BKV, img);
                }
            }
            __syncwarp();
            if (!last) mbar_wait(&k_full[s1], ((j + 1) >> 1) & 1);
            mbar_wait(&v_full[s], (j >> 1) & 1);
            mbar_wait(p_ready, j & 1);  // P_j is in tensor memory
            tc_fence_after();
            if (elect_one()) {
                const uint64_t v_desc = v_desc0 + s * (Q4_KV_BYTES >> 4);
                const int pv_steps = last ? pv_steps_last : Q4_BKV / 16;
#pragma unroll
                for (int k = 0; k < Q4_BKV / 16; ++k) {
                    if (k < pv_steps) umma_f16_ts(tmem_O, tmem_S + k * 8, v_desc + k * (2048 >> 4), IDESC_O, (j | k) != 0);
                }
                if (!last) {
                    const uint64_t k_desc = k_desc0 + s1 * (Q4_KV_BYTES >> 4);
                    const uint32_t idesc = (j + 2 == nkb) ? idesc_s_last : IDESC_S;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, q_desc + k * 2, k_desc + k * 2, idesc, k != 0);
                }
                umma_commit(s_full);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax warps: one thread per query row =====================
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t* kmask = nullptr;
        if (p.key_valid != nullptr || (p.T % Q4_BKV) != 0) {
            uint32_t* words = reinterpret_cast<uint32_t*>(smem + Q4_OFF_KMASK);
            const uint8_t* kvalid = p.key_valid ? p.key_valid + (size_t)img * p.T : nullptr;
            for (int w = warp - 1; w < nkb * (Q4_BKV / 32); w += 4) {
                const int key = w * 32 + lane;
                const bool ok = key < p.T && (kvalid == nullptr || kvalid[key] != 0);
                const uint32_t word = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) words[w] = word;
            }
            named_bar_sync(1, 128);
            kmask = words;
        }
        const bool warp_active = qt * ATT_BQ + quarter * 32 < p.T;
        SoftmaxState st{0.f, 0.f, 0.f};
        for (int j = 0; j < nkb; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            if (!warp_active) {
                tc_fence_before();
                mbar_arrive(p_ready);
                continue;
            }
            const bool last = j + 1 == nkb;
            const uint32_t* km = (p.key_valid != nullptr || last) ? kmask : nullptr;
            if (!last || p.n_last > 32) softmax_block_q4<2, POLY>(st, j, p, km, tmem_S + lane_off, tmem_O + lane_off, p_ready);
            else softmax_block_q4<1, POLY>(st, j, p, km, tmem_S + lane_off, tmem_O + lane_off, p_ready);
        }
        // ---- epilogue: O / l, bf16, staged in the (finished) Q tile, TMA store
        mbar_wait(s_full, nkb & 1);
        tc_fence_after();
        const float inv_l = st.l_run > 0.f ? 1.0f / st.l_run : 0.f;
        uint8_t* o_row = smem + Q4_OFF_Q + r * 128;
        if (warp_active) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
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
        }
        tc_fence_before();
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (threadIdx.x == 32) {
            tma_store_3d(&tmO, smem + Q4_OFF_Q, head * ATT_D, qt * ATT_BQ, img);
            tma_store_commit();
            tma_store_wait_read<0>();  // the staging tile has been read; the writes complete on their own
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<Q4_TMEM_COLS>(tmem_base);
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
    WG_REQUIRE(T <= ATT_MAX_T, "wg_attention_d64: T=%d exceeds the supported maximum %d", T, ATT_MAX_T);
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
    p.extra_key = 0;
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    // tuning knobs (defaults measured on B200): WG_ATTN_POLY = exponentials per 8 evaluated on the FMA pipe,
    // WG_ATTN_SPLIT = softmax threads per query row
    struct Knobs { int poly, split, smem_pad, persist, extra, q4, poly_set; };
    static const Knobs knobs = [] {  // read once (C++11 guarantees a thread-safe initialisation)
        Knobs k;
        const char* pad = getenv("WG_ATTN_SMEM_PAD");  // debug: extra dynamic smem (forces one CTA per SM when large)
        k.smem_pad = pad ? atoi(pad) : 0;
        const char* e = getenv("WG_ATTN_POLY");
        k.poly = e ? atoi(e) : ATT_POLY_DEFAULT;
        k.poly_set = e != nullptr;
        if (k.poly < 0 || k.poly > 4) k.poly = ATT_POLY_DEFAULT;
        e = getenv("WG_ATTN_SPLIT");
        k.split = e ? atoi(e) : ATT_SPLIT_DEFAULT;
        if (k.split != 1 && k.split != 2) k.split = ATT_SPLIT_DEFAULT;
        e = getenv("WG_ATTN_PERSIST");
        k.persist = (e == nullptr || atoi(e) != 0) ? 1 : 0;
        e = getenv("WG_ATTN_EXTRA_KEY");  // 0: the 1025th key gets its own (narrow) key block, as in round 1
        k.extra = (e == nullptr || atoi(e) != 0) ? 1 : 0;
        e = getenv("WG_ATTN_Q4");  // four CTAs per SM with 64-key blocks (attention_d64_q4_kernel)
        k.q4 = e ? atoi(e) : ATT_Q4_DEFAULT;
        return k;
    }();
    const int poly = knobs.poly, split = knobs.split, smem_pad = knobs.smem_pad, persist = knobs.persist;
    // The 1025th token of the CLIP sequence simply gets a ninth (mostly empty) query tile and a ninth, 16-key-wide key block:
    // idle softmax warps skip the arithmetic and the last S / PV MMAs are narrow.  Measured on B200 (B=64, 16 heads): T=1024
    // 0.43 ms, T=1025 this way 0.51 ms; a separate CUDA-core kernel for the last query row (it must stream K and V from HBM
    // once more) 0.53 ms; the leftover key folded into every row on CUDA cores instead of a ninth key block 0.55 ms; the last
    // row folded into the producer warps 0.65-0.83 ms.
    Prof prof("attention_d64", stream, 4.0 * B * heads * (double)T * T * ATT_D, 2.0 * 4.0 * B * (double)T * heads * ATT_D);
    if (knobs.q4 && split == 1) {
        CUtensorMap tmKV;
        uint64_t dims[3] = {3 * HD, (uint64_t)T, (uint64_t)B};
        uint64_t strides[2] = {3 * HD * 2, (uint64_t)T * 3 * HD * 2};
        uint32_t box[3] = {ATT_D, Q4_BKV, 1};
        WG_TRY(make_tensor_map(&tmKV, qkv, 2, 3, dims, strides, box));
        p.num_kv_blocks = (T + Q4_BKV - 1) / Q4_BKV;
        p.n_last = ((T - (p.num_kv_blocks - 1) * Q4_BKV) + 15) / 16 * 16;
        dim3 grid((T + ATT_BQ - 1) / ATT_BQ, heads, B);
#define WG_ATT_Q4LAUNCH(P)                                                                                                      \
    do {                                                                                                                        \
        WG_SMEM_OPT_IN(attention_d64_q4_kernel<P>, Q4_SMEM_BYTES);                                                              \
        attention_d64_q4_kernel<P><<<grid, Q4_THREADS, Q4_SMEM_BYTES, stream>>>(tmQKV, tmKV, tmO, p);                          \
    } while (0)
        switch (knobs.poly_set ? poly : ATT_Q4_POLY_DEFAULT) {
            case 1: WG_ATT_Q4LAUNCH(1); break;
            case 2: WG_ATT_Q4LAUNCH(2); break;
            case 3: WG_ATT_Q4LAUNCH(3); break;
            case 4: WG_ATT_Q4LAUNCH(4); break;
            default: WG_ATT_Q4LAUNCH(0); break;
        }
#undef WG_ATT_Q4LAUNCH
        WG_CHECK_CUDA(cudaGetLastError());
        return WG_OK;
    }
    // persistent kernel (two CTAs per SM walking the work list) from 4 key blocks per tile on; WG_ATTN_PERSIST=0 falls back to one
    // CTA per (tile, head, image)
    const long long n_items = (long long)B * heads * ((T + ATT_BQ - 1) / ATT_BQ);
    if (persist && split == 1 && p.num_kv_blocks >= 4 && n_items < (1ll << 30)) {
        if (knobs.extra && T % ATT_BKV == 1 && T / ATT_BKV >= 4) {  // T = 128 k + 1: k full key blocks + the last key on CUDA cores
            p.extra_key = 1;
            p.num_kv_blocks = T / ATT_BKV;
            p.n_last = ATT_BKV;
        }
        p.n_full_tiles = T / ATT_BQ;
        p.batch = B;
        p.n_items = (int)n_items;
        const int slots = 2 * device_sm_count();
        const int grid = p.n_items < slots ? p.n_items : slots;
#define WG_ATT_PLAUNCH(P)                                                                                                       \
    do {                                                                                                                        \
        WG_SMEM_OPT_IN(attention_d64_persist_kernel<P>, PQ_SMEM_BYTES);                                                         \
        attention_d64_persist_kernel<P><<<grid, 192, PQ_SMEM_BYTES, stream>>>(tmQKV, tmO, p);                                   \
    } while (0)
        switch (poly) {
            case 1: WG_ATT_PLAUNCH(1); break;
            case 2: WG_ATT_PLAUNCH(2); break;
            case 3: WG_ATT_PLAUNCH(3); break;
            case 4: WG_ATT_PLAUNCH(4); break;
            default: WG_ATT_PLAUNCH(0); break;
        }
#undef WG_ATT_PLAUNCH
        WG_CHECK_CUDA(cudaGetLastError());
        return WG_OK;
    }
    {
        dim3 grid((T + ATT_BQ - 1) / ATT_BQ, heads, B);
#define WG_ATT_LAUNCH(P, S)                                                                                                     \
    do {                                                                                                                        \
        WG_SMEM_OPT_IN((attention_d64_kernel<P, S>), ATT_SMEM_BYTES + smem_pad);                                                \
        attention_d64_kernel<P, S><<<grid, att_threads(S), ATT_SMEM_BYTES + smem_pad, stream>>>(tmQKV, tmO, p);                            \
    } while (0)
        if (split == 2) {
            switch (poly) {
                case 1: WG_ATT_LAUNCH(1, 2); break;
                case 2: WG_ATT_LAUNCH(2, 2); break;
                case 3: WG_ATT_LAUNCH(3, 2); break;
                case 4: WG_ATT_LAUNCH(4, 2); break;
                default: WG_ATT_LAUNCH(0, 2); break;
            }
        } else {
            switch (poly) {
                case 1: WG_ATT_LAUNCH(1, 1); break;
                case 2: WG_ATT_LAUNCH(2, 1); break;
                case 3: WG_ATT_LAUNCH(3, 1); break;
                case 4: WG_ATT_LAUNCH(4, 1); break;
                default: WG_ATT_LAUNCH(0, 1); break;
            }
        }
#undef WG_ATT_LAUNCH
        WG_CHECK_CUDA(cudaGetLastError());
    }
    return WG_OK;
}

#ifdef ATT_TRACE
extern "C" __attribute__((visibility("default"))) int wg_debug_attn_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, wg::g_att_trace, sizeof(long long) * 3 * 3 * 16 * 8);
}
#endif
