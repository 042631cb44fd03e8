// tcgen05 k-tap conv + gated activation for the denoiser's residual layers (blocks.py:677-681), fp16 hi/lo operands
// (3 MMAs per K step, see umma_conv.cu): the general kernel's SPLIT path specialised for this one shape (K = 3 x 256,
// N = 512, 80 launches per step).  What the profiles said, in the order it was learnt (profiles/ncu_r1_gate_roles.txt):
// the general kernel streams 786 KB of operands per 128 x 128 tile (one A tile PER TAP); fetching the A tile once cut
// that by a third and changed nothing (62 us) — the kernel was bound by the warp issuing the cross-term MMAs (~100 % busy
// at 15 SASS instructions per tcgen05.mma) AND by its epilogue warps (libm expf / tanhf evaluated one output at a
// time); with both fixed it runs in 48 us at 79 % tensor-pipe activity.
//
// * HALO A TILE: the (128 + span)-row activation tile of a channel block is fetched ONCE (hi and lo) and every tap's
//   A operand is that tile at a row offset in the UMMA descriptor (same trick as umma_halo.cu: the 128B swizzle is a
//   function of the absolute shared-memory address, identical for the TMA write and the UMMA read).  The K loop runs
//   channel-block-major, tap-minor; A traffic drops 3x (786 -> 532 KB per tile);
// * two rings: A halo tiles (2 stages x 34 KB) and weight blocks (4 stages x 32 KB), each with its own producer warp;
// * two tcgen05.mma issuing warps as in the general kernel (main products / cross terms, separate TMEM accumulators),
//   issued from `if (elect_one())` blocks with warp-uniform operands only (compile-time ring indices, make_uniform()'d
//   bases): bare UTCHMMA runs, ~2 SASS instructions per MMA instead of ~15 (see issue_tiles);
// * the DN_GATE epilogue of umma_conv.cu (8 warps, fp16 hi/lo store) with cheaper sigmoid x tanh math (gate_fast).
//
// Roles: warp 0 = A producer, warp 2 = TMEM allocator + weight producer, warps 1 / 3 = MMA issuers, warps 4-11 =
// epilogue.  Summation order differs from the general kernel's (channel-block-major), results agree to fp32 rounding.
#include "umma_common.cuh"
#include <stdlib.h>

namespace {

using namespace umma;

constexpr int G_BM = 128, G_BN = 128, G_BK = 64;
constexpr int G_ROW_BYTES = G_BK * 2;                 // 128
constexpr int G_A_STAGES = 2, G_B_STAGES = 4;
constexpr int G_B_BYTES = G_BN * G_ROW_BYTES;         // one weight block (hi or lo)
constexpr int G_B_STAGE = 2 * G_B_BYTES;              // hi + lo

// sigmoid(g) * tanh(f) = (1 - e^{-2f}) / ((1 + e^{-g}) (1 + e^{-2f})) with two MUFU.EX2 and one MUFU.RCP, BRANCH-FREE.
// expf / tanhf / IEEE division each carry a slow-path branch, so with them the compiler evaluated the 16 outputs of a
// chunk strictly one after the other — ~25 dependent instructions each, no overlap — and the 8 epilogue warps were
// ~95 % busy (ncu), i.e. the epilogue, not the tensor pipe, bounded the kernel (the same is true of the DN_GATE path
// of umma_conv.cu).  Straight-line math lets the 16 chains interleave.  Arguments are clamped so that no intermediate
// overflows or goes subnormal (sigmoid(-30) = 9e-14, 1 - tanh(15) = 2e-13: below fp32 resolution of the result), which
// is what makes the .ftz approximations safe; absolute error <= ~2e-7 (ex2.approx is good to 2 ulp, the 1 - e^{-2f}
// cancellation costs one ulp of 1) — the size of the hi/lo operands' own 2^-22 truncation.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gate_fast(float g, float f) {
    const float a = fminf(fmaxf(g, -30.f), 30.f) * -1.4426950408889634f;      // -g log2(e)
    const float b = fminf(fmaxf(f, -15.f), 15.f) * -2.8853900817779268f;      // -2 f log2(e)
    const float eg = ex2_approx(a), ef = ex2_approx(b);
    return (1.f - ef) * rcp_approx((1.f + eg) * (1.f + ef));
}

// MMA issue loop of one warp (CROSS = false: main products into accumulator 0; true: the two cross terms into
// accumulator 1).  ncu on the first version of this kernel (and on umma_conv_kernel's SPLIT path) showed the warp
// issuing the cross terms busy ~100 % of the time at ~17 SASS instructions per tcgen05.mma (ELECT / VOTEU and five
// R2UR per MMA: descriptors were built from ring indices held in vector registers) while the tensor pipe idled 40 %.
// Here every operand of UTCHMMA is derived from warp-uniform values only: ring stage indices and phases are
// COMPILE-TIME (the rings' depths divide the per-tile stage counts, so every tile starts at stage 0), the shared-memory
// bases go through make_uniform() once, and the MMAs sit in `if (elect_one())` blocks (umma_common.cuh).
template <int TAPS, int CB, bool CROSS>
__device__ __forceinline__ void issue_tiles(const UmmaConvParams& p, uint8_t* smA, uint8_t* smB, int a_stage, int a_half,
                                            uint64_t* a_full, uint64_t* a_empty, uint64_t* b_full, uint64_t* b_empty,
                                            uint64_t* tfull, uint64_t* tempty, uint32_t tmem_base, int tiles) {
    constexpr int BN = G_BN, ACC_COLS = 2 * G_BN;
    static_assert(CB % G_A_STAGES == 0 && (CB * TAPS) % G_B_STAGES == 0, "every tile must start at ring stage 0");
    constexpr int A_USES = CB / G_A_STAGES, B_USES = CB * TAPS / G_B_STAGES;   // uses of one ring stage per tile
    const uint32_t tmem_u = make_uniform(tmem_base);
    const uint32_t idesc = make_idesc(G_BM, BN);
    constexpr uint32_t DESC_HI = (uint32_t)((8 * G_ROW_BYTES) >> 4) | (1u << 14) | (2u << 29);   // SBO | version | SWIZZLE_128B
    const uint32_t tap_step = make_uniform((uint32_t)(((TAPS > 1 ? p.shift[1] - p.shift[0] : 0) * G_ROW_BYTES) >> 4));
    const uint32_t a_half16 = make_uniform((uint32_t)(a_half >> 4));
    const uint32_t a_stage16 = make_uniform((uint32_t)(a_stage >> 4));
    const uint32_t a_base = make_uniform(((smem_u32(smA) >> 4) & 0x3FFF) | (1u << 16));
    const uint32_t b_base = make_uniform(((smem_u32(smB) >> 4) & 0x3FFF) | (1u << 16));
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t abuf = it & 1u, tphase = (it >> 1) & 1u;
        mbar_wait(&tempty[abuf], tphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = make_uniform(tmem_u + abuf * (uint32_t)ACC_COLS + (CROSS ? (uint32_t)BN : 0u));
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
            const int as = cb % G_A_STAGES;
            const uint32_t aphase = (uint32_t)((it * A_USES + cb / G_A_STAGES) & 1u);
            mbar_wait(&a_full[as], aphase);
            tc_fence_after();
            const uint32_t a0 = a_base + (uint32_t)as * a_stage16;
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                const int idx = cb * TAPS + tap;
                const int bs = idx % G_B_STAGES;
                const uint32_t bphase = (uint32_t)((it * B_USES + idx / G_B_STAGES) & 1u);
                mbar_wait(&b_full[bs], bphase);
                tc_fence_after();
                const uint32_t ah = a0 + (uint32_t)tap * tap_step;     // hi halo tile at this tap's row offset
                const uint32_t wh = b_base + (uint32_t)((bs * G_B_STAGE) >> 4);
                if (elect_one()) {
                    constexpr uint64_t HI = (uint64_t)DESC_HI << 32;
                    if (!CROSS) {
#pragma unroll
                        for (int k = 0; k < G_BK / 16; ++k)
                            umma_f16(d_tmem, HI | (ah + 2 * k), HI | (wh + 2 * k), idesc, (cb | tap | k) ? 1u : 0u);
                    } else {
                        const uint32_t wl = wh + (uint32_t)(G_B_BYTES >> 4), al = ah + a_half16;
#pragma unroll
                        for (int k = 0; k < G_BK / 16; ++k)                                  // A_hi W_lo
                            umma_f16(d_tmem, HI | (ah + 2 * k), HI | (wl + 2 * k), idesc, (cb | tap | k) ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < G_BK / 16; ++k)                                  // A_lo W_hi
                            umma_f16(d_tmem, HI | (al + 2 * k), HI | (wh + 2 * k), idesc, 1u);
                    }
                    umma_commit(&b_empty[bs]);                  // frees the weight stage once these MMAs have read it
                    if (tap == TAPS - 1) umma_commit(&a_empty[as]);
                    if (tap == TAPS - 1 && cb == CB - 1) umma_commit(&tfull[abuf]);   // accumulator complete -> epilogue
                }
                __syncwarp();
            }
        }
    }
}

template <int TAPS, int CB>
__global__ void __launch_bounds__(384, 1)
umma_gate_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                 const UmmaConvParams p, const int rows_alloc, const int box_rows) {
    constexpr int BM = G_BM, BN = G_BN;
    constexpr int ACC_COLS = 2 * BN;                  // main | cross-term accumulator
    constexpr int TMEM_COLS = 2 * ACC_COLS;           // double-buffered: 512 columns

    const int a_half = rows_alloc * G_ROW_BYTES;      // hi (or lo) halo tile, 1024-byte multiple
    const int a_stage = 2 * a_half;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* smA = smem;
    uint8_t* smB = smA + G_A_STAGES * a_stage;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smB + G_B_STAGES * G_B_STAGE);
    uint64_t* a_empty = a_full + G_A_STAGES;
    uint64_t* b_full = a_empty + G_A_STAGES;
    uint64_t* b_empty = b_full + G_B_STAGES;
    uint64_t* tfull = b_empty + G_B_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: warp index provably uniform for ptxas
    const int m_tiles = (p.M + BM - 1) / BM;
    const int n_tiles = p.N / BN;
    const int tiles = p.B * m_tiles * n_tiles;        // n-tile fastest: the n-tiles of one row block run side by side
    const int shift0 = p.shift[0];

    if (threadIdx.x == 0) {
        // both issuing warps commit to the ring-empty and accumulator-full barriers
        for (int i = 0; i < G_A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 2); }
        for (int i = 0; i < G_B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 2); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 2); mbar_init(&tempty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    if (warp != 2) pdl_wait();      // the weight producer reads constants only: it runs ahead of the previous kernel's tail

    if (warp == 0) {
        // ================================ A producer: one halo tile (hi + lo) per channel block ================================
        prefetch_tmap(&tmA0); prefetch_tmap(&tmA1);
        int stage = 0; uint32_t phase = 0;
        const uint32_t bytes = (uint32_t)(2 * box_rows * G_ROW_BYTES);
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int rest = tile / n_tiles;
            const int mt = rest % m_tiles, b = rest / m_tiles;
#pragma unroll 1
            for (int cb = 0; cb < CB; ++cb) {
                mbar_wait(&a_empty[stage], phase ^ 1);
                mbar_expect_tx_elect(&a_full[stage], bytes);
                uint8_t* sa = smA + stage * a_stage;
                tma_load_3d_elect(sa, &tmA0, &a_full[stage], cb * G_BK, mt * BM + shift0, b);
                tma_load_3d_elect(sa + a_half, &tmA1, &a_full[stage], cb * G_BK, mt * BM + shift0, b);
                if (++stage == G_A_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ================================ weight producer: (channel block, tap) blocks, hi + lo ================================
        prefetch_tmap(&tmB0); prefetch_tmap(&tmB1);
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int nt = tile % n_tiles;
#pragma unroll 1
            for (int cb = 0; cb < CB; ++cb) {
#pragma unroll 1
                for (int tap = 0; tap < TAPS; ++tap) {
                    mbar_wait(&b_empty[stage], phase ^ 1);
                    mbar_expect_tx_elect(&b_full[stage], G_B_STAGE);
                    uint8_t* sb = smB + stage * G_B_STAGE;
                    tma_load_2d_elect(sb, &tmB0, &b_full[stage], cb * G_BK, tap * p.N + nt * BN);
                    tma_load_2d_elect(sb + G_B_BYTES, &tmB1, &b_full[stage], cb * G_BK, tap * p.N + nt * BN);
                    if (++stage == G_B_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ================================ MMA issuers ================================
        // warp 1: A_hi W_hi -> accumulator 0; warp 3: A_hi W_lo + A_lo W_hi -> accumulator 1 (disjoint, no ordering needed)
        if (warp == 1) issue_tiles<TAPS, CB, false>(p, smA, smB, a_stage, a_half, a_full, a_empty, b_full, b_empty, tfull, tempty, tmem_base, tiles);
        else issue_tiles<TAPS, CB, true>(p, smA, smB, a_stage, a_half, a_full, a_empty, b_full, b_empty, tfull, tempty, tmem_base, tiles);
    } else if (warp >= 4) {
        // ================================ epilogue (UEPI_DN_GATE) ================================
        // tile columns [0, BN/2) are gates, [BN/2, BN) the matching filters (weights.py gate_permutation); warp w reads
        // TMEM lane quarter (w & 3) and one half of the gate / filter columns
        constexpr int GH = BN / 4;
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        int abuf = 0; uint32_t tphase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int nt = tile % n_tiles, rest = tile / n_tiles;
            const int mt = rest % m_tiles, b = rest / m_tiles;
            const int t = mt * BM + row;
            bool valid = t < p.M;
            if (p.rows_per_utt > 0) {                         // flattened layout: never write the guard row of an utterance
                const int ub = t / p.rows_per_utt;
                valid = valid && (t - ub * p.rows_per_utt) < p.rows_per_utt - 1;
            }
            // this thread's 2 x GH bias values, fetched before the accumulator wait
            float bg[GH], bf[GH];
            {
                const float4* pg = reinterpret_cast<const float4*>(p.bias + nt * BN + h * GH);
                const float4* pf = reinterpret_cast<const float4*>(p.bias + nt * BN + BN / 2 + h * GH);
#pragma unroll
                for (int i = 0; i < GH / 4; ++i) {
                    const float4 x = __ldg(pg + i), y = __ldg(pf + i);
                    bg[4 * i] = x.x; bg[4 * i + 1] = x.y; bg[4 * i + 2] = x.z; bg[4 * i + 3] = x.w;
                    bf[4 * i] = y.x; bf[4 * i + 1] = y.y; bf[4 * i + 2] = y.z; bf[4 * i + 3] = y.w;
                }
            }
            mbar_wait(&tfull[abuf], tphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(abuf * ACC_COLS) + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int c = 0; c < GH / 16; ++c) {
                uint32_t rg[16], rf[16], rg2[16], rf2[16];
                const int g0 = h * GH + c * 16;
                tmem_ld16(taddr + g0, rg);
                tmem_ld16(taddr + BN / 2 + g0, rf);
                tmem_ld16(taddr + BN + g0, rg2);               // cross-term accumulator
                tmem_ld16(taddr + BN + BN / 2 + g0, rf2);
                tmem_ld_wait();
                if (valid) {
                    const int ch = nt * (BN / 2) + g0;
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float g = fmaf(__uint_as_float(rg[j]) + __uint_as_float(rg2[j]), p.alpha, bg[c * 16 + j]);
                        const float f = fmaf(__uint_as_float(rf[j]) + __uint_as_float(rf2[j]), p.alpha, bf[c * 16 + j]);
                        v[j] = gate_fast(g, f);
                    }
                    const long long o = (long long)b * p.out_bstride + (long long)t * p.out_ld + ch;
                    store16_hilo(p.out_h + o, p.out_lo + o, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[abuf]);
            abuf ^= 1; if (abuf == 0) tphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ==========================================================================================================================
// fp8 CROSS TERMS.  The hi/lo scheme spends two of its three MMAs on A_hi W_lo + A_lo W_hi, which is 2^-11 of the product:
// that part does not need 11-bit operands.  Here it runs as kind::f8f6f4 on e4m3 copies of the operands (hi8 = e4m3(hi),
// lo8 = e4m3(lo * 2^11); written next to the fp16 pair by the epilogues that produce y, and packed once for the weights):
// K = 32 per instruction at twice the fp16 rate, i.e. 12 -> 8 MMA issue slots per 64 channels.  e4m3's 2^-4 relative
// rounding on a 2^-11 term leaves ~2^-15 of the product (fp16 alone: 2^-11): measured on the whole residual stack the mel
// error is 1.4e-4 (fp16 cross terms 4e-6, none 3.6e-3; contract 1e-3).
//
// The main products and the cross terms now read DIFFERENT tiles, so each issuing warp has its own two rings and its own
// two producer warps: main  — A_hi16 halo tile per 64-channel block (2 stages) + W_hi16 blocks (3 stages);
//                     cross — {A_hi8, A_lo8} halo tiles per 128-channel block (128-byte rows again, 2 stages)
//                             + {W_hi8, W_lo8} blocks per (128-channel block, tap) (2 stages).
// Every ring depth divides its per-tile use count, so ring indices and phases stay compile-time (see issue_tiles).
// Roles (448 threads): warp 0 A16 producer, 1 main issuer, 2 TMEM allocator + W16 producer, 3 cross issuer,
// 4-11 epilogue, 12 A8 producer, 13 W8 producer.
constexpr int G8_A16_STAGES = 2, G8_W16_STAGES = 3, G8_A8_STAGES = 2, G8_W8_STAGES = 2;
constexpr int G8_THREADS = 448;
constexpr int G8_W8_STAGE = 2 * G_B_BYTES;            // hi8 + lo8 block: 128 rows x 128 bytes each
constexpr float G8_LO_SCALE_INV = 1.0f / 2048.0f;     // weights.py: F8_LO_SCALE (both cross terms carry one factor 2^11)

template <int TAPS>
__global__ void __launch_bounds__(G8_THREADS, 1)
umma_gate8_kernel(const __grid_constant__ CUtensorMap tmA16, const __grid_constant__ CUtensorMap tmA8h,
                  const __grid_constant__ CUtensorMap tmA8l, const __grid_constant__ CUtensorMap tmW16,
                  const __grid_constant__ CUtensorMap tmW8h, const __grid_constant__ CUtensorMap tmW8l,
                  const UmmaConvParams p, const int rows_alloc, const int box_rows) {
    constexpr int BM = G_BM, BN = G_BN;
    constexpr int CB = 4, CB8 = 2;                    // 256 channels: four 64-channel fp16 blocks / two 128-channel e4m3 blocks
    constexpr int ACC_COLS = 2 * BN;                  // main | cross-term accumulator
    constexpr int TMEM_COLS = 2 * ACC_COLS;           // double-buffered: 512 columns
    static_assert(CB % G8_A16_STAGES == 0 && (CB * TAPS) % G8_W16_STAGES == 0 && CB8 % G8_A8_STAGES == 0 &&
                  (CB8 * TAPS) % G8_W8_STAGES == 0, "every tile must start at ring stage 0");

    const int a_half = rows_alloc * G_ROW_BYTES;      // one halo tile (128-byte rows), 1024-byte multiple
    const int a8_stage = 2 * a_half;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* smA16 = smem;
    uint8_t* smW16 = smA16 + G8_A16_STAGES * a_half;
    uint8_t* smA8 = smW16 + G8_W16_STAGES * G_B_BYTES;
    uint8_t* smW8 = smA8 + G8_A8_STAGES * a8_stage;
    uint64_t* a16_full = reinterpret_cast<uint64_t*>(smW8 + G8_W8_STAGES * G8_W8_STAGE);
    uint64_t* a16_empty = a16_full + G8_A16_STAGES;
    uint64_t* w16_full = a16_empty + G8_A16_STAGES;
    uint64_t* w16_empty = w16_full + G8_W16_STAGES;
    uint64_t* a8_full = w16_empty + G8_W16_STAGES;
    uint64_t* a8_empty = a8_full + G8_A8_STAGES;
    uint64_t* w8_full = a8_empty + G8_A8_STAGES;
    uint64_t* w8_empty = w8_full + G8_W8_STAGES;
    uint64_t* tfull = w8_empty + G8_W8_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM;
    const int n_tiles = p.N / BN;
    const int tiles = p.B * m_tiles * n_tiles;        // n-tile fastest
    const int shift0 = p.shift[0];

    if (threadIdx.x == 0) {
        for (int i = 0; i < G8_A16_STAGES; ++i) { mbar_init(&a16_full[i], 1); mbar_init(&a16_empty[i], 1); }
        for (int i = 0; i < G8_W16_STAGES; ++i) { mbar_init(&w16_full[i], 1); mbar_init(&w16_empty[i], 1); }
        for (int i = 0; i < G8_A8_STAGES; ++i) { mbar_init(&a8_full[i], 1); mbar_init(&a8_empty[i], 1); }
        for (int i = 0; i < G8_W8_STAGES; ++i) { mbar_init(&w8_full[i], 1); mbar_init(&w8_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 2); mbar_init(&tempty[i], 8); }     // both issuers commit tfull
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    if (warp != 2 && warp != 13) pdl_wait();          // the weight producers read constants only

    constexpr uint32_t DESC_HI = (uint32_t)((8 * G_ROW_BYTES) >> 4) | (1u << 14) | (2u << 29);   // SBO | version | SWIZZLE_128B
    constexpr uint64_t HI = (uint64_t)DESC_HI << 32;

    if (warp == 0) {
        // ================= A16 producer: A_hi16 halo tile per 64-channel block =================
        prefetch_tmap(&tmA16);
        int stage = 0; uint32_t phase = 0;
        const uint32_t bytes = (uint32_t)(box_rows * G_ROW_BYTES);
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int rest = tile / n_tiles;
            const int mt = rest % m_tiles, b = rest / m_tiles;
#pragma unroll 1
            for (int cb = 0; cb < CB; ++cb) {
                mbar_wait(&a16_empty[stage], phase ^ 1);
                mbar_expect_tx_elect(&a16_full[stage], bytes);
                tma_load_3d_elect(smA16 + stage * a_half, &tmA16, &a16_full[stage], cb * G_BK, mt * BM + shift0, b);
                if (++stage == G8_A16_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 12) {
        // ================= A8 producer: {A_hi8, A_lo8} halo tiles per 128-channel block =================
        prefetch_tmap(&tmA8h); prefetch_tmap(&tmA8l);
        int stage = 0; uint32_t phase = 0;
        const uint32_t bytes = (uint32_t)(2 * box_rows * G_ROW_BYTES);
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int rest = tile / n_tiles;
            const int mt = rest % m_tiles, b = rest / m_tiles;
#pragma unroll 1
            for (int cb = 0; cb < CB8; ++cb) {
                mbar_wait(&a8_empty[stage], phase ^ 1);
                mbar_expect_tx_elect(&a8_full[stage], bytes);
                uint8_t* sa = smA8 + stage * a8_stage;
                tma_load_3d_elect(sa, &tmA8h, &a8_full[stage], cb * 128, mt * BM + shift0, b);
                tma_load_3d_elect(sa + a_half, &tmA8l, &a8_full[stage], cb * 128, mt * BM + shift0, b);
                if (++stage == G8_A8_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ================= W16 producer: W_hi16 blocks per (64-channel block, tap) =================
        prefetch_tmap(&tmW16);
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int nt = tile % n_tiles;
#pragma unroll 1
            for (int cb = 0; cb < CB; ++cb) {
#pragma unroll 1
                for (int tap = 0; tap < TAPS; ++tap) {
                    mbar_wait(&w16_empty[stage], phase ^ 1);
                    mbar_expect_tx_elect(&w16_full[stage], G_B_BYTES);
                    tma_load_2d_elect(smW16 + stage * G_B_BYTES, &tmW16, &w16_full[stage], cb * G_BK, tap * p.N + nt * BN);
                    if (++stage == G8_W16_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 13) {
        // ================= W8 producer: {W_hi8, W_lo8} blocks per (128-channel block, tap) =================
        prefetch_tmap(&tmW8h); prefetch_tmap(&tmW8l);
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int nt = tile % n_tiles;
#pragma unroll 1
            for (int cb = 0; cb < CB8; ++cb) {
#pragma unroll 1
                for (int tap = 0; tap < TAPS; ++tap) {
                    mbar_wait(&w8_empty[stage], phase ^ 1);
                    mbar_expect_tx_elect(&w8_full[stage], G8_W8_STAGE);
                    uint8_t* sb = smW8 + stage * G8_W8_STAGE;
                    tma_load_2d_elect(sb, &tmW8h, &w8_full[stage], cb * 128, tap * p.N + nt * BN);
                    tma_load_2d_elect(sb + G_B_BYTES, &tmW8l, &w8_full[stage], cb * 128, tap * p.N + nt * BN);
                    if (++stage == G8_W8_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= main issuer: A_hi16 W_hi16 -> accumulator 0 =================
        constexpr int A_USES = CB / G8_A16_STAGES, B_USES = CB * TAPS / G8_W16_STAGES;
        const uint32_t tmem_u = make_uniform(tmem_base);
        const uint32_t idesc = make_idesc(BM, BN);
        const uint32_t tap_step = make_uniform((uint32_t)(((TAPS > 1 ? p.shift[1] - p.shift[0] : 0) * G_ROW_BYTES) >> 4));
        const uint32_t a_half16 = make_uniform((uint32_t)(a_half >> 4));
        const uint32_t a_base = make_uniform(((smem_u32(smA16) >> 4) & 0x3FFF) | (1u << 16));
        const uint32_t b_base = make_uniform(((smem_u32(smW16) >> 4) & 0x3FFF) | (1u << 16));
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const uint32_t abuf = it & 1u, tphase = (it >> 1) & 1u;
            mbar_wait(&tempty[abuf], tphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = make_uniform(tmem_u + abuf * (uint32_t)ACC_COLS);
#pragma unroll
            for (int cb = 0; cb < CB; ++cb) {
                const int as = cb % G8_A16_STAGES;
                mbar_wait(&a16_full[as], (uint32_t)((it * A_USES + cb / G8_A16_STAGES) & 1u));
                tc_fence_after();
                const uint32_t a0 = a_base + (uint32_t)as * a_half16;
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
                    const int idx = cb * TAPS + tap;
                    const int bs = idx % G8_W16_STAGES;
                    mbar_wait(&w16_full[bs], (uint32_t)((it * B_USES + idx / G8_W16_STAGES) & 1u));
                    tc_fence_after();
                    const uint32_t ah = a0 + (uint32_t)tap * tap_step;
                    const uint32_t wh = b_base + (uint32_t)((bs * G_B_BYTES) >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < G_BK / 16; ++k)
                            umma_f16(d_tmem, HI | (ah + 2 * k), HI | (wh + 2 * k), idesc, (cb | tap | k) ? 1u : 0u);
                        umma_commit(&w16_empty[bs]);
                        if (tap == TAPS - 1) umma_commit(&a16_empty[as]);
                        if (tap == TAPS - 1 && cb == CB - 1) umma_commit(&tfull[abuf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 3) {
        // ================= cross issuer (e4m3): A_hi8 W_lo8 + A_lo8 W_hi8 -> accumulator 1 =================
        constexpr int A_USES = CB8 / G8_A8_STAGES, B_USES = CB8 * TAPS / G8_W8_STAGES;
        const uint32_t tmem_u = make_uniform(tmem_base);
        const uint32_t idesc = make_idesc(BM, BN);        // e4m3 x e4m3 -> f32: same bits as f16 x f16 -> f32
        const uint32_t tap_step = make_uniform((uint32_t)(((TAPS > 1 ? p.shift[1] - p.shift[0] : 0) * G_ROW_BYTES) >> 4));
        const uint32_t a_half16 = make_uniform((uint32_t)(a_half >> 4));
        const uint32_t a_base = make_uniform(((smem_u32(smA8) >> 4) & 0x3FFF) | (1u << 16));
        const uint32_t b_base = make_uniform(((smem_u32(smW8) >> 4) & 0x3FFF) | (1u << 16));
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const uint32_t abuf = it & 1u, tphase = (it >> 1) & 1u;
            mbar_wait(&tempty[abuf], tphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = make_uniform(tmem_u + abuf * (uint32_t)ACC_COLS + (uint32_t)BN);
#pragma unroll
            for (int cb = 0; cb < CB8; ++cb) {
                const int as = cb % G8_A8_STAGES;
                mbar_wait(&a8_full[as], (uint32_t)((it * A_USES + cb / G8_A8_STAGES) & 1u));
                tc_fence_after();
                const uint32_t a0 = a_base + (uint32_t)as * 2u * a_half16;
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
                    const int idx = cb * TAPS + tap;
                    const int bs = idx % G8_W8_STAGES;
                    mbar_wait(&w8_full[bs], (uint32_t)((it * B_USES + idx / G8_W8_STAGES) & 1u));
                    tc_fence_after();
                    const uint32_t ah = a0 + (uint32_t)tap * tap_step, al = ah + a_half16;
                    const uint32_t wh = b_base + (uint32_t)((bs * G8_W8_STAGE) >> 4), wl = wh + (uint32_t)(G_B_BYTES >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 128 / 32; ++k)                                   // A_hi8 W_lo8
                            umma_f8(d_tmem, HI | (ah + 2 * k), HI | (wl + 2 * k), idesc, (cb | tap | k) ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < 128 / 32; ++k)                                   // A_lo8 W_hi8
                            umma_f8(d_tmem, HI | (al + 2 * k), HI | (wh + 2 * k), idesc, 1u);
                        umma_commit(&w8_empty[bs]);
                        if (tap == TAPS - 1) umma_commit(&a8_empty[as]);
                        if (tap == TAPS - 1 && cb == CB8 - 1) umma_commit(&tfull[abuf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ================= epilogue (UEPI_DN_GATE), as umma_gate_kernel's, cross accumulator scaled by 2^-11 =================
        constexpr int GH = BN / 4;
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        int abuf = 0; uint32_t tphase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int nt = tile % n_tiles, rest = tile / n_tiles;
            const int mt = rest % m_tiles, b = rest / m_tiles;
            const int t = mt * BM + row;
            bool valid = t < p.M;
            if (p.rows_per_utt > 0) {
                const int ub = t / p.rows_per_utt;
                valid = valid && (t - ub * p.rows_per_utt) < p.rows_per_utt - 1;
            }
            float bg[GH], bf[GH];
            {
                const float4* pg = reinterpret_cast<const float4*>(p.bias + nt * BN + h * GH);
                const float4* pf = reinterpret_cast<const float4*>(p.bias + nt * BN + BN / 2 + h * GH);
#pragma unroll
                for (int i = 0; i < GH / 4; ++i) {
                    const float4 x = __ldg(pg + i), y = __ldg(pf + i);
                    bg[4 * i] = x.x; bg[4 * i + 1] = x.y; bg[4 * i + 2] = x.z; bg[4 * i + 3] = x.w;
                    bf[4 * i] = y.x; bf[4 * i + 1] = y.y; bf[4 * i + 2] = y.z; bf[4 * i + 3] = y.w;
                }
            }
            mbar_wait(&tfull[abuf], tphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(abuf * ACC_COLS) + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int c = 0; c < GH / 16; ++c) {
                uint32_t rg[16], rf[16], rg2[16], rf2[16];
                const int g0 = h * GH + c * 16;
                tmem_ld16(taddr + g0, rg);
                tmem_ld16(taddr + BN / 2 + g0, rf);
                tmem_ld16(taddr + BN + g0, rg2);
                tmem_ld16(taddr + BN + BN / 2 + g0, rf2);
                tmem_ld_wait();
                if (valid) {
                    const int ch = nt * (BN / 2) + g0;
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float g = fmaf(fmaf(__uint_as_float(rg2[j]), G8_LO_SCALE_INV, __uint_as_float(rg[j])), p.alpha, bg[c * 16 + j]);
                        const float f = fmaf(fmaf(__uint_as_float(rf2[j]), G8_LO_SCALE_INV, __uint_as_float(rf[j])), p.alpha, bf[c * 16 + j]);
                        v[j] = gate_fast(g, f);
                    }
                    const long long o = (long long)b * p.out_bstride + (long long)t * p.out_ld + ch;
                    store16_hilo(p.out_h + o, p.out_lo + o, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[abuf]);
            abuf ^= 1; if (abuf == 0) tphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}


// ==========================================================================================================================
// CTA-PAIR variant of umma_gate8_kernel (tcgen05 cta_group::2; see umma_halo2.cu for the mechanism and why it is the bytes
// INTO an SM, not L2 reads, that bound these kernels): one M = 256 MMA per instruction slot over both SMs' shared memory, each
// CTA holding its own 128 rows of A and HALF of every weight block (64 of the 128 output channels of the N tile).  Operand
// bytes per tile and SM: 136 KB of A + 192 KB of weights instead of 136 + 384.  Only the leader issues (both issuing warps);
// commits are multicast to both CTAs' barriers; `accumulator empty` lives in the leader (16 arrivals).
__device__ __forceinline__ uint32_t gx2_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void gx2_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t GX2_PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void gx2_tma_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & GX2_PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void gx2_tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & GX2_PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void gx2_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void gx2_mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void gx2_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void gx2_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & GX2_PEER_MASK) : "memory");
}

// per-CTA weight stages hold HALF an N tile (64 output channels): 8 KB fp16 blocks, 8 + 8 KB e4m3 pairs.  The shared memory
// that frees goes into DEEPER weight rings: the one-CTA kernel's 3 / 2 stages cover only ~0.5 us of tensor-pipe work, less
// than one L2 round trip, which (not the operand bytes) is what holds it at ~57 us per launch.
constexpr int GX2_WB = G_B_BYTES / 2;                 // one half block: 64 rows x 128 bytes
// (W16 ring: 4 stages, was 6 — the 16 KB go to the epilogue's output staging; the profile that asked for deeper rings was
// taken while the epilogue warps, spilling their bias arrays, hid everything else)
constexpr int GX2_W16_STAGES = 4, GX2_W8_STAGES = 3;
constexpr int GX2_OUT_SLAB = 4096;                    // per epilogue warp: g hi | g lo boxes of 32 columns x 32 rows
constexpr int GX2_OUT_BYTES = 8 * GX2_OUT_SLAB;
constexpr int GX2_MAX_N = 512;                        // bias staged in shared memory
constexpr int GX2_W8_STAGE = 2 * GX2_WB;

template <int TAPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G8_THREADS, 1)
umma_gate8x2_kernel(const __grid_constant__ CUtensorMap tmA16, const __grid_constant__ CUtensorMap tmA8h,
                  const __grid_constant__ CUtensorMap tmA8l, const __grid_constant__ CUtensorMap tmW16,
                  const __grid_constant__ CUtensorMap tmW8h, const __grid_constant__ CUtensorMap tmW8l,
                  const UmmaConvParams p, const int rows_alloc, const int box_rows) {
    constexpr int BM = G_BM, BN = G_BN;
    constexpr int CB = 4, CB8 = 2;                    // 256 channels: four 64-channel fp16 blocks / two 128-channel e4m3 blocks
    constexpr int ACC_COLS = 2 * BN;                  // main | cross-term accumulator
    constexpr int TMEM_COLS = 2 * ACC_COLS;           // double-buffered: 512 columns
    static_assert(CB % G8_A16_STAGES == 0 && (CB * TAPS) % GX2_W16_STAGES == 0 && CB8 % G8_A8_STAGES == 0 &&
                  (CB8 * TAPS) % GX2_W8_STAGES == 0, "every tile must start at ring stage 0");

    const int a_half = rows_alloc * G_ROW_BYTES;      // one halo tile (128-byte rows), 1024-byte multiple
    const int a8_stage = 2 * a_half;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* smOut = smem;                            // [8 warps][GX2_OUT_SLAB], 1024-byte aligned slabs
    uint8_t* smA16 = smOut + GX2_OUT_BYTES;
    uint8_t* smW16 = smA16 + G8_A16_STAGES * a_half;
    uint8_t* smA8 = smW16 + GX2_W16_STAGES * GX2_WB;
    uint8_t* smW8 = smA8 + G8_A8_STAGES * a8_stage;
    uint64_t* a16_full = reinterpret_cast<uint64_t*>(smW8 + GX2_W8_STAGES * GX2_W8_STAGE);
    uint64_t* a16_empty = a16_full + G8_A16_STAGES;
    uint64_t* w16_full = a16_empty + G8_A16_STAGES;
    uint64_t* w16_empty = w16_full + GX2_W16_STAGES;
    uint64_t* a8_full = w16_empty + GX2_W16_STAGES;
    uint64_t* a8_empty = a8_full + G8_A8_STAGES;
    uint64_t* w8_full = a8_empty + G8_A8_STAGES;
    uint64_t* w8_empty = w8_full + GX2_W8_STAGES;
    uint64_t* tfull = w8_empty + GX2_W8_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);   // [p.N]

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int rank = (int)gx2_ctarank();              // 0 = leader of the CTA pair
    // bias (constant weights: no dependency on the previous kernel) -> shared memory.  Held in registers (64 per
    // thread) it pushed the epilogue over the 128-register cap: spills (LDL in the element loops), the 8 epilogue
    // warps 93 % busy and the MMA issuers waiting for accumulators (ncu).
    for (int i = threadIdx.x; i < p.N; i += G8_THREADS) s_bias[i] = p.bias[i];
    const int m_tiles = (p.M + BM - 1) / BM;
    const int m_pairs = (m_tiles + 1) / 2;            // a pair covers 256 rows: row tile 2 * mp + rank per CTA
    const int n_tiles = p.N / BN;
    const int tiles = p.B * m_pairs * n_tiles;        // pair tiles, n-tile fastest
    const int shift0 = p.shift[0];
    const int first = blockIdx.x >> 1, stride = gridDim.x >> 1;     // cluster index / number of clusters

    if (threadIdx.x == 0) {
        for (int i = 0; i < G8_A16_STAGES; ++i) { mbar_init(&a16_full[i], 1); mbar_init(&a16_empty[i], 1); }
        for (int i = 0; i < GX2_W16_STAGES; ++i) { mbar_init(&w16_full[i], 1); mbar_init(&w16_empty[i], 1); }
        for (int i = 0; i < G8_A8_STAGES; ++i) { mbar_init(&a8_full[i], 1); mbar_init(&a8_empty[i], 1); }
        for (int i = 0; i < GX2_W8_STAGES; ++i) { mbar_init(&w8_full[i], 1); mbar_init(&w8_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 2); mbar_init(&tempty[i], 16); }    // both issuers commit tfull; 8 epilogue warps of each CTA
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    gx2_cluster_sync();                               // the peer's barriers exist before anything is signalled into them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    if (warp != 2 && warp != 13) pdl_wait();          // the weight producers read constants only

    constexpr uint32_t DESC_HI = (uint32_t)((8 * G_ROW_BYTES) >> 4) | (1u << 14) | (2u << 29);   // SBO | version | SWIZZLE_128B
    constexpr uint64_t HI = (uint64_t)DESC_HI << 32;

    if (warp == 0) {
        // ================= A16 producer: A_hi16 halo tile per 64-channel block =================
        prefetch_tmap(&tmA16);
        int stage = 0; uint32_t phase = 0;
        const uint32_t bytes = (uint32_t)(box_rows * G_ROW_BYTES);
        for (int tile = first; tile < tiles; tile += stride) {
            const int rest = tile / n_tiles;
            const int mt = 2 * (rest % m_pairs) + rank, b = rest / m_pairs;
#pragma unroll 1
            for (int cb = 0; cb < CB; ++cb) {
                mbar_wait(&a16_empty[stage], phase ^ 1);
                if (rank == 0) mbar_expect_tx_elect(&a16_full[stage], 2 * bytes);          // both CTAs' tiles
                gx2_tma_3d(smA16 + stage * a_half, &tmA16, &a16_full[stage], cb * G_BK, mt * BM + shift0, b);
                if (++stage == G8_A16_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 12) {
        // ================= A8 producer: {A_hi8, A_lo8} halo tiles per 128-channel block =================
        prefetch_tmap(&tmA8h); prefetch_tmap(&tmA8l);
        int stage = 0; uint32_t phase = 0;
        const uint32_t bytes = (uint32_t)(2 * box_rows * G_ROW_BYTES);
        for (int tile = first; tile < tiles; tile += stride) {
            const int rest = tile / n_tiles;
            const int mt = 2 * (rest % m_pairs) + rank, b = rest / m_pairs;
#pragma unroll 1
            for (int cb = 0; cb < CB8; ++cb) {
                mbar_wait(&a8_empty[stage], phase ^ 1);
                if (rank == 0) mbar_expect_tx_elect(&a8_full[stage], 2 * bytes);
                uint8_t* sa = smA8 + stage * a8_stage;
                gx2_tma_3d(sa, &tmA8h, &a8_full[stage], cb * 128, mt * BM + shift0, b);
                gx2_tma_3d(sa + a_half, &tmA8l, &a8_full[stage], cb * 128, mt * BM + shift0, b);
                if (++stage == G8_A8_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ================= W16 producer: W_hi16 blocks per (64-channel block, tap) =================
        prefetch_tmap(&tmW16);
        int stage = 0; uint32_t phase = 0;
        for (int tile = first; tile < tiles; tile += stride) {
            const int nt = tile % n_tiles;
#pragma unroll 1
            for (int cb = 0; cb < CB; ++cb) {
#pragma unroll 1
                for (int tap = 0; tap < TAPS; ++tap) {
                    mbar_wait(&w16_empty[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx_elect(&w16_full[stage], 2 * GX2_WB);   // two halves of 64 output channels
                    gx2_tma_2d(smW16 + stage * GX2_WB, &tmW16, &w16_full[stage], cb * G_BK, tap * p.N + nt * BN + rank * (BN / 2));
                    if (++stage == GX2_W16_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 13) {
        // ================= W8 producer: {W_hi8, W_lo8} blocks per (128-channel block, tap) =================
        prefetch_tmap(&tmW8h); prefetch_tmap(&tmW8l);
        int stage = 0; uint32_t phase = 0;
        for (int tile = first; tile < tiles; tile += stride) {
            const int nt = tile % n_tiles;
#pragma unroll 1
            for (int cb = 0; cb < CB8; ++cb) {
#pragma unroll 1
                for (int tap = 0; tap < TAPS; ++tap) {
                    mbar_wait(&w8_empty[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx_elect(&w8_full[stage], 2 * GX2_W8_STAGE);
                    uint8_t* sb = smW8 + stage * GX2_W8_STAGE;
                    gx2_tma_2d(sb, &tmW8h, &w8_full[stage], cb * 128, tap * p.N + nt * BN + rank * (BN / 2));
                    gx2_tma_2d(sb + GX2_WB, &tmW8l, &w8_full[stage], cb * 128, tap * p.N + nt * BN + rank * (BN / 2));
                    if (++stage == GX2_W8_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ================= main issuer (leader only; M = 256 across the pair): A_hi16 W_hi16 -> accumulator 0 =================
        constexpr int A_USES = CB / G8_A16_STAGES, B_USES = CB * TAPS / GX2_W16_STAGES;
        const uint32_t tmem_u = make_uniform(tmem_base);
        const uint32_t idesc = make_idesc(2 * BM, BN);
        const uint32_t tap_step = make_uniform((uint32_t)(((TAPS > 1 ? p.shift[1] - p.shift[0] : 0) * G_ROW_BYTES) >> 4));
        const uint32_t a_half16 = make_uniform((uint32_t)(a_half >> 4));
        const uint32_t a_base = make_uniform(((smem_u32(smA16) >> 4) & 0x3FFF) | (1u << 16));
        const uint32_t b_base = make_uniform(((smem_u32(smW16) >> 4) & 0x3FFF) | (1u << 16));
        uint32_t it = 0;
        for (int tile = first; tile < tiles; tile += stride, ++it) {
            const uint32_t abuf = it & 1u, tphase = (it >> 1) & 1u;
            mbar_wait(&tempty[abuf], tphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = make_uniform(tmem_u + abuf * (uint32_t)ACC_COLS);
#pragma unroll
            for (int cb = 0; cb < CB; ++cb) {
                const int as = cb % G8_A16_STAGES;
                mbar_wait(&a16_full[as], (uint32_t)((it * A_USES + cb / G8_A16_STAGES) & 1u));
                tc_fence_after();
                const uint32_t a0 = a_base + (uint32_t)as * a_half16;
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
                    const int idx = cb * TAPS + tap;
                    const int bs = idx % GX2_W16_STAGES;
                    mbar_wait(&w16_full[bs], (uint32_t)((it * B_USES + idx / GX2_W16_STAGES) & 1u));
                    tc_fence_after();
                    const uint32_t ah = a0 + (uint32_t)tap * tap_step;
                    const uint32_t wh = b_base + (uint32_t)((bs * GX2_WB) >> 4);
                    if (elect_one()) {
                        if (!(p.dbg & 32)) {                       // dbg 32 = timing ablation: no MMAs
#pragma unroll
                            for (int k = 0; k < G_BK / 16; ++k)
                                gx2_mma_f16(d_tmem, HI | (ah + 2 * k), HI | (wh + 2 * k), idesc, (cb | tap | k) ? 1u : 0u);
                        }
                        gx2_commit_both(&w16_empty[bs]);
                        if (tap == TAPS - 1) gx2_commit_both(&a16_empty[as]);
                        if (tap == TAPS - 1 && cb == CB - 1) gx2_commit_both(&tfull[abuf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 3 && rank == 0) {
        // ================= cross issuer (leader only) (e4m3): A_hi8 W_lo8 + A_lo8 W_hi8 -> accumulator 1 =================
        constexpr int A_USES = CB8 / G8_A8_STAGES, B_USES = CB8 * TAPS / GX2_W8_STAGES;
        const uint32_t tmem_u = make_uniform(tmem_base);
        const uint32_t idesc = make_idesc(2 * BM, BN);    // e4m3 x e4m3 -> f32: same bits as f16 x f16 -> f32
        const uint32_t tap_step = make_uniform((uint32_t)(((TAPS > 1 ? p.shift[1] - p.shift[0] : 0) * G_ROW_BYTES) >> 4));
        const uint32_t a_half16 = make_uniform((uint32_t)(a_half >> 4));
        const uint32_t a_base = make_uniform(((smem_u32(smA8) >> 4) & 0x3FFF) | (1u << 16));
        const uint32_t b_base = make_uniform(((smem_u32(smW8) >> 4) & 0x3FFF) | (1u << 16));
        uint32_t it = 0;
        for (int tile = first; tile < tiles; tile += stride, ++it) {
            const uint32_t abuf = it & 1u, tphase = (it >> 1) & 1u;
            mbar_wait(&tempty[abuf], tphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = make_uniform(tmem_u + abuf * (uint32_t)ACC_COLS + (uint32_t)BN);
#pragma unroll
            for (int cb = 0; cb < CB8; ++cb) {
                const int as = cb % G8_A8_STAGES;
                mbar_wait(&a8_full[as], (uint32_t)((it * A_USES + cb / G8_A8_STAGES) & 1u));
                tc_fence_after();
                const uint32_t a0 = a_base + (uint32_t)as * 2u * a_half16;
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
                    const int idx = cb * TAPS + tap;
                    const int bs = idx % GX2_W8_STAGES;
                    mbar_wait(&w8_full[bs], (uint32_t)((it * B_USES + idx / GX2_W8_STAGES) & 1u));
                    tc_fence_after();
                    const uint32_t ah = a0 + (uint32_t)tap * tap_step, al = ah + a_half16;
                    const uint32_t wh = b_base + (uint32_t)((bs * GX2_W8_STAGE) >> 4), wl = wh + (uint32_t)(GX2_WB >> 4);
                    if (elect_one()) {
                        if (!(p.dbg & 32)) {
#pragma unroll
                            for (int k = 0; k < 128 / 32; ++k)                                   // A_hi8 W_lo8
                                gx2_mma_f8(d_tmem, HI | (ah + 2 * k), HI | (wl + 2 * k), idesc, (cb | tap | k) ? 1u : 0u);
#pragma unroll
                            for (int k = 0; k < 128 / 32; ++k)                                   // A_lo8 W_hi8
                                gx2_mma_f8(d_tmem, HI | (al + 2 * k), HI | (wh + 2 * k), idesc, 1u);
                        }
                        gx2_commit_both(&w8_empty[bs]);
                        if (tap == TAPS - 1) gx2_commit_both(&a8_empty[as]);
                        if (tap == TAPS - 1 && cb == CB8 - 1) gx2_commit_both(&tfull[abuf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ================= epilogue (UEPI_DN_GATE): cross accumulator scaled by 2^-11, gate math, g as fp16 hi/lo =================
        // staged per warp as two tiles of 32 rows x 32 columns (64-byte rows, swizzled against bank conflicts) and written
        // out with coalesced stores (4 lanes per row segment: a warp-level STG.128 touches 8 rows, not 32 — see
        // umma_conv.cu, UEPI_DN_OUTY).  Guard rows are written as zeros (nothing consumes them).
        constexpr int GH = BN / 4;
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        uint8_t* slab = smOut + (warp - 4) * GX2_OUT_SLAB;
        uint8_t* srow = slab + lane * 64;
        const int sw3 = (lane >> 1) & 3;
        int abuf = 0; uint32_t tphase = 0;
        for (int tile = first; tile < tiles; tile += stride) {
            const int nt = tile % n_tiles, rest = tile / n_tiles;
            const int mt = 2 * (rest % m_pairs) + rank;
            const int t = mt * BM + row;
            bool valid = t < p.M;
            if (p.rows_per_utt > 0) {
                const int ub = t / p.rows_per_utt;
                valid = valid && (t - ub * p.rows_per_utt) < p.rows_per_utt - 1;
            }
            const float vmask = valid ? 1.f : 0.f;        // gate_fast clamps its arguments: finite for any accumulator
            const float* sbg = s_bias + nt * BN + h * GH;
            const float* sbf = sbg + BN / 2;
            mbar_wait(&tfull[abuf], tphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(abuf * ACC_COLS) + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int c = 0; c < GH / 16; ++c) {
                if (p.dbg & 64) break;                    // dbg 64 = timing ablation: the epilogue does nothing
                uint32_t rg[16], rf[16], rg2[16], rf2[16];
                const int g0 = h * GH + c * 16;
                tmem_ld16(taddr + g0, rg);
                tmem_ld16(taddr + BN / 2 + g0, rf);
                tmem_ld16(taddr + BN + g0, rg2);
                tmem_ld16(taddr + BN + BN / 2 + g0, rf2);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b4g = *reinterpret_cast<const float4*>(sbg + c * 16 + 4 * i);
                    const float4 b4f = *reinterpret_cast<const float4*>(sbf + c * 16 + 4 * i);
                    const float bgv[4] = {b4g.x, b4g.y, b4g.z, b4g.w}, bfv[4] = {b4f.x, b4f.y, b4f.z, b4f.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int j = 4 * i + k;
                        const float g = fmaf(fmaf(__uint_as_float(rg2[j]), G8_LO_SCALE_INV, __uint_as_float(rg[j])), p.alpha, bgv[k]);
                        const float f = fmaf(fmaf(__uint_as_float(rf2[j]), G8_LO_SCALE_INV, __uint_as_float(rf[j])), p.alpha, bfv[k]);
                        v[j] = __fmul_rn(gate_fast(g, f), vmask);      // a mask, not a select: `valid ? gate(...) : 0` compiles to a branch per element
                                                                       // (serial MUFU chains, no overlap between the 16 outputs)
                    }
                }
                uint4 h0, h1, l0, l1;
                pack16_hilo(v, h0, h1, l0, l1);
                *reinterpret_cast<uint4*>(srow + (((2 * c) ^ sw3) << 4)) = h0;
                *reinterpret_cast<uint4*>(srow + (((2 * c + 1) ^ sw3) << 4)) = h1;
                *reinterpret_cast<uint4*>(srow + 2048 + (((2 * c) ^ sw3) << 4)) = l0;
                *reinterpret_cast<uint4*>(srow + 2048 + (((2 * c + 1) ^ sw3) << 4)) = l1;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) gx2_arrive_leader(&tempty[abuf]);
            if (!(p.dbg & (16 | 64))) {                   // dbg 16 = timing ablation: no copy-out
                const int c0 = nt * (BN / 2) + h * GH, r0 = mt * BM + q * 32;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int pc = it * 32 + lane, rr = pc >> 2, un = pc & 3;
                    if (r0 + rr < p.M) {
                        const uint4 vh = *reinterpret_cast<const uint4*>(slab + rr * 64 + ((un ^ ((rr >> 1) & 3)) << 4));
                        const uint4 vl = *reinterpret_cast<const uint4*>(slab + 2048 + rr * 64 + ((un ^ ((rr >> 1) & 3)) << 4));
                        const long long o = (long long)(r0 + rr) * p.out_ld + c0 + un * 8;
                        *reinterpret_cast<uint4*>(p.out_h + o) = vh;
                        *reinterpret_cast<uint4*>(p.out_lo + o) = vl;
                    }
                }
            }
            __syncwarp();                                 // slab free for the next tile
            abuf ^= 1; if (abuf == 0) tphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    gx2_cluster_sync();                               // nobody leaves while the peer may still signal into this CTA
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}


template <int TAPS>
int launch_gate8x2_cfg(const UmmaConvParams& p, cudaStream_t s) {
    const int span = p.shift[p.taps - 1] - p.shift[0];
    const int box_rows = 128 + span;
    const int rows_alloc = (box_rows + 7) / 8 * 8;
    const size_t a_half = (size_t)rows_alloc * G_ROW_BYTES;
    const size_t smem = GX2_OUT_BYTES + G8_A16_STAGES * a_half + (size_t)GX2_W16_STAGES * GX2_WB + G8_A8_STAGES * 2 * a_half +
                        (size_t)GX2_W8_STAGES * GX2_W8_STAGE +
                        (2 * (G8_A16_STAGES + GX2_W16_STAGES + G8_A8_STAGES + GX2_W8_STAGES) + 4) * 8 + 32 + GX2_MAX_N * 4 + 1024;
    if (smem > 227 * 1024 || box_rows > 256 || p.N > GX2_MAX_N || p.B != 1 || p.rows_per_utt <= 0 || p.out_ld % 8 != 0 ||
        ((uintptr_t)p.out_h % 16) != 0 || ((uintptr_t)p.out_lo % 16) != 0)
        return CMTTS_ERR_UNSUPPORTED;
    auto kern = umma_gate8x2_kernel<TAPS>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cmtts_set_error("umma_gate8x2: cannot set dynamic shared memory size", __FILE__, __LINE__);
            return CMTTS_ERR_CUDA;
        }
        attr_done = true;
    }
    CUtensorMap a16, a8h, a8l, w16, w8h, w8l;
    // weight boxes: HALF an N tile (64 output channels) per CTA
    if (!make_act_map(&a16, p.a_hi, p.Cin, p.Lin, p.B, p.a_ld, p.a_bstride, G_BK, box_rows) ||
        !make_act_map8(&a8h, p.a8_hi, p.Cin, p.Lin, p.B, p.Cin, (long long)p.Lin * p.Cin, box_rows) ||
        !make_act_map8(&a8l, p.a8_lo, p.Cin, p.Lin, p.B, p.Cin, (long long)p.Lin * p.Cin, box_rows) ||
        !make_w_map(&w16, p.w_hi, p.Cin, p.taps * p.N, G_BK, G_BN / 2) ||
        !make_w_map8(&w8h, p.w8_hi, p.Cin, p.taps * p.N, G_BN / 2) || !make_w_map8(&w8l, p.w8_lo, p.Cin, p.taps * p.N, G_BN / 2)) {
        cmtts_set_error("umma_gate8x2: cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    const int m_tiles = (p.M + 127) / 128;
    const int tiles = p.B * ((m_tiles + 1) / 2) * (p.N / G_BN);
    int grid = 2 * tiles < num_sms() ? 2 * tiles : num_sms();
    grid &= ~1;
    if (grid < 2) return CMTTS_ERR_UNSUPPORTED;
    if (g_cmtts_prof_on) {
        const double rows = (double)p.B * p.M;
        char lbl[96];
        snprintf(lbl, sizeof(lbl), "umma_gate8x2<%d> t%d %d->%d (hi/lo, e4m3 cross terms, CTA pairs)", p.taps, p.taps, p.Cin, p.N);
        cmtts_prof_note(lbl, 2.0 * rows * p.N * p.taps * p.Cin,
                        rows * p.Cin * 4.0 + rows * (p.N / 2) * 4.0 + (double)p.taps * p.N * p.Cin * 4.0);
    }
    launch_pdl(kern, grid, G8_THREADS, smem, s, a16, a8h, a8l, w16, w8h, w8l, p, rows_alloc, box_rows);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

template <int TAPS>
int launch_gate8_cfg(const UmmaConvParams& p, cudaStream_t s) {
    const int span = p.shift[p.taps - 1] - p.shift[0];
    const int box_rows = 128 + span;
    const int rows_alloc = (box_rows + 7) / 8 * 8;
    const size_t a_half = (size_t)rows_alloc * G_ROW_BYTES;
    const size_t smem = G8_A16_STAGES * a_half + (size_t)G8_W16_STAGES * G_B_BYTES + G8_A8_STAGES * 2 * a_half +
                        (size_t)G8_W8_STAGES * G8_W8_STAGE +
                        (2 * (G8_A16_STAGES + G8_W16_STAGES + G8_A8_STAGES + G8_W8_STAGES) + 4) * 8 + 16 + 1024;
    if (smem > 227 * 1024 || box_rows > 256) return CMTTS_ERR_UNSUPPORTED;
    auto kern = umma_gate8_kernel<TAPS>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cmtts_set_error("umma_gate8: cannot set dynamic shared memory size", __FILE__, __LINE__);
            return CMTTS_ERR_CUDA;
        }
        attr_done = true;
    }
    CUtensorMap a16, a8h, a8l, w16, w8h, w8l;
    if (!make_act_map(&a16, p.a_hi, p.Cin, p.Lin, p.B, p.a_ld, p.a_bstride, G_BK, box_rows) ||
        !make_act_map8(&a8h, p.a8_hi, p.Cin, p.Lin, p.B, p.Cin, (long long)p.Lin * p.Cin, box_rows) ||
        !make_act_map8(&a8l, p.a8_lo, p.Cin, p.Lin, p.B, p.Cin, (long long)p.Lin * p.Cin, box_rows) ||
        !make_w_map(&w16, p.w_hi, p.Cin, p.taps * p.N, G_BK, G_BN) ||
        !make_w_map8(&w8h, p.w8_hi, p.Cin, p.taps * p.N, G_BN) || !make_w_map8(&w8l, p.w8_lo, p.Cin, p.taps * p.N, G_BN)) {
        cmtts_set_error("umma_gate8: cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    const int tiles = p.B * ((p.M + 127) / 128) * (p.N / G_BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    if (g_cmtts_prof_on) {
        const double rows = (double)p.B * p.M;
        char lbl[96];
        snprintf(lbl, sizeof(lbl), "umma_gate8<%d> t%d %d->%d (hi/lo, e4m3 cross terms)", p.taps, p.taps, p.Cin, p.N);
        cmtts_prof_note(lbl, 2.0 * rows * p.N * p.taps * p.Cin,
                        rows * p.Cin * 4.0 + rows * (p.N / 2) * 4.0 + (double)p.taps * p.N * p.Cin * 4.0);
    }
    launch_pdl(kern, grid, G8_THREADS, smem, s, a16, a8h, a8l, w16, w8h, w8l, p, rows_alloc, box_rows);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}


template <int TAPS, int CB>
int launch_gate_cfg(const UmmaConvParams& p, cudaStream_t s) {
    const int span = p.shift[p.taps - 1] - p.shift[0];
    const int box_rows = 128 + span;
    const int rows_alloc = (box_rows + 7) / 8 * 8;                          // 8 rows x 128 B = one swizzle atom
    const size_t smem = (size_t)G_A_STAGES * 2 * rows_alloc * G_ROW_BYTES + (size_t)G_B_STAGES * G_B_STAGE +
                        (2 * G_A_STAGES + 2 * G_B_STAGES + 4) * 8 + 16 + 1024;
    if (smem > 227 * 1024 || box_rows > 256) return CMTTS_ERR_UNSUPPORTED;
    auto kern = umma_gate_kernel<TAPS, CB>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cmtts_set_error("umma_gate: cannot set dynamic shared memory size", __FILE__, __LINE__);
            return CMTTS_ERR_CUDA;
        }
        attr_done = true;
    }
    CUtensorMap a0, a1, b0, b1;
    if (!make_act_map(&a0, p.a_hi, p.Cin, p.Lin, p.B, p.a_ld, p.a_bstride, G_BK, box_rows) ||
        !make_act_map(&a1, p.a_lo, p.Cin, p.Lin, p.B, p.a_ld, p.a_bstride, G_BK, box_rows) ||
        !make_w_map(&b0, p.w_hi, p.Cin, p.taps * p.N, G_BK, G_BN) ||
        !make_w_map(&b1, p.w_lo, p.Cin, p.taps * p.N, G_BK, G_BN)) {
        cmtts_set_error("umma_gate: cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    const int tiles = p.B * ((p.M + 127) / 128) * (p.N / G_BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    if (g_cmtts_prof_on) {
        const double rows = (double)p.B * p.M;
        char lbl[96];
        snprintf(lbl, sizeof(lbl), "umma_gate<%d> t%d %d->%d (hi/lo)", p.taps, p.taps, p.Cin, p.N);
        cmtts_prof_note(lbl, 2.0 * rows * p.N * p.taps * p.Cin,
                        rows * p.Cin * 4.0 + rows * (p.N / 2) * 4.0 + (double)p.taps * p.N * p.Cin * 4.0);
    }
    launch_pdl(kern, grid, 384, smem, s, a0, a1, b0, b1, p, rows_alloc, box_rows);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

}  // namespace

// Returns CMTTS_ERR_UNSUPPORTED (without setting an error) when the shape is not covered, so the caller can use the
// general kernel.  Covered: split operands, UEPI_DN_GATE, 3 uniformly spaced taps, Cin == 256, N % 128 == 0.
int launch_umma_gate(const UmmaConvParams& p, cudaStream_t s) {
    if (!p.split || p.epi != UEPI_DN_GATE || p.a2_hi || p.a_tap_dim || p.tap_split_n) return CMTTS_ERR_UNSUPPORTED;
    if (p.taps != 3 || p.Cin != 256 || p.N % G_BN != 0 || !p.bias || !p.a_lo || !p.w_lo) return CMTTS_ERR_UNSUPPORTED;
    if (p.shift[1] - p.shift[0] != p.shift[2] - p.shift[1] || p.shift[1] <= p.shift[0]) return CMTTS_ERR_UNSUPPORTED;
    if (((uintptr_t)p.bias % 16) != 0) return CMTTS_ERR_UNSUPPORTED;      // float4 bias loads
    if (p.B == 0 || p.M == 0) return CMTTS_OK;
    // e4m3 cross terms when the caller supplies the fp8 operand copies (CMTTS_UMMA_DBG bit 256 keeps them in fp16)
    if (p.a8_hi && p.a8_lo && p.w8_hi && p.w8_lo && !(p.dbg & 256) && p.B == 1) {
        static int pair_env = -1;                          // CMTTS_GATE_PAIR=0: one-CTA kernel instead of the CTA-pair (cta_group::2) one
        if (pair_env < 0) { const char* e = getenv("CMTTS_GATE_PAIR"); pair_env = e ? atoi(e) : 1; }
        if (pair_env && !(p.dbg & 1024)) {
            const int rc = launch_gate8x2_cfg<3>(p, s);
            if (rc != CMTTS_ERR_UNSUPPORTED) return rc;
        }
        return launch_gate8_cfg<3>(p, s);
    }
    return launch_gate_cfg<3, 4>(p, s);
}
