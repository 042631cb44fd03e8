// Fused HiFi-GAN ResBlock iteration on tcgen05:   y' = c2(lrelu(c1(lrelu(y)) + b1)) + b2 + y
// (hifigan/models.py:96-103) for the HBM-bound levels (C = 64, 32).
//
// The unfused pipeline moves every activation 5.4 times per iteration (read y, write t, read t with
// halo, read y again as the residual, write y').  Here one kernel does both convolutions per tile:
//   * the input halo tile (128 + (k-1)*dil rows of lrelu(y), fp16) is fetched ONCE by TMA; conv1's
//     taps are row offsets into it (same trick as umma_halo.cu);
//   * conv1's accumulator (TMEM) goes through bias + leaky-ReLU in the epilogue warps and is written
//     as fp16 into a shared-memory tile in the UMMA K-major swizzled layout — it never sees HBM;
//     rows outside the utterance are written as zeros (conv2 zero-pads t, not c1(padding));
//   * conv2 (dilation 1) reads that tile, again by row offsets; each tile yields 128 - (k-1) valid
//     output rows (92-98 % of the MMA rows);
//   * the residual y comes from the input halo tile itself (its rows P2 + p1d .. are the output rows; stored as
//     lrelu(y), inverted exactly in the epilogue): an input stage stays allocated until epilogue 2 of its tile
//     has read them.  (It used to be re-read from global memory one tile ahead, "L2-hot" — ncu showed 60 % of
//     those reads going to DRAM, 660 MB read per launch for 416 MB of input, and the epilogue-2 warps waiting on
//     them 40 % of their time: ~10 tiles x 148 CTAs x 64 KB of traffic pass between the TMA load of a tile and its
//     epilogue 2, about the size of the L2.)  The output is staged in the TMA box layout and TMA-stored.
// HBM traffic per iteration: read y once (+halo), write y' once.
//
// Pipelining: conv2 of a tile can only start after conv1 -> commit -> epilogue 1 -> t tile, a chain of
// ~1.5k cycles, so conv1 runs up to nb-1 tiles AHEAD of conv2 (separate issuing warps; conv1 accumulators,
// t tiles and conv2 accumulators are nb-deep rings, nb = 3 or 4).  Epilogue 1
// (TMEM -> t tile) and epilogue 2 (output) run on separate warps concurrently — EIGHT warps each (lane
// quarter x column half): a lone warp per SM sub-partition issues its dependent ALU chain at ~4 cycles per
// instruction, which made the epilogues (not HBM, not the tensor pipe) the limiter (ncu: MMA warp stalled on
// acc2_empty, epilogue-2 warps > 80 % busy).  Biases sit in shared memory; leaky-ReLU is max / min(v, s v).
// An input stage is released by the conv1 commit AND the eight epilogue-2 warps (9 arrivals).
//
// Roles (608 threads): warp 0 TMEM allocator + TMA producer, 1 conv1 issuer, 2 conv2 issuer, 3-10 epilogue 1,
// 11-18 epilogue 2 (an epilogue warp's TMEM lane quarter is warp % 4, its column half the warp's position in its
// group of eight).
// Both weight sets stay resident in shared memory for the whole persistent loop.
//
// PAIRED modes (MODE 1 / 2; C = 32 tensors, kernel instantiated with C = 64).  At C = 32 an M = 128, N = 32, K = 16 MMA
// spends 32 cycles reading its A slab from shared memory for 16 cycles of math: the tensor pipe cannot exceed 50 % and
// the plain kernel sat at that cap for k = 7 / 11 (ncu: 644 / 706 TFLOP/s).  Here the (L, 32) tensors are viewed as
// (L / 2, 64): one 128-byte row holds the time steps 2 r and 2 r + 1, and one accumulator row holds both outputs
// (N = 64: columns [0, 32) = step 2 r, [32, 64) = step 2 r + 1).  For a dilation-1 conv the two outputs of a row share
// K - 1 of their K input positions: the conv becomes K + 1 "units" of K = 32 (one per input position e = -h .. h + 1,
// a 64-byte half row), each with the N = 64 weight block [W[e + h] ; W[e + h - 1]] (zero where the tap does not
// exist; weights.py: pair_pack_d1) — 2 (K + 1) MMAs per 256 time steps instead of 4 K.  conv2 always has dilation 1;
// conv1 has it in the first iteration of a ResBlock (MODE 1).  For odd dilations > 1 (MODE 2) the two outputs share
// no input position: conv1 runs its K taps as two N = 32 MMAs each (the same weight block W[k], input halves at
// positions 1 + k d and 2 + k d, accumulator columns [0, 32) and [32, 64)) — the MMA time of the plain kernel, but on
// 128-byte TMA rows and with half as many tiles, epilogue hand-offs and barrier round trips per time step.  Epilogues,
// rings and the TMA store are those of the C = 64 instantiation on the paired view (bias duplicated).
#include "umma_common.cuh"
#include <stdlib.h>

namespace {

using namespace umma;

constexpr int RB_MAX_STAGES = 8;
constexpr int RB_MAX_NB = 4;           // depth of the conv1-accumulator / t-tile rings
constexpr int T_ROWS_ALLOC = 144;      // 128 + (k_max - 1) rounded to the swizzle atom

struct RbCfg {
    int a_stages, rows_alloc, box_rows, valid, m_tiles, nb, nb2;   // nb2: depth of the conv2-accumulator ring (2..4)
    int p1d;   // conv1 half-span in rows of the halo tile
    int pf;    // L2 prefetch distance of the halo tiles, in tiles (0 = off; CMTTS_PF)
    int dbg;   // timing ablations (CMTTS_RB_DBG, results become wrong): 1 epilogue 2 does no work, 2 epilogue 1 does no
               // work, 4 no MMAs are issued, 8 no residual read, 16 no output store
};

__device__ __forceinline__ float lrelu_fwd(float v, float slope) { return fmaxf(v, v * slope); }          // slope <= 1
__device__ __forceinline__ float lrelu_inv(float v, float inv_slope) { return fminf(v, v * inv_slope); }  // inv_slope >= 1

constexpr int RB_THREADS = 608;          // 19 warps

template <int C, int TAPS, bool HAS_SUM, int MODE>
__global__ void __launch_bounds__(RB_THREADS, 1)
umma_resblock_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO,
                     const __grid_constant__ CUtensorMap tmOtail, const UmmaResblockParams p, const RbCfg cfg) {
    constexpr int BM = 128;
    constexpr int BK = C;                          // one channel block (C = 64 -> SW128, C = 32 -> SW64)
    constexpr int ROW_BYTES = BK * 2;
    constexpr int W_BLK = C * ROW_BYTES;           // one tap of one conv
    constexpr int T_ALLOC = T_ROWS_ALLOC * ROW_BYTES;
    constexpr int CH = C / 2;                      // columns per epilogue warp
    constexpr int OROW = CH * 2;                   // bytes per staged output row of one epilogue-2 warp
    constexpr int O_SLAB = 32 * OROW;
    constexpr bool TMA_STORE = C >= 64;            // C = 32: 32-byte half rows, direct stores
    constexpr int TMEM_COLS = pow2_cols(2 * RB_MAX_NB * C);
    static_assert(MODE == 0 || C == 64, "paired modes run on the (L/2, 64) view");
    constexpr int NBLK = (TAPS + 1) / 2;           // paired modes: weight blocks of two K = 32 units each (128-byte rows)
    constexpr int W1_BYTES = MODE == 1 ? NBLK * 64 * 128 : MODE == 2 ? NBLK * 32 * 128 : TAPS * W_BLK;
    constexpr int W2_BYTES = MODE ? NBLK * 64 * 128 : TAPS * W_BLK;
    // conv2 half-span in rows of the tile: plain (TAPS-1)/2; paired: input positions -h .. h+1 = pair rows -(h+1)/2 .. (h+1)/2
    constexpr int P2 = MODE ? ((TAPS - 1) / 2 + 1) / 2 : (TAPS - 1) / 2;
    constexpr uint32_t DESC_HI = (uint32_t)((8 * ROW_BYTES) >> 4) | (1u << 14) | ((BK == 64 ? 2u : 4u) << 29);

    const int a_alloc = cfg.rows_alloc * ROW_BYTES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* smA = smem;
    uint8_t* smT = smA + cfg.a_stages * a_alloc;
    uint8_t* smO = smT + cfg.nb * T_ALLOC;
    uint8_t* smW1 = smO + 8 * O_SLAB;
    uint8_t* smW2 = smW1 + W1_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smW2 + W2_BYTES);
    uint64_t* a_empty = a_full + RB_MAX_STAGES;
    uint64_t* w_full = a_empty + RB_MAX_STAGES;    // [1]
    uint64_t* acc1_full = w_full + 1;              // [RB_MAX_NB]
    uint64_t* acc1_empty = acc1_full + RB_MAX_NB;
    uint64_t* t_full = acc1_empty + RB_MAX_NB;
    uint64_t* t_empty = t_full + RB_MAX_NB;
    uint64_t* acc2_full = t_empty + RB_MAX_NB;     // [RB_MAX_NB]
    uint64_t* acc2_empty = acc2_full + RB_MAX_NB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + RB_MAX_NB);
    float* s_b1 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);   // float4 reads
    float* s_b2 = s_b1 + C;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: warp index provably uniform for ptxas
    const int tiles = p.B * cfg.m_tiles;
    const int p1d = cfg.p1d;                       // conv1 half-span in rows of the tile (plain: (TAPS-1)/2 * dil)
    const int Lr = MODE ? (p.L >> 1) : p.L;        // rows of the tensors as this kernel sees them

    if (threadIdx.x == 0) {
        for (int i = 0; i < RB_MAX_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 9); }   // conv1 commit + 8 epilogue-2 warps
        mbar_init(&w_full[0], 1);
        for (int i = 0; i < RB_MAX_NB; ++i) {
            mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], 8);
            mbar_init(&t_full[i], 8); mbar_init(&t_empty[i], 1);
            mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 96 && threadIdx.x < 96 + C) {
        const int ci = MODE ? ((threadIdx.x - 96) & 31) : (threadIdx.x - 96);      // paired: [b | b]
        s_b1[threadIdx.x - 96] = p.b1[ci]; s_b2[threadIdx.x - 96] = p.b2[ci];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    if (warp != 0) pdl_wait();      // PDL: warp 0 issues the (constant) resident weights first, see below
    // TMEM columns: acc1[b] at b*C (b < nb), acc2[b] at RB_MAX_NB*C + b*C
    const int nb = cfg.nb;

    if (warp == 0) {
        // ======================= producer: resident weights, then one halo tile per output tile =======
        {
            prefetch_tmap(&tmA); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
            mbar_expect_tx_elect(&w_full[0], (uint32_t)(W1_BYTES + W2_BYTES));
            if constexpr (MODE == 0) {
                for (int tap = 0; tap < TAPS; ++tap) {
                    tma_load_2d_elect(smW1 + tap * W_BLK, &tmW1, &w_full[0], 0, tap * C);
                    tma_load_2d_elect(smW2 + tap * W_BLK, &tmW2, &w_full[0], 0, tap * C);
                }
            } else {
                constexpr int R1 = MODE == 1 ? 64 : 32;       // rows of one conv1 weight block
                for (int blk = 0; blk < NBLK; ++blk) {
                    tma_load_2d_elect(smW1 + blk * R1 * 128, &tmW1, &w_full[0], 0, blk * R1);
                    tma_load_2d_elect(smW2 + blk * 64 * 128, &tmW2, &w_full[0], 0, blk * 64);
                }
            }
            pdl_wait();
            int stage = 0; uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)(((cfg.dbg & 32) ? 16 : cfg.box_rows) * ROW_BYTES);   // 32: 16-row boxes (timing)
            // optional L2 prefetch of the halo tiles PF tiles ahead of the ring (CMTTS_PF, off by default: measured no gain,
            // profiles/ablation_r1_ring.txt — the kernel is not bound by the latency of its input loads)
            const int PF = cfg.pf;
            for (int i = 0; i < PF; ++i) {
                const int tl = blockIdx.x + i * gridDim.x;
                if (tl < tiles) tma_prefetch_3d_elect(&tmA, 0, (tl % cfg.m_tiles) * cfg.valid - P2 - p1d, tl / cfg.m_tiles);
            }
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int mt = tile % cfg.m_tiles, b = tile / cfg.m_tiles;
                const int row0 = mt * cfg.valid - P2 - p1d;          // first input row of the halo tile
                if (PF > 0) {
                    const int tl = tile + PF * gridDim.x;
                    if (tl < tiles) tma_prefetch_3d_elect(&tmA, 0, (tl % cfg.m_tiles) * cfg.valid - P2 - p1d, tl / cfg.m_tiles);
                }
                mbar_wait(&a_empty[stage], phase ^ 1);
                mbar_expect_tx_elect(&a_full[stage], bytes);
                tma_load_3d_elect(smA + stage * a_alloc, &tmA, &a_full[stage], 0, row0, b);
                if (++stage == cfg.a_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ======================= conv1 issuer =======================
        // (the whole warp runs the loop convergently; one lane's tcgen05 instructions are predicated on.)
        // Every tcgen05.mma costs ~10 uniform-datapath instructions of descriptor set-up (~60-70 cycles), more than
        // a 128 x C x 16 MMA occupies the tensor pipe for C <= 64, so conv1 and conv2 are issued by TWO warps:
        // a single issuer left the pipe idle ~45 % of the time with every other role waiting on it (ncu).
        {
            // Issue form: `if (elect_one())` blocks, operands derived from warp-uniform values only (umma_common.cuh)
            const uint32_t tmem_u = make_uniform(tmem_base);
            const uint32_t idesc = make_idesc(BM, C);
            // rows (plain) or 64-byte positions (paired) between conv1 taps, in 16-byte descriptor units
            const uint32_t tap_step1 = make_uniform(MODE ? (uint32_t)(p.dil * 4) : (uint32_t)((p.dil * ROW_BYTES) >> 4));
            const uint32_t w1_lo = make_uniform(((smem_u32(smW1) >> 4) & 0x3FFF) | (1u << 16));
            const uint32_t a_base = make_uniform(((smem_u32(smA) >> 4) & 0x3FFF) | (1u << 16));
            const uint32_t a_alloc16 = make_uniform((uint32_t)(a_alloc >> 4));
            mbar_wait(&w_full[0], 0);
            tc_fence_after();
            int stage = 0; uint32_t phase = 0;
            int b1 = 0; uint32_t ph1 = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                mbar_wait(&acc1_empty[b1], ph1 ^ 1);
                mbar_wait(&a_full[stage], phase);
                tc_fence_after();
                const uint32_t d_tmem = make_uniform(tmem_u + (uint32_t)(b1 * C));
                const uint32_t a_lo = make_uniform(a_base + (uint32_t)stage * a_alloc16);
                if (elect_one()) {
                    if (!(cfg.dbg & 4)) {
                    if constexpr (MODE == 0) {
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16(d_tmem, ((uint64_t)DESC_HI << 32) | (a_lo + tap * tap_step1 + 2 * k),
                                     ((uint64_t)DESC_HI << 32) | (w1_lo + (uint32_t)((tap * W_BLK) >> 4) + 2 * k), idesc,
                                     (tap | k) ? 1u : 0u);
                    }
                    } else if constexpr (MODE == 1) {
                        // dilation 1, paired: unit u = input position u - h = byte offset (u + 1) * 64 of the halo tile
#pragma unroll
                        for (int u = 0; u <= TAPS; ++u) {
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                umma_f16(d_tmem, ((uint64_t)DESC_HI << 32) | (a_lo + (uint32_t)((u + 1) * 4 + 2 * k)),
                                         ((uint64_t)DESC_HI << 32) | (w1_lo + (uint32_t)((u >> 1) * 512 + (u & 1) * 4 + 2 * k)), idesc,
                                         (u | k) ? 1u : 0u);
                        }
                    } else {
                        // odd dilation > 1, paired: tap j reads positions 1 + j d (even output) and 2 + j d (odd output)
                        const uint32_t idesc32 = make_idesc(BM, 32);
#pragma unroll
                        for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const uint32_t a_e = a_lo + 4u + (uint32_t)tap * tap_step1 + (uint32_t)(2 * k);
                                const uint64_t bd = ((uint64_t)DESC_HI << 32) | (w1_lo + (uint32_t)((tap >> 1) * 256 + (tap & 1) * 4 + 2 * k));
                                umma_f16(d_tmem, ((uint64_t)DESC_HI << 32) | a_e, bd, idesc32, (tap | k) ? 1u : 0u);
                                umma_f16(d_tmem + 32u, ((uint64_t)DESC_HI << 32) | (a_e + 4u), bd, idesc32, (tap | k) ? 1u : 0u);
                            }
                        }
                    }
                    }
                    umma_commit(&a_empty[stage]);      // input stage free once conv1 has read it
                    umma_commit(&acc1_full[b1]);
                }
                __syncwarp();
                if (++stage == cfg.a_stages) { stage = 0; phase ^= 1; }
                if (++b1 == nb) { b1 = 0; ph1 ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ======================= conv2 issuer =======================
        {
            const uint32_t tmem_u = make_uniform(tmem_base);
            const uint32_t idesc = make_idesc(BM, C);
            constexpr uint32_t tap_step2 = (uint32_t)(ROW_BYTES >> 4);
            const uint32_t w2_lo = make_uniform(((smem_u32(smW2) >> 4) & 0x3FFF) | (1u << 16));
            const uint32_t t_base = make_uniform(((smem_u32(smT) >> 4) & 0x3FFF) | (1u << 16));
            mbar_wait(&w_full[0], 0);
            tc_fence_after();
            int b2 = 0; uint32_t ph2 = 0;                     // t-tile ring
            int ab = 0; uint32_t aph = 0;                     // conv2 accumulator ring
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                mbar_wait(&t_full[b2], ph2);
                mbar_wait(&acc2_empty[ab], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = make_uniform(tmem_u + (uint32_t)(RB_MAX_NB * C + ab * C));
                const uint32_t t_lo = make_uniform(t_base + (uint32_t)b2 * (uint32_t)(T_ALLOC >> 4));
                if (elect_one()) {
                    if (!(cfg.dbg & 4)) {
                    if constexpr (MODE == 0) {
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16(d_tmem, ((uint64_t)DESC_HI << 32) | (t_lo + tap * tap_step2 + 2 * k),
                                     ((uint64_t)DESC_HI << 32) | (w2_lo + (uint32_t)((tap * W_BLK) >> 4) + 2 * k), idesc,
                                     (tap | k) ? 1u : 0u);
                    }
                    } else {
#pragma unroll
                        for (int u = 0; u <= TAPS; ++u) {
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                umma_f16(d_tmem, ((uint64_t)DESC_HI << 32) | (t_lo + (uint32_t)((u + 1) * 4 + 2 * k)),
                                         ((uint64_t)DESC_HI << 32) | (w2_lo + (uint32_t)((u >> 1) * 512 + (u & 1) * 4 + 2 * k)), idesc,
                                         (u | k) ? 1u : 0u);
                        }
                    }
                    }
                    umma_commit(&t_empty[b2]);
                    umma_commit(&acc2_full[ab]);
                }
                __syncwarp();
                if (++b2 == nb) { b2 = 0; ph2 ^= 1; }
                if (++ab == cfg.nb2) { ab = 0; aph ^= 1; }
            }
        }
    } else {
        // ======================= epilogue warps =======================
        const bool is_e1 = warp < 11;     // warps 3-10: epilogue 1 (conv1 -> t tile); warps 11-18: epilogue 2 (output)
        const int q = warp & 3;                                             // TMEM lane quarter
        const int h = ((warp - 3) >> 2) & 1;                                // column half
        const int ew = (warp - 3) & 7;                                      // index inside the group of eight
        const int row = q * 32 + lane;
        const int n_base = h * CH;
        const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
        int it = 0;
        if (is_e1) {
            // ---- epilogue 1: t = lrelu(c1 + b1) -> fp16 swizzled smem tile (zeros outside the utterance)
            const int swz = (BK == 64) ? (row & 7) : ((row >> 1) & 3);
            int bb = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
                const int mt = tile % cfg.m_tiles;
                mbar_wait(&acc1_full[bb], ph);
                mbar_wait(&t_empty[bb], ph ^ 1);
                tc_fence_after();
                const int trow = mt * cfg.valid - P2 + row;                 // global row of this t row
                const bool inside = trow >= 0 && trow < Lr;
                const uint32_t taddr = tmem_base + (uint32_t)(bb * C + n_base) + lane_addr;
                uint8_t* trow_ptr = smT + bb * T_ALLOC + row * ROW_BYTES;
                if (!(cfg.dbg & 2))
#pragma unroll
                for (int c = 0; c < CH / 16; ++c) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c * 16, r);
                    tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bq = *reinterpret_cast<const float4*>(s_b1 + n_base + c * 16 + 4 * j);
                        v[4 * j] = lrelu_fwd(__uint_as_float(r[4 * j]) + bq.x, p.t_slope);
                        v[4 * j + 1] = lrelu_fwd(__uint_as_float(r[4 * j + 1]) + bq.y, p.t_slope);
                        v[4 * j + 2] = lrelu_fwd(__uint_as_float(r[4 * j + 2]) + bq.z, p.t_slope);
                        v[4 * j + 3] = lrelu_fwd(__uint_as_float(r[4 * j + 3]) + bq.w, p.t_slope);
                    }
                    uint4 u0 = make_uint4(0u, 0u, 0u, 0u), u1 = u0;
                    if (inside) {
                        __half2* p0 = reinterpret_cast<__half2*>(&u0);
                        __half2* p1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            p0[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                            p1[i] = __floats2half2_rn(v[8 + 2 * i], v[8 + 2 * i + 1]);
                        }
                    }
                    const int j0 = n_base / 8 + 2 * c;                      // 16-byte chunk index inside the row
                    *reinterpret_cast<uint4*>(trow_ptr + ((j0 ^ swz) << 4)) = u0;
                    *reinterpret_cast<uint4*>(trow_ptr + (((j0 + 1) ^ swz) << 4)) = u1;
                }
                tc_fence_before();
                fence_proxy_async_smem();          // t tile (generic-proxy writes) -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) { mbar_arrive(&acc1_empty[bb]); mbar_arrive(&t_full[bb]); }
                if (++bb == nb) { bb = 0; ph ^= 1; }
            }
        } else {
            // ---- epilogue 2: y' = lrelu_out(c2 + b2 + inv_lrelu(a) [+ sum]) -> staged + TMA store (direct for C = 32)
            const int swo = (OROW == 64) ? ((lane >> 1) & 3) : (lane & 1);
            uint8_t* slab = smO + ew * O_SLAB + lane * OROW;
            // MRF partial sum of this thread's output row: global reads, software-pipelined ONE TILE AHEAD so that their
            // latency never sits between the accumulator wait and the stores.  The residual lrelu(y) is read from the
            // input stage in shared memory (rows P2 + p1d + row of the halo tile, swizzled like the TMA wrote them).
            uint4 snext[HAS_SUM ? CH / 8 : 1];
            auto prefetch = [&](int tile_) {
                if constexpr (HAS_SUM) {
                    const int mt_ = tile_ % cfg.m_tiles, b_ = tile_ / cfg.m_tiles;
                    const int o_ = mt_ * cfg.valid + row;
                    const bool ok = tile_ < tiles && row < cfg.valid && o_ < Lr;
                    const long long g_ = ((long long)b_ * Lr + o_) * C + n_base;
#pragma unroll
                    for (int i = 0; i < CH / 8; ++i)
                        snext[i] = ok ? reinterpret_cast<const uint4*>(p.sum_h + g_)[i] : make_uint4(0u, 0u, 0u, 0u);
                }
            };
            prefetch((int)blockIdx.x);
            const int arow = row + P2 + p1d;                                // this thread's row of the input halo tile
            const int asw = (BK == 64) ? (arow & 7) : ((arow >> 1) & 3);
            int st = 0; uint32_t sph = 0;                                   // input-stage ring position of the tile
            int bb = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
                const int mt = tile % cfg.m_tiles, b = tile / cfg.m_tiles;
                const int o = mt * cfg.valid + row;                         // output row of this thread
                const bool valid = row < cfg.valid && o < Lr;
                const long long goff = ((long long)b * Lr + o) * C + n_base;
                uint4 rres[CH / 8], rsum[HAS_SUM ? CH / 8 : 1];
                mbar_wait(&a_full[st], sph);                                // (long complete: conv1 of this tile has run)
                {
                    const uint8_t* ap = smA + st * a_alloc + arow * ROW_BYTES;
#pragma unroll
                    for (int i = 0; i < CH / 8; ++i)
                        rres[i] = (cfg.dbg & 8) ? make_uint4(0u, 0u, 0u, 0u)
                                                : *reinterpret_cast<const uint4*>(ap + (((n_base / 8 + i) ^ asw) << 4));
                }
#pragma unroll
                for (int i = 0; i < CH / 8; ++i) {
                    if constexpr (HAS_SUM) rsum[i] = snext[i];
                }
                prefetch(tile + (int)gridDim.x);
                mbar_wait(&acc2_full[bb], ph);
                tc_fence_after();
                if (TMA_STORE) {
                    if (lane == 0) tma_store_wait_read();                   // previous store has finished reading the slab
                    __syncwarp();
                }
                const uint32_t taddr = tmem_base + (uint32_t)(RB_MAX_NB * C + bb * C + n_base) + lane_addr;
                if (!(cfg.dbg & 1))
#pragma unroll
                for (int c = 0; c < CH / 16; ++c) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c * 16, r);
                    tmem_ld_wait();
                    const int n = n_base + c * 16;
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bq = *reinterpret_cast<const float4*>(s_b2 + n + 4 * j);
                        v[4 * j] = fmaf(__uint_as_float(r[4 * j]), p.alpha2, bq.x);
                        v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), p.alpha2, bq.y);
                        v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), p.alpha2, bq.z);
                        v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), p.alpha2, bq.w);
                    }
                    const __half2* h0 = reinterpret_cast<const __half2*>(&rres[2 * c]);
                    const __half2* h1 = reinterpret_cast<const __half2*>(&rres[2 * c + 1]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 a = __half22float2(h0[i]), bq = __half22float2(h1[i]);
                        v[2 * i] += lrelu_inv(a.x, p.res_inv_slope); v[2 * i + 1] += lrelu_inv(a.y, p.res_inv_slope);
                        v[8 + 2 * i] += lrelu_inv(bq.x, p.res_inv_slope); v[8 + 2 * i + 1] += lrelu_inv(bq.y, p.res_inv_slope);
                    }
                    if constexpr (HAS_SUM) {
                        const __half2* s0 = reinterpret_cast<const __half2*>(&rsum[2 * c]);
                        const __half2* s1 = reinterpret_cast<const __half2*>(&rsum[2 * c + 1]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 a = __half22float2(s0[i]), bq = __half22float2(s1[i]);
                            v[2 * i] += a.x; v[2 * i + 1] += a.y; v[8 + 2 * i] += bq.x; v[8 + 2 * i + 1] += bq.y;
                        }
                    }
                    uint4 u0, u1;
                    __half2* p0 = reinterpret_cast<__half2*>(&u0);
                    __half2* p1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        p0[i] = __floats2half2_rn(lrelu_fwd(v[2 * i], p.out_slope), lrelu_fwd(v[2 * i + 1], p.out_slope));
                        p1[i] = __floats2half2_rn(lrelu_fwd(v[8 + 2 * i], p.out_slope), lrelu_fwd(v[8 + 2 * i + 1], p.out_slope));
                    }
                    if (TMA_STORE) {
                        *reinterpret_cast<uint4*>(slab + (((2 * c) ^ swo) << 4)) = u0;
                        *reinterpret_cast<uint4*>(slab + (((2 * c + 1) ^ swo) << 4)) = u1;
                    } else if (valid && !(cfg.dbg & 16)) {
                        __half* op = p.out_h + goff + c * 16;
                        *reinterpret_cast<uint4*>(op) = u0;
                        *reinterpret_cast<uint4*>(op + 8) = u1;
                    }
                }
                tc_fence_before();
                if (TMA_STORE) fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&acc2_empty[bb]);
                    mbar_arrive(&a_empty[st]);      // residual rows consumed (their values went through the adds above)
                    if (TMA_STORE && !(cfg.dbg & 17)) {
                        // rows [q*32, q*32+32) of the tile; the last lane quarter only owns `valid - 96` rows
                        const int r0 = mt * cfg.valid + q * 32;
                        if (q < 3) tma_store_3d(&tmO, smO + ew * O_SLAB, n_base, r0, b);
                        else tma_store_3d(&tmOtail, smO + ew * O_SLAB, n_base, r0, b);
                        tma_store_commit();
                    }
                }
                if (++bb == cfg.nb2) { bb = 0; ph ^= 1; }
                if (++st == cfg.a_stages) { st = 0; sph ^= 1; }
            }
            if (TMA_STORE && lane == 0) tma_store_wait_all();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int C, int TAPS, bool HAS_SUM, int MODE>
int launch_rb_cfg(const UmmaResblockParams& p, cudaStream_t s) {
    constexpr int BK = C;
    constexpr int ROW_BYTES = BK * 2;
    constexpr int ROW_ALIGN = 1024 / ROW_BYTES;
    constexpr int NBLK = (TAPS + 1) / 2;
    constexpr int R1 = MODE == 1 ? 64 : 32;                // rows of one paired conv1 weight block
    constexpr size_t W_BYTES = MODE ? (size_t)NBLK * (R1 + 64) * 128 : (size_t)2 * TAPS * C * ROW_BYTES;
    constexpr size_t T_ALLOC = (size_t)T_ROWS_ALLOC * ROW_BYTES;
    constexpr size_t O_BYTES = (size_t)8 * 32 * C;         // 8 epilogue-2 warps x 32 rows x C/2 fp16
    constexpr size_t FIXED = (2 * RB_MAX_STAGES + 1 + 6 * RB_MAX_NB) * 8 + 32 + 2 * C * 4 + 1024;
    constexpr size_t LIMIT = 227 * 1024;
    static_assert((T_ROWS_ALLOC * ROW_BYTES) % 1024 == 0, "t tile must keep the swizzle alignment");

    RbCfg cfg{};
    constexpr int H = (TAPS - 1) / 2;
    const int Lr = MODE ? p.L / 2 : p.L;                   // rows as the kernel sees them (paired: two time steps per row)
    if (MODE) {
        constexpr int HQ = (H + 1) / 2;                    // pair rows either side of a dilation-1 conv
        cfg.valid = 128 - 2 * HQ;
        cfg.p1d = (H * p.dil + 1) / 2;                     // positions -H d .. H d + 1
    } else {
        cfg.valid = 128 - (TAPS - 1);
        cfg.p1d = H * p.dil;
    }
    cfg.box_rows = 128 + 2 * cfg.p1d;
    if (cfg.box_rows > 256) return CMTTS_ERR_UNSUPPORTED;
    cfg.rows_alloc = (cfg.box_rows + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    cfg.m_tiles = (Lr + cfg.valid - 1) / cfg.valid;
    const size_t a_alloc = (size_t)cfg.rows_alloc * ROW_BYTES;
    // An input stage lives from its TMA load to the end of epilogue 2 of its tile (the residual is read from it), i.e.
    // across the whole conv1 -> epilogue 1 -> conv2 -> epilogue 2 chain: the deepest conv1 -> conv2 lag nb that still
    // leaves nb + 2 input stages, else nb = 2 with at least 2 (CMTTS_RB_NB overrides, 2..4)
    static int nb_env = -1;
    if (nb_env < 0) { const char* e = getenv("CMTTS_RB_NB"); nb_env = e ? atoi(e) : 0; }
    cfg.nb = 0;
    size_t rest = 0;
    for (int nb = RB_MAX_NB; nb >= 2; --nb) {
        if (nb_env >= 2 && nb_env <= RB_MAX_NB && nb != nb_env) continue;
        rest = W_BYTES + (size_t)nb * T_ALLOC + O_BYTES + FIXED;
        const size_t min_stages = (nb == 2 || nb == nb_env) ? 2 : (size_t)nb + 2;
        if (rest + min_stages * a_alloc <= LIMIT) { cfg.nb = nb; break; }
    }
    if (cfg.nb == 0) return CMTTS_ERR_UNSUPPORTED;
    static int nb2_env = -1;
    if (nb2_env < 0) { const char* e = getenv("CMTTS_RB_NB2"); nb2_env = e ? atoi(e) : 0; }
    cfg.nb2 = (nb2_env >= 2 && nb2_env <= RB_MAX_NB) ? nb2_env : RB_MAX_NB;
    static int dbg_env = -1;
    if (dbg_env < 0) { const char* e = getenv("CMTTS_RB_DBG"); dbg_env = e ? atoi(e) : 0; }
    cfg.dbg = dbg_env;
    static int pf_env = -1;
    if (pf_env < 0) { const char* e = getenv("CMTTS_PF"); pf_env = e ? atoi(e) : 0; }
    cfg.pf = pf_env;
    size_t st = (LIMIT - rest) / a_alloc;
    cfg.a_stages = (int)(st > RB_MAX_STAGES ? RB_MAX_STAGES : st);
    const size_t smem = rest + (size_t)cfg.a_stages * a_alloc;

    auto kern = umma_resblock_kernel<C, TAPS, HAS_SUM, MODE>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIMIT) != cudaSuccess) {
            cmtts_set_error("umma_resblock: cannot set dynamic shared memory size", __FILE__, __LINE__);
            return CMTTS_ERR_CUDA;
        }
        attr_done = true;
    }
    CUtensorMap a_map, w1_map, w2_map, o_map, ot_map;
    const long long bs = (long long)Lr * C;
    bool ok = make_act_map(&a_map, p.a, C, Lr, p.B, C, bs, BK, (cfg.dbg & 32) ? 16 : cfg.box_rows) &&
              make_act_map(&o_map, p.out_h, C, Lr, p.B, C, bs, 32, 32) &&              // per-warp box: C/2 (<= 32) channels
              make_act_map(&ot_map, p.out_h, C, Lr, p.B, C, bs, 32, cfg.valid - 96);
    if (MODE) ok = ok && make_w_map(&w1_map, p.w1p, 64, NBLK * R1, 64, R1) && make_w_map(&w2_map, p.w2p, 64, NBLK * 64, 64, 64);
    else ok = ok && make_w_map(&w1_map, p.w1, C, TAPS * C, BK, C) && make_w_map(&w2_map, p.w2, C, TAPS * C, BK, C);
    if (!ok) {
        cmtts_set_error("umma_resblock: cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    const int tiles = p.B * cfg.m_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    if (g_cmtts_prof_on) {
        const double rows = (double)p.B * p.L;
        char lbl[96];
        snprintf(lbl, sizeof(lbl), "umma_resblock<%d%s> k%d d%d%s (conv1+conv2)", p.C, MODE ? ", paired rows" : "", p.taps, p.dil, p.sum_h ? " +sum" : "");
        cmtts_prof_note(lbl, 2.0 * 2.0 * rows * p.C * p.C * p.taps,
                        rows * p.C * 2.0 * 2.0 + (p.sum_h ? rows * p.C * 2.0 : 0.0) + 2.0 * p.taps * p.C * p.C * 2.0);
    }
    launch_pdl(kern, grid, RB_THREADS, smem, s, a_map, w1_map, w2_map, o_map, ot_map, p, cfg);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

}  // namespace

// CMTTS_ERR_UNSUPPORTED (no error text) when the shape is not covered -> caller runs the two convs separately.
int launch_umma_resblock(const UmmaResblockParams& p, cudaStream_t s) {
    if (p.B == 0 || p.L == 0) return CMTTS_OK;
    if (p.dil < 1 || !p.a || !p.w1 || !p.w2 || !p.b1 || !p.b2 || !p.out_h) return CMTTS_ERR_UNSUPPORTED;
    if (((uintptr_t)p.a % 16) || ((uintptr_t)p.out_h % 16)) return CMTTS_ERR_UNSUPPORTED;
    // max / min form of leaky-ReLU: forward slopes in (0, 1], inverse slope >= 1
    if (!(p.t_slope > 0.f && p.t_slope <= 1.f && p.out_slope > 0.f && p.out_slope <= 1.f && p.res_inv_slope >= 1.f))
        return CMTTS_ERR_UNSUPPORTED;
#define RB_DISPATCH(C_, MODE_)                                            \
    switch (p.taps) {                                                     \
        case 3: return p.sum_h ? launch_rb_cfg<C_, 3, true, MODE_>(p, s) : launch_rb_cfg<C_, 3, false, MODE_>(p, s);      \
        case 7: return p.sum_h ? launch_rb_cfg<C_, 7, true, MODE_>(p, s) : launch_rb_cfg<C_, 7, false, MODE_>(p, s);      \
        case 11: return p.sum_h ? launch_rb_cfg<C_, 11, true, MODE_>(p, s) : launch_rb_cfg<C_, 11, false, MODE_>(p, s);   \
        default: return CMTTS_ERR_UNSUPPORTED;                            \
    }
    if (p.C == 32) {
        // paired rows (two time steps per 128-byte row): needs the pair-packed weights, an even length, an odd dilation
        static int pair_env = -1;                              // CMTTS_RB_PAIR=0 or CMTTS_UMMA_DBG bit 1024: the plain C = 32 kernel (A/B)
        if (pair_env < 0) { const char* e = getenv("CMTTS_RB_PAIR"); pair_env = e ? atoi(e) : 1; }
        if (g_cmtts_umma_dbg < 0) { const char* e = getenv("CMTTS_UMMA_DBG"); g_cmtts_umma_dbg = e ? atoi(e) : 0; }
        if (pair_env && !(g_cmtts_umma_dbg & 1024) && p.w1p && p.w2p && (p.L % 2) == 0 && (p.dil % 2) == 1 && ((uintptr_t)p.w1p % 16) == 0 && ((uintptr_t)p.w2p % 16) == 0) {
            if (p.dil == 1) { RB_DISPATCH(64, 1) }
            RB_DISPATCH(64, 2)
        }
        RB_DISPATCH(32, 0)
    }
    if (p.C == 64) { RB_DISPATCH(64, 0) }
#undef RB_DISPATCH
    return CMTTS_ERR_UNSUPPORTED;
}
