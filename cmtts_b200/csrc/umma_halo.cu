// tcgen05 conv1d for the HiFi-GAN ResBlock convolutions (Cin == Cout == C in {32, 64, 128}),
// restructured around the things that bound the plain implicit-GEMM kernel (umma_conv.cu) on these
// shapes — L2->SM operand traffic and memory latency with too few bytes in flight:
//
// * HALO TILE: the activation tile is fetched ONCE per output tile as (128 + span) rows
//   (span = (k-1)*dilation <= 64) and every tap's A operand is the same shared-memory tile addressed
//   at a row offset — the UMMA descriptor start address moves by (shift_tap - shift_0) * row_bytes.
//   The 128B/64B swizzle is a function of the absolute shared-memory address bits, identical for
//   the TMA write and the UMMA read, so a row offset needs no re-layout (verified on B200).  This
//   cuts the A traffic by the tap count (3..11x).
// * RESIDENT WEIGHTS: all taps of the layer's weights (<= 96 KB) are loaded into shared memory
//   once per CTA and reused for every tile of the persistent loop; when they do not fit (C = 128,
//   k = 7/11) they stream through their own mbarrier ring.
// * DEEP RING: as many activation stages as shared memory allows (up to 8 tiles in flight per SM); the
//   residual rows of the epilogue are read from global memory (L2) into registers before the
//   accumulator wait.
//
// * WIDE EPILOGUE: 8 epilogue warps (lane quarter x column half).  A single warp per SM sub-partition
//   issues its ~12 instructions per output element back to back at the ALU dependency latency, which
//   made the 4-warp epilogue (not HBM, not the tensor pipe) the limiter of every shape here (ncu:
//   epilogue warps >80 % busy, MMA warp waiting on the accumulator-empty barrier).  Each warp owns its
//   staging slab and TMA store (column half = one channel block, or half of one), the bias sits in
//   shared memory, and leaky-ReLU is max(v, s v) / min(v, s v).
//
// Roles: warp 0 = activation TMA producer (+ resident weights), warp 3 = streamed-weight
// TMA producer, warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator (+ second issuer when the weights
// are resident), warps 4-11 = epilogue.
#include "umma_common.cuh"
#include <stdlib.h>

namespace {

using namespace umma;

constexpr int MAX_STAGES = 8;

struct HaloCfg {
    int a_stages, w_stages;             // ring depths (w_stages = 0: weights resident)
    int rows_alloc;                     // rows reserved per activation block (>= 128 + span)
    int box_rows;                       // 128 + span
    int pf;                             // L2 prefetch distance in tiles (0 = off; CMTTS_PF)
};

// leaky-ReLU without a compare/select: slope <= 1 (forward) and slope >= 1 (the exact inverse on stored values)
__device__ __forceinline__ float lrelu_fwd(float v, float slope) { return fmaxf(v, v * slope); }
__device__ __forceinline__ float lrelu_inv(float v, float inv_slope) { return fminf(v, v * inv_slope); }

template <int BN, int BK, int CB, int TAPS>
__global__ void __launch_bounds__(384, 1)
umma_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmO,
                 const UmmaConvParams p, const HaloCfg cfg) {
    constexpr int BM = 128;
    constexpr int ROW_BYTES = BK * 2;
    constexpr int W_BLK = BN * ROW_BYTES;
    constexpr int BNH = BN / 2;                    // output columns per epilogue warp
    constexpr int OROW = BNH * 2;                  // bytes per staged output row of one warp (128 / 64 / 32)
    constexpr int O_SLAB = 32 * OROW;              // one epilogue warp's 32 rows x BNH columns
    constexpr int O_BYTES = 8 * O_SLAB;            // whole 128 x BN fp16 output tile
    constexpr int TMEM_COLS = pow2_cols(2 * BN);
    constexpr bool TMA_STORE = BN >= 64;           // C = 32: a warp's 32 rows are one contiguous 2 KB run, direct stores win

    const int a_alloc = cfg.rows_alloc * ROW_BYTES;
    const int a_stage = CB * a_alloc;
    const bool wres = cfg.w_stages == 0;
    const bool has_res = p.res_h != nullptr;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* smA = smem;
    uint8_t* smO = smA + cfg.a_stages * a_stage;
    uint8_t* smW = smO + O_BYTES;
    const int w_blocks = wres ? p.taps * CB : cfg.w_stages;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smW + (size_t)w_blocks * W_BLK);
    uint64_t* a_empty = a_full + MAX_STAGES;
    uint64_t* w_full = a_empty + MAX_STAGES;
    uint64_t* w_empty = w_full + MAX_STAGES;
    uint64_t* tfull = w_empty + MAX_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);   // float4 reads

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: warp index provably uniform for ptxas
    const int m_tiles = (p.M + BM - 1) / BM;
    const int tiles = p.B * m_tiles;
    const int shift0 = p.shift[0];

    if (threadIdx.x == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1);
            mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 128 && threadIdx.x < 128 + BN) s_bias[threadIdx.x - 128] = p.bias[threadIdx.x - 128];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    // PDL: weights are constants — warp 0 issues the resident weights before waiting for the previous kernel, the
    // streamed-weight producer (warp 3) never waits; every other role touches activations and waits here
    if (warp != 0 && warp != 3) pdl_wait();

    if (warp == 0) {
        // ======================= activation producer (+ resident weights) =======================
        {
            prefetch_tmap(&tmA); prefetch_tmap(&tmW);
            if (wres) {
                mbar_expect_tx_elect(&w_full[0], (uint32_t)(p.taps * CB * W_BLK));
                for (int tap = 0; tap < p.taps; ++tap)
                    for (int cb = 0; cb < CB; ++cb)
                        tma_load_2d_elect(smW + (size_t)(tap * CB + cb) * W_BLK, &tmW, &w_full[0], cb * BK, tap * p.N);
            }
            pdl_wait();
            int stage = 0; uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)(CB * cfg.box_rows * ROW_BYTES);
            // optional L2 prefetch of the halo tiles PF tiles ahead of the ring (CMTTS_PF, off by default: measured no gain)
            const int PF = cfg.pf;
            auto prefetch_tile = [&](int tl) {
                if (PF > 0 && tl < tiles)
                    for (int cb = 0; cb < CB; ++cb) tma_prefetch_3d_elect(&tmA, cb * BK, (tl % m_tiles) * BM + shift0, tl / m_tiles);
            };
            for (int i = 0; i < PF; ++i) prefetch_tile(blockIdx.x + i * gridDim.x);
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int mt = tile % m_tiles, b = tile / m_tiles;
                prefetch_tile(tile + PF * gridDim.x);
                mbar_wait(&a_empty[stage], phase ^ 1);
                mbar_expect_tx_elect(&a_full[stage], bytes);
                for (int cb = 0; cb < CB; ++cb)
                    tma_load_3d_elect(smA + stage * a_stage + cb * a_alloc, &tmA, &a_full[stage], cb * BK, mt * BM + shift0, b);
                if (++stage == cfg.a_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ======================= streamed-weight producer =======================
        if (!wres) {
            int ws = 0; uint32_t wphase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                for (int tap = 0; tap < p.taps; ++tap)
                    for (int cb = 0; cb < CB; ++cb) {
                        mbar_wait(&w_empty[ws], wphase ^ 1);
                        mbar_expect_tx_elect(&w_full[ws], W_BLK);
                        tma_load_2d_elect(smW + ws * W_BLK, &tmW, &w_full[ws], cb * BK, tap * p.N);
                        if (++ws == cfg.w_stages) { ws = 0; wphase ^= 1; }
                    }
            }
        }
    } else if (warp == 1 || (warp == 2 && wres)) {
        // ======================= MMA issuer(s) =======================
        // With resident weights TWO warps issue, on alternating tiles (warp 1: even, warp 2: odd; the accumulator
        // buffer is the tile parity), a split that dates from the predicated issue form (13-17 SASS instructions per
        // tcgen05.mma, more than a 128 x 64 x 16 MMA occupies the tensor pipe); it is kept with the elect-branch form
        // because the two warps also overlap their barrier waits.  With streamed weights the ring is consumed in tile
        // order, so one warp issues (those shapes are bound by the L2 -> SM weight traffic anyway).
        {
            const int first = wres ? warp - 1 : 0, step = wres ? 2 : 1;
            const uint32_t tmem_u = make_uniform(tmem_base);
            // With the tap count a template parameter the loops unroll completely and all descriptor offsets fold
            // into immediates / one UIADD3 each.
            const uint32_t idesc = make_idesc(BM, BN);
            constexpr uint32_t DESC_HI = (uint32_t)((8 * ROW_BYTES) >> 4) | (1u << 14) | ((BK == 64 ? 2u : 4u) << 29);
            const uint32_t tap_step = (uint32_t)(((p.taps > 1 ? p.shift[1] - p.shift[0] : 0) * ROW_BYTES) >> 4);
            const uint32_t cb_step = (uint32_t)(a_alloc >> 4);
            // Issue form: `if (elect_one())` blocks whose operands all derive from warp-uniform values (make_uniform()'d
            // bases, scalar ring counters) -> bare UTCHMMA runs with descriptors in uniform registers (umma_common.cuh).
            constexpr uint64_t HI = (uint64_t)DESC_HI << 32;
            const uint32_t w_lo0 = make_uniform(((smem_u32(smW) >> 4) & 0x3FFF) | (1u << 16));
            const uint32_t a_base = make_uniform(((smem_u32(smA) >> 4) & 0x3FFF) | (1u << 16));
            const uint32_t a_stage16 = make_uniform((uint32_t)(a_stage >> 4));
            const uint32_t tap_step_u = make_uniform(tap_step), cb_step_u = make_uniform(cb_step);
            const int ntaps = TAPS > 0 ? TAPS : p.taps;
            int stage = first % cfg.a_stages; uint32_t phase = (uint32_t)((first / cfg.a_stages) & 1);
            int ws = 0; uint32_t wphase = 0;
            if (wres) { mbar_wait(&w_full[0], 0); tc_fence_after(); }
            int it = first;
            for (int tile = blockIdx.x + first * gridDim.x; tile < tiles; tile += step * gridDim.x, it += step) {
                const int abuf = it & 1; const uint32_t aphase = (uint32_t)((it >> 1) & 1);
                mbar_wait(&tempty[abuf], aphase ^ 1);
                mbar_wait(&a_full[stage], phase);
                tc_fence_after();
                const uint32_t d_tmem = make_uniform(tmem_u + (uint32_t)(abuf * BN));
                const uint32_t a_lo0 = make_uniform(a_base + (uint32_t)stage * a_stage16);
                if (wres) {
                    if (elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < ntaps; ++tap) {
#pragma unroll
                            for (int cb = 0; cb < CB; ++cb) {
                                const uint32_t a_lo = a_lo0 + tap * tap_step_u + cb * cb_step_u;
                                const uint32_t w_lo = w_lo0 + (uint32_t)(((tap * CB + cb) * W_BLK) >> 4);
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    umma_f16(d_tmem, HI | (a_lo + 2 * k), HI | (w_lo + 2 * k), idesc, (tap | cb | k) ? 1u : 0u);
                            }
                        }
                        umma_commit(&a_empty[stage]);
                        umma_commit(&tfull[abuf]);
                    }
                    __syncwarp();
                } else {
#pragma unroll
                    for (int tap = 0; tap < ntaps; ++tap) {
#pragma unroll
                        for (int cb = 0; cb < CB; ++cb) {
                            mbar_wait(&w_full[ws], wphase);
                            tc_fence_after();
                            const uint32_t a_lo = a_lo0 + tap * tap_step_u + cb * cb_step_u;
                            const uint32_t w_lo = make_uniform(w_lo0 + (uint32_t)((ws * W_BLK) >> 4));
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    umma_f16(d_tmem, HI | (a_lo + 2 * k), HI | (w_lo + 2 * k), idesc, (tap | cb | k) ? 1u : 0u);
                                umma_commit(&w_empty[ws]);
                            }
                            __syncwarp();
                            if (++ws == cfg.w_stages) { ws = 0; wphase ^= 1; }
                        }
                    }
                    umma_commit_pred(&a_empty[stage], 0u);
                    umma_commit_pred(&tfull[abuf], 0u);
                }
                stage += step;
                if (stage >= cfg.a_stages) { stage -= cfg.a_stages; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ======================= epilogue (UEPI_VOC), 8 warps =======================
        const int q = warp & 3;                    // TMEM lane quarter
        const int h = (warp - 4) >> 2;             // column half
        const int row = q * 32 + lane;
        // staging slab of this warp: rows of OROW bytes in the TMA box layout (128B / 64B swizzle by row pitch)
        const int swz_o = (OROW == 128) ? (lane & 7) : ((lane >> 1) & 3);
        uint8_t* slab = smO + (warp - 4) * O_SLAB + lane * OROW;
        const int n_base = h * BNH;
        int abuf = 0; uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int mt = tile % m_tiles, b = tile / m_tiles;
            const int t = mt * BM + row;
            const bool valid = t < p.M;
            // residual lrelu(y) of this thread's row segment: global reads issued before the accumulator wait (the tensor
            // was written by the previous launch, most of it is still in L2)
            {
                // pull the NEXT tile's residual / partial-sum segment of this thread into L2 (one 64-128 byte line each):
                // by the time the register loads below are issued for that tile they no longer pay a DRAM round trip
                const int tn = tile + gridDim.x;
                if (tn < tiles) {
                    const int t2 = (tn % m_tiles) * BM + row, b2 = tn / m_tiles;
                    if (t2 < p.M) {
                        if (has_res) prefetch_l2(p.res_h + (long long)b2 * p.res_bstride + (long long)t2 * p.res_ld + n_base);
                        if (p.sum_h) prefetch_l2(p.sum_h + (long long)b2 * p.out_bstride + (long long)t2 * p.out_ld + n_base);
                    }
                }
            }
            uint4 rres[BNH / 8];
            if (has_res) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.res_h + (long long)b * p.res_bstride + (long long)t * p.res_ld + n_base);
#pragma unroll
                for (int i = 0; i < BNH / 8; ++i) rres[i] = valid ? rp[i] : make_uint4(0u, 0u, 0u, 0u);
            }
            // MRF partial sum (last iteration of a resblock): global reads issued before the accumulator wait
            uint4 rsum[BNH / 8];
            const bool has_sum = p.sum_h != nullptr && valid;
            if (has_sum) {
                const uint4* sp = reinterpret_cast<const uint4*>(p.sum_h + (long long)b * p.out_bstride + (long long)t * p.out_ld + n_base);
#pragma unroll
                for (int i = 0; i < BNH / 8; ++i) rsum[i] = sp[i];
            }
            mbar_wait(&tfull[abuf], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(abuf * BN + n_base) + ((uint32_t)(q * 32) << 16);
            // the previous tile's TMA store must have finished READING this warp's staging slab
            if (TMA_STORE) {
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
            }
#pragma unroll
            for (int c = 0; c < BNH / 16; ++c) {
                uint32_t r[16];
                tmem_ld16(taddr + c * 16, r);
                tmem_ld_wait();
                const int n = n_base + c * 16;
                float v[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bq = *reinterpret_cast<const float4*>(s_bias + n + 4 * j);
                    v[4 * j] = fmaf(__uint_as_float(r[4 * j]), p.alpha, bq.x);
                    v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), p.alpha, bq.y);
                    v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), p.alpha, bq.z);
                    v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), p.alpha, bq.w);
                }
                if (has_res) {
                    const __half2* h0 = reinterpret_cast<const __half2*>(&rres[2 * c]);
                    const __half2* h1 = reinterpret_cast<const __half2*>(&rres[2 * c + 1]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 a = __half22float2(h0[i]), bb = __half22float2(h1[i]);
                        v[2 * i] += lrelu_inv(a.x, p.res_inv_slope); v[2 * i + 1] += lrelu_inv(a.y, p.res_inv_slope);
                        v[8 + 2 * i] += lrelu_inv(bb.x, p.res_inv_slope); v[8 + 2 * i + 1] += lrelu_inv(bb.y, p.res_inv_slope);
                    }
                }
                if (has_sum) {
                    const __half2* s0 = reinterpret_cast<const __half2*>(&rsum[2 * c]);
                    const __half2* s1 = reinterpret_cast<const __half2*>(&rsum[2 * c + 1]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 a = __half22float2(s0[i]), bb = __half22float2(s1[i]);
                        v[2 * i] += a.x; v[2 * i + 1] += a.y; v[8 + 2 * i] += bb.x; v[8 + 2 * i + 1] += bb.y;
                    }
                }
                // stage as fp16 in the TMA box layout: 16-byte chunk j of a row lives at chunk (j ^ swz)
                uint4 u0, u1;
                __half2* p0 = reinterpret_cast<__half2*>(&u0);
                __half2* p1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    p0[i] = __floats2half2_rn(lrelu_fwd(v[2 * i], p.out_slope), lrelu_fwd(v[2 * i + 1], p.out_slope));
                    p1[i] = __floats2half2_rn(lrelu_fwd(v[8 + 2 * i], p.out_slope), lrelu_fwd(v[8 + 2 * i + 1], p.out_slope));
                }
                if (TMA_STORE) {
                    *reinterpret_cast<uint4*>(slab + (((2 * c) ^ swz_o) << 4)) = u0;
                    *reinterpret_cast<uint4*>(slab + (((2 * c + 1) ^ swz_o) << 4)) = u1;
                } else if (valid) {
                    __half* op = p.out_h + (long long)b * p.out_bstride + (long long)t * p.out_ld + n;
                    *reinterpret_cast<uint4*>(op) = u0;
                    *reinterpret_cast<uint4*>(op + 8) = u1;
                }
            }
            tc_fence_before();
            if (TMA_STORE) fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&tempty[abuf]);
                if (TMA_STORE) {
                    tma_store_3d(&tmO, smO + (warp - 4) * O_SLAB, n_base, mt * BM + q * 32, b);
                    tma_store_commit();
                }
            }
            abuf ^= 1; if (abuf == 0) aphase ^= 1;
        }
        if (TMA_STORE && lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int BN, int BK, int CB, int TAPS>
int launch_halo_cfg(const UmmaConvParams& p, cudaStream_t s) {
    constexpr int ROW_BYTES = BK * 2;
    constexpr int W_BLK = BN * ROW_BYTES;
    constexpr int ROW_ALIGN = 1024 / ROW_BYTES;        // rows per 1024-byte swizzle-aligned unit
    constexpr size_t LIMIT = 227 * 1024;
    constexpr size_t O_BYTES = (size_t)8 * 32 * BN;        // 8 warps x 32 rows x BN/2 fp16
    constexpr size_t FIXED = (4 * MAX_STAGES + 4) * 8 + 32 + BN * 4 + 1024 + O_BYTES;

    HaloCfg cfg{};
    const int span = p.shift[p.taps - 1] - p.shift[0];
    cfg.box_rows = 128 + span;
    cfg.rows_alloc = (cfg.box_rows + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    const size_t a_stage = (size_t)CB * cfg.rows_alloc * ROW_BYTES;
    const size_t w_res = (size_t)p.taps * CB * W_BLK;
    // weights stay resident when that leaves room for >= 3 activation stages
    size_t budget = LIMIT - FIXED;
    size_t w_bytes;
    if (w_res + 3 * a_stage <= budget) { cfg.w_stages = 0; w_bytes = w_res; }
    else { cfg.w_stages = 4; w_bytes = (size_t)cfg.w_stages * W_BLK; }
    if (w_bytes + 2 * a_stage > budget) return CMTTS_ERR_UNSUPPORTED;
    budget -= w_bytes;
    size_t a = budget / a_stage;
    cfg.a_stages = (int)(a > MAX_STAGES ? MAX_STAGES : a);
    if (cfg.a_stages < 2) return CMTTS_ERR_UNSUPPORTED;
    const size_t smem = (size_t)cfg.a_stages * a_stage + w_bytes + FIXED;

    static int pf_env = -1;
    if (pf_env < 0) { const char* e = getenv("CMTTS_PF"); pf_env = e ? atoi(e) : 0; }
    cfg.pf = pf_env;
    auto kern = umma_halo_kernel<BN, BK, CB, TAPS>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIMIT) != cudaSuccess) {
            cmtts_set_error("umma_halo: cannot set dynamic shared memory size", __FILE__, __LINE__);
            return CMTTS_ERR_CUDA;
        }
        attr_done = true;
    }
    CUtensorMap a_map, w_map, o_map;
    if (!make_act_map(&a_map, p.a_hi, p.Cin, p.Lin, p.B, p.a_ld, p.a_bstride, BK, cfg.box_rows) ||
        !make_w_map(&w_map, p.w_hi, p.Cin, p.taps * p.N, BK, BN)) {
        cmtts_set_error("umma_halo: cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    // output boxes: one per epilogue warp, BN/2 channels x 32 rows (128B swizzle for 64 channels, 64B for 32)
    if (!make_act_map(&o_map, p.out_h, p.N, p.M, p.B, p.out_ld, p.out_bstride, BN >= 128 ? 64 : 32, 32)) {
        cmtts_set_error("umma_halo: cuTensorMapEncodeTiled failed (output)", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    const int tiles = p.B * ((p.M + 127) / 128);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    if (g_cmtts_prof_on) {
        const double rows = (double)p.B * p.M;
        char lbl[96];
        snprintf(lbl, sizeof(lbl), "umma_halo<%d> k%d d%d%s%s", p.N, p.taps, p.taps > 1 ? p.shift[1] - p.shift[0] : 1,
                 p.res_h ? " +res" : "", p.sum_h ? " +sum" : "");
        cmtts_prof_note(lbl, 2.0 * rows * p.N * p.taps * p.Cin,
                        rows * (p.Cin + p.N) * 2.0 + (p.res_h ? rows * p.N * 2.0 : 0.0) + (p.sum_h ? rows * p.N * 2.0 : 0.0) +
                            (double)p.taps * p.N * p.Cin * 2.0);
    }
    launch_pdl(kern, grid, 384, smem, s, a_map, w_map, o_map, p, cfg);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

}  // namespace

// Returns CMTTS_ERR_UNSUPPORTED (without setting an error) when the shape is not covered, so the
// caller can use the general kernel.
int launch_umma_halo(const UmmaConvParams& p, cudaStream_t s) {
    if (p.split || p.epi != UEPI_VOC || p.Cin != p.N || p.taps < 1 || !p.bias) return CMTTS_ERR_UNSUPPORTED;
    for (int i = 1; i < p.taps; ++i)
        if (p.shift[i] <= p.shift[i - 1]) return CMTTS_ERR_UNSUPPORTED;
    if (p.shift[p.taps - 1] - p.shift[0] > 96) return CMTTS_ERR_UNSUPPORTED;
    if (p.out_ld % 8 != 0 || p.out_bstride % 8 != 0 || ((uintptr_t)p.out_h % 16) != 0) return CMTTS_ERR_UNSUPPORTED;
    if (p.res_h && (p.res_ld % 8 != 0 || p.res_bstride % 8 != 0 || ((uintptr_t)p.res_h % 16) != 0)) return CMTTS_ERR_UNSUPPORTED;
    // max / min form of leaky-ReLU: forward slope in (0, 1], inverse slope >= 1
    if (!(p.out_slope > 0.f && p.out_slope <= 1.f && p.res_inv_slope >= 1.f)) return CMTTS_ERR_UNSUPPORTED;
    if (p.B == 0 || p.M == 0) return CMTTS_OK;
    for (int i = 2; i < p.taps; ++i)   // uniform tap spacing (dilation)
        if (p.shift[i] - p.shift[i - 1] != p.shift[1] - p.shift[0]) return CMTTS_ERR_UNSUPPORTED;
#define HALO_DISPATCH(BN_, BK_, CB_)                                              \
    switch (p.taps) {                                                             \
        case 3: return launch_halo_cfg<BN_, BK_, CB_, 3>(p, s);                   \
        case 7: return launch_halo_cfg<BN_, BK_, CB_, 7>(p, s);                   \
        case 11: return launch_halo_cfg<BN_, BK_, CB_, 11>(p, s);                 \
        default: return launch_halo_cfg<BN_, BK_, CB_, 0>(p, s);                  \
    }
    if (p.N == 128 || p.N == 256) {   // streamed weights: the CTA-pair kernel halves the weight bytes each SM takes in
        const int rc = launch_umma_halo2(p, s);
        if (rc != CMTTS_ERR_UNSUPPORTED) return rc;
    }
    switch (p.N) {
        case 32: HALO_DISPATCH(32, 32, 1)
        case 64: HALO_DISPATCH(64, 64, 1)
        case 128: HALO_DISPATCH(128, 64, 2)
        default: return CMTTS_ERR_UNSUPPORTED;
    }
#undef HALO_DISPATCH
}
