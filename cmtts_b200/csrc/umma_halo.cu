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
// * DEEP RINGS: as many activation stages as shared memory allows (up to 8 tiles in flight per SM)
//   and the residual tile of the epilogue is ALSO fetched by TMA into its own ring, so the
//   HBM-bound levels keep tens of KB in flight per SM instead of one row per thread.
//
// Roles: warp 0 = activation TMA producer (+ resident weights), warp 3 = residual + streamed-weight
// TMA producer, warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue.
#include "umma_common.cuh"

namespace {

using namespace umma;

constexpr int MAX_STAGES = 8;

struct HaloCfg {
    int a_stages, r_stages, w_stages;   // ring depths (w_stages = 0: weights resident)
    int rows_alloc;                     // rows reserved per activation block (>= 128 + span)
    int box_rows;                       // 128 + span
};

template <int BN, int BK, int CB, int TAPS>
__global__ void __launch_bounds__(256, 1)
umma_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                 const UmmaConvParams p, const HaloCfg cfg) {
    constexpr int BM = 128;
    constexpr int ROW_BYTES = BK * 2;
    constexpr int W_BLK = BN * ROW_BYTES;
    constexpr int R_BLK = BM * ROW_BYTES;          // one channel block of the residual tile
    constexpr int R_STAGE = CB * R_BLK;
    constexpr int O_SLAB = 32 * ROW_BYTES;         // one epilogue warp's 32 output rows of one channel block
    constexpr int O_BYTES = 4 * CB * O_SLAB;       // whole 128 x BN fp16 output tile
    constexpr int TMEM_COLS = pow2_cols(2 * BN);
    constexpr bool TMA_STORE = BN >= 64;           // C = 32: a warp's 32 rows are one contiguous 2 KB run, direct stores win

    const int a_alloc = cfg.rows_alloc * ROW_BYTES;
    const int a_stage = CB * a_alloc;
    const bool wres = cfg.w_stages == 0;
    const bool has_res = p.res_h != nullptr;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smA = smem;
    uint8_t* smR = smA + cfg.a_stages * a_stage;
    uint8_t* smO = smR + cfg.r_stages * R_STAGE;
    uint8_t* smW = smO + O_BYTES;
    const int w_blocks = wres ? p.taps * CB : cfg.w_stages;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smW + (size_t)w_blocks * W_BLK);
    uint64_t* a_empty = a_full + MAX_STAGES;
    uint64_t* w_full = a_empty + MAX_STAGES;
    uint64_t* w_empty = w_full + MAX_STAGES;
    uint64_t* r_full = w_empty + MAX_STAGES;
    uint64_t* r_empty = r_full + MAX_STAGES;
    uint64_t* tfull = r_empty + MAX_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM;
    const int tiles = p.B * m_tiles;
    const int shift0 = p.shift[0];

    if (threadIdx.x == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1);
            mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
            mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 4);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
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
            int stage = 0; uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)(CB * cfg.box_rows * ROW_BYTES);
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int mt = tile % m_tiles, b = tile / m_tiles;
                mbar_wait(&a_empty[stage], phase ^ 1);
                mbar_expect_tx_elect(&a_full[stage], bytes);
                for (int cb = 0; cb < CB; ++cb)
                    tma_load_3d_elect(smA + stage * a_stage + cb * a_alloc, &tmA, &a_full[stage], cb * BK, mt * BM + shift0, b);
                if (++stage == cfg.a_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ======================= residual + streamed-weight producer =======================
        if (has_res || !wres) {
            if (has_res) prefetch_tmap(&tmR);
            int ws = 0; uint32_t wphase = 0;
            int rs = 0; uint32_t rphase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int mt = tile % m_tiles, b = tile / m_tiles;
                if (has_res) {
                    mbar_wait(&r_empty[rs], rphase ^ 1);
                    mbar_expect_tx_elect(&r_full[rs], R_STAGE);
                    for (int cb = 0; cb < CB; ++cb)
                        tma_load_3d_elect(smR + rs * R_STAGE + cb * R_BLK, &tmR, &r_full[rs], cb * BK, mt * BM, b);
                    if (++rs == cfg.r_stages) { rs = 0; rphase ^= 1; }
                }
                if (!wres) {
                    for (int tap = 0; tap < p.taps; ++tap)
                        for (int cb = 0; cb < CB; ++cb) {
                            mbar_wait(&w_empty[ws], wphase ^ 1);
                            mbar_expect_tx_elect(&w_full[ws], W_BLK);
                            tma_load_2d_elect(smW + ws * W_BLK, &tmW, &w_full[ws], cb * BK, tap * p.N);
                            if (++ws == cfg.w_stages) { ws = 0; wphase ^= 1; }
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        {
            const uint32_t tmem_u = make_uniform(tmem_base);
            // The issuing thread is a single lane: every integer instruction on its path delays the next
            // tcgen05.mma.  With the tap count a template parameter the loops unroll completely, all
            // descriptor offsets fold into immediates / one IADD each, and the tensor pipe stays fed.
            const uint32_t idesc = make_idesc(BM, BN);
            constexpr uint32_t DESC_HI = (uint32_t)((8 * ROW_BYTES) >> 4) | (1u << 14) | ((BK == 64 ? 2u : 4u) << 29);
            const uint32_t tap_step = (uint32_t)(((p.taps > 1 ? p.shift[1] - p.shift[0] : 0) * ROW_BYTES) >> 4);
            const uint32_t cb_step = (uint32_t)(a_alloc >> 4);
            const uint32_t w_lo0 = ((smem_u32(smW) >> 4) & 0x3FFF) | (1u << 16);
            const int ntaps = TAPS > 0 ? TAPS : p.taps;
            int stage = 0; uint32_t phase = 0;
            int ws = 0; uint32_t wphase = 0;
            int abuf = 0; uint32_t aphase = 0;
            if (wres) { mbar_wait(&w_full[0], 0); tc_fence_after(); }
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                mbar_wait(&tempty[abuf], aphase ^ 1);
                mbar_wait(&a_full[stage], phase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_u + (uint32_t)(abuf * BN);
                const uint32_t a_lo0 = ((smem_u32(smA + stage * a_stage) >> 4) & 0x3FFF) | (1u << 16);
                if (wres) {
#pragma unroll
                    for (int tap = 0; tap < ntaps; ++tap) {
#pragma unroll
                        for (int cb = 0; cb < CB; ++cb) {
                            const uint32_t a_lo = a_lo0 + tap * tap_step + cb * cb_step;
                            const uint32_t w_lo = w_lo0 + (uint32_t)(((tap * CB + cb) * W_BLK) >> 4);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_f16_pred(d_tmem, ((uint64_t)DESC_HI << 32) | (a_lo + 2 * k), ((uint64_t)DESC_HI << 32) | (w_lo + 2 * k),
                                         idesc, (tap | cb | k) ? 1u : 0u, 0u);
                        }
                    }
                } else {
#pragma unroll
                    for (int tap = 0; tap < ntaps; ++tap) {
#pragma unroll
                        for (int cb = 0; cb < CB; ++cb) {
                            mbar_wait(&w_full[ws], wphase);
                            tc_fence_after();
                            const uint32_t a_lo = a_lo0 + tap * tap_step + cb * cb_step;
                            const uint32_t w_lo = w_lo0 + (uint32_t)((ws * W_BLK) >> 4);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_f16_pred(d_tmem, ((uint64_t)DESC_HI << 32) | (a_lo + 2 * k), ((uint64_t)DESC_HI << 32) | (w_lo + 2 * k),
                                         idesc, (tap | cb | k) ? 1u : 0u, 0u);
                            umma_commit_pred(&w_empty[ws], 0u);
                            if (++ws == cfg.w_stages) { ws = 0; wphase ^= 1; }
                        }
                    }
                }
                umma_commit_pred(&a_empty[stage], 0u);
                umma_commit_pred(&tfull[abuf], 0u);
                if (++stage == cfg.a_stages) { stage = 0; phase ^= 1; }
                abuf ^= 1; if (abuf == 0) aphase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ======================= epilogue (UEPI_VOC) =======================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        // 16-byte chunk swizzle of the TMA-written residual tile (same pattern as the operands)
        const int swz = (BK == 64) ? (row & 7) : ((row >> 1) & 3);
        constexpr int CHUNKS = ROW_BYTES / 16;     // 16-byte chunks per channel-block row: 8 or 4
        const int swz_o = (BK == 64) ? (lane & 7) : ((lane >> 1) & 3);   // same pattern, row index inside the warp slab
        int abuf = 0; uint32_t aphase = 0;
        int rs = 0; uint32_t rphase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int mt = tile % m_tiles, b = tile / m_tiles;
            const int t = mt * BM + row;
            const bool valid = t < p.M;
            uint4 rres[BN / 8];
            if (has_res) {
                mbar_wait(&r_full[rs], rphase);
                const uint8_t* rb = smR + rs * R_STAGE + row * ROW_BYTES;
#pragma unroll
                for (int i = 0; i < BN / 8; ++i) {
                    const int cb = i / CHUNKS, j = i % CHUNKS;
                    rres[i] = *reinterpret_cast<const uint4*>(rb + cb * R_BLK + ((j ^ swz) << 4));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&r_empty[rs]);
                if (++rs == cfg.r_stages) { rs = 0; rphase ^= 1; }
            }
            mbar_wait(&tfull[abuf], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(abuf * BN) + ((uint32_t)(q * 32) << 16);
            // the previous tile's TMA store must have finished READING this warp's staging slab
            if (TMA_STORE) {
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
            }
            uint8_t* slab = smO + q * (CB * O_SLAB) + lane * ROW_BYTES;     // this thread's row inside the warp slab
#pragma unroll
            for (int c = 0; c < BN / 16; ++c) {
                uint32_t r[16];
                tmem_ld16(taddr + c * 16, r);
                tmem_ld_wait();
                const int n = c * 16;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(r[j]), p.alpha, p.bias[n + j]);
                if (has_res) {
                    const __half2* h0 = reinterpret_cast<const __half2*>(&rres[2 * c]);
                    const __half2* h1 = reinterpret_cast<const __half2*>(&rres[2 * c + 1]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 a = __half22float2(h0[i]), bb = __half22float2(h1[i]);
                        v[2 * i] += lrelu(a.x, p.res_inv_slope); v[2 * i + 1] += lrelu(a.y, p.res_inv_slope);
                        v[8 + 2 * i] += lrelu(bb.x, p.res_inv_slope); v[8 + 2 * i + 1] += lrelu(bb.y, p.res_inv_slope);
                    }
                }
                if (p.sum_h && valid) {
                    float ss[16];
                    load16h(p.sum_h + (long long)b * p.out_bstride + (long long)t * p.out_ld + n, ss);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += ss[j];
                }
                // stage as fp16 in the TMA box layout: 16-byte chunk j of a row lives at chunk (j ^ swz)
                uint4 u0, u1;
                __half2* p0 = reinterpret_cast<__half2*>(&u0);
                __half2* p1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    p0[i] = __floats2half2_rn(lrelu(v[2 * i], p.out_slope), lrelu(v[2 * i + 1], p.out_slope));
                    p1[i] = __floats2half2_rn(lrelu(v[8 + 2 * i], p.out_slope), lrelu(v[8 + 2 * i + 1], p.out_slope));
                }
                if (TMA_STORE) {
                    const int cb = (2 * c) / CHUNKS, j0 = (2 * c) % CHUNKS;
                    *reinterpret_cast<uint4*>(slab + cb * O_SLAB + ((j0 ^ swz_o) << 4)) = u0;
                    *reinterpret_cast<uint4*>(slab + cb * O_SLAB + (((j0 + 1) ^ swz_o) << 4)) = u1;
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
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb)
                        tma_store_3d(&tmO, smO + q * (CB * O_SLAB) + cb * O_SLAB, cb * BK, mt * BM + q * 32, b);
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
    constexpr int R_STAGE = CB * 128 * ROW_BYTES;
    constexpr int ROW_ALIGN = 1024 / ROW_BYTES;        // rows per 1024-byte swizzle-aligned unit
    constexpr size_t LIMIT = 227 * 1024;
    constexpr size_t O_BYTES = (size_t)4 * CB * 32 * ROW_BYTES;
    constexpr size_t FIXED = (6 * MAX_STAGES + 4) * 8 + 16 + 1024 + O_BYTES;

    HaloCfg cfg{};
    const int span = p.shift[p.taps - 1] - p.shift[0];
    cfg.box_rows = 128 + span;
    cfg.rows_alloc = (cfg.box_rows + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    const size_t a_stage = (size_t)CB * cfg.rows_alloc * ROW_BYTES;
    const size_t w_res = (size_t)p.taps * CB * W_BLK;
    const bool has_res = p.res_h != nullptr;
    // weights stay resident when that leaves room for >= 3 activation stages (+ 2 residual stages)
    size_t budget = LIMIT - FIXED;
    const size_t need_min = 3 * a_stage + (has_res ? 2 * R_STAGE : 0);
    size_t w_bytes;
    if (w_res + need_min <= budget) { cfg.w_stages = 0; w_bytes = w_res; }
    else { cfg.w_stages = 4; w_bytes = (size_t)cfg.w_stages * W_BLK; }
    if (w_bytes + 2 * a_stage + (has_res ? R_STAGE : 0) > budget) return CMTTS_ERR_UNSUPPORTED;
    budget -= w_bytes;
    // split the rest: residual ring gets ~1/3 of the bytes in flight when present
    cfg.r_stages = 0;
    if (has_res) {
        size_t r = budget / 3 / R_STAGE;
        cfg.r_stages = (int)(r < 1 ? 1 : (r > MAX_STAGES ? MAX_STAGES : r));
        budget -= (size_t)cfg.r_stages * R_STAGE;
    }
    size_t a = budget / a_stage;
    cfg.a_stages = (int)(a > MAX_STAGES ? MAX_STAGES : a);
    if (cfg.a_stages < 2) return CMTTS_ERR_UNSUPPORTED;
    const size_t smem = (size_t)cfg.a_stages * a_stage + (size_t)cfg.r_stages * R_STAGE + w_bytes + FIXED;

    auto kern = umma_halo_kernel<BN, BK, CB, TAPS>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIMIT) != cudaSuccess) {
            cmtts_set_error("umma_halo: cannot set dynamic shared memory size", __FILE__, __LINE__);
            return CMTTS_ERR_CUDA;
        }
        attr_done = true;
    }
    CUtensorMap a_map, w_map, r_map, o_map;
    if (!make_act_map(&a_map, p.a_hi, p.Cin, p.Lin, p.B, p.a_ld, p.a_bstride, BK, cfg.box_rows) ||
        !make_w_map(&w_map, p.w_hi, p.Cin, p.taps * p.N, BK, BN)) {
        cmtts_set_error("umma_halo: cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    r_map = a_map;
    if (has_res && !make_act_map(&r_map, p.res_h, p.N, p.M, p.B, p.res_ld, p.res_bstride, BK, 128)) {
        cmtts_set_error("umma_halo: cuTensorMapEncodeTiled failed (residual)", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    if (!make_act_map(&o_map, p.out_h, p.N, p.M, p.B, p.out_ld, p.out_bstride, BK, 32)) {
        cmtts_set_error("umma_halo: cuTensorMapEncodeTiled failed (output)", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    const int tiles = p.B * ((p.M + 127) / 128);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    kern<<<grid, 256, smem, s>>>(a_map, w_map, r_map, o_map, p, cfg);
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
    switch (p.N) {
        case 32: HALO_DISPATCH(32, 32, 1)
        case 64: HALO_DISPATCH(64, 64, 1)
        case 128: HALO_DISPATCH(128, 64, 2)
        default: return CMTTS_ERR_UNSUPPORTED;
    }
#undef HALO_DISPATCH
}
