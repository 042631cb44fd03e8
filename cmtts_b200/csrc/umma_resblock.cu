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
//   * the residual y is read back from the input tile already in shared memory (stored as lrelu(y),
//     inverted exactly in the epilogue), the output is staged in the TMA box layout and TMA-stored.
// HBM traffic per iteration: read y once (+halo), write y' once.
//
// Pipelining: the MMA-issuing warp alternates conv1(i+1) / conv2(i); epilogue 1 (TMEM -> t tile) and
// epilogue 2 (output) run on separate warp quartets concurrently; accumulators and the t tile are
// double-buffered.
// Both weight sets stay resident in shared memory for the whole persistent loop.
#include "umma_common.cuh"

namespace {

using namespace umma;

constexpr int RB_MAX_STAGES = 8;
constexpr int T_ROWS_ALLOC = 144;      // 128 + (k_max - 1) rounded to the swizzle atom

struct RbCfg {
    int a_stages, rows_alloc, box_rows, valid, m_tiles;
};

template <int C, int TAPS>
__global__ void __launch_bounds__(384, 1)
umma_resblock_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO,
                     const __grid_constant__ CUtensorMap tmOtail, const UmmaResblockParams p, const RbCfg cfg) {
    constexpr int BM = 128;
    constexpr int BK = C;                          // one channel block (C = 64 -> SW128, C = 32 -> SW64)
    constexpr int ROW_BYTES = BK * 2;
    constexpr int CHUNKS = ROW_BYTES / 16;
    constexpr int W_BLK = C * ROW_BYTES;           // one tap of one conv
    constexpr int T_ALLOC = T_ROWS_ALLOC * ROW_BYTES;
    constexpr int O_SLAB = 32 * ROW_BYTES;
    constexpr int TMEM_COLS = pow2_cols(4 * C);
    constexpr int P2 = (TAPS - 1) / 2;
    constexpr uint32_t DESC_HI = (uint32_t)((8 * ROW_BYTES) >> 4) | (1u << 14) | ((BK == 64 ? 2u : 4u) << 29);

    const int a_alloc = cfg.rows_alloc * ROW_BYTES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smA = smem;
    uint8_t* smT = smA + cfg.a_stages * a_alloc;
    uint8_t* smO = smT + 2 * T_ALLOC;
    uint8_t* smW1 = smO + 4 * O_SLAB;
    uint8_t* smW2 = smW1 + TAPS * W_BLK;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smW2 + TAPS * W_BLK);
    uint64_t* a_empty = a_full + RB_MAX_STAGES;
    uint64_t* w_full = a_empty + RB_MAX_STAGES;    // [1]
    uint64_t* acc1_full = w_full + 1;              // [2]
    uint64_t* acc1_empty = acc1_full + 2;
    uint64_t* acc2_full = acc1_empty + 2;
    uint64_t* acc2_empty = acc2_full + 2;
    uint64_t* t_full = acc2_empty + 2;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = p.B * cfg.m_tiles;
    const int p1d = ((TAPS - 1) / 2) * p.dil;      // conv1 half-span in rows

    if (threadIdx.x == 0) {
        for (int i = 0; i < RB_MAX_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 4); }
        mbar_init(&w_full[0], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], 4);
            mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], 4);
            mbar_init(&t_full[i], 4); mbar_init(&t_empty[i], 1);
        }
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
    // TMEM columns: acc1[b] at b*C, acc2[b] at 2C + b*C

    if (warp == 0) {
        // ======================= producer: resident weights, then one halo tile per output tile =======
        {
            prefetch_tmap(&tmA); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
            mbar_expect_tx_elect(&w_full[0], (uint32_t)(2 * TAPS * W_BLK));
            for (int tap = 0; tap < TAPS; ++tap) {
                tma_load_2d_elect(smW1 + tap * W_BLK, &tmW1, &w_full[0], 0, tap * C);
                tma_load_2d_elect(smW2 + tap * W_BLK, &tmW2, &w_full[0], 0, tap * C);
            }
            int stage = 0; uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)(cfg.box_rows * ROW_BYTES);
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int mt = tile % cfg.m_tiles, b = tile / cfg.m_tiles;
                const int row0 = mt * cfg.valid - P2 - p1d;          // first input row of the halo tile
                mbar_wait(&a_empty[stage], phase ^ 1);
                mbar_expect_tx_elect(&a_full[stage], bytes);
                tma_load_3d_elect(smA + stage * a_alloc, &tmA, &a_full[stage], 0, row0, b);
                if (++stage == cfg.a_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        {
            // the whole warp runs this loop convergently; one lane's tcgen05 instructions are predicated on
            const uint32_t issue = 0;
            const uint32_t tmem_u = make_uniform(tmem_base);
            const uint32_t idesc = make_idesc(BM, C);
            const uint32_t tap_step1 = (uint32_t)((p.dil * ROW_BYTES) >> 4);
            constexpr uint32_t tap_step2 = (uint32_t)(ROW_BYTES >> 4);
            const uint32_t w1_lo = ((smem_u32(smW1) >> 4) & 0x3FFF) | (1u << 16);
            const uint32_t w2_lo = ((smem_u32(smW2) >> 4) & 0x3FFF) | (1u << 16);
            mbar_wait(&w_full[0], 0);
            tc_fence_after();
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            auto conv2 = [&](int j) {
                const int bb = j & 1; const uint32_t ph = (uint32_t)((j >> 1) & 1);
                mbar_wait(&t_full[bb], ph);
                mbar_wait(&acc2_empty[bb], ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_u + (uint32_t)(2 * C + bb * C);
                const uint32_t t_lo = ((smem_u32(smT + bb * T_ALLOC) >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_f16_pred(d_tmem, ((uint64_t)DESC_HI << 32) | (t_lo + tap * tap_step2 + 2 * k),
                                 ((uint64_t)DESC_HI << 32) | (w2_lo + (uint32_t)((tap * W_BLK) >> 4) + 2 * k), idesc,
                                 (tap | k) ? 1u : 0u, issue);
                }
                umma_commit_pred(&t_empty[bb], issue);
                umma_commit_pred(&acc2_full[bb], issue);
            };
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
                const int bb = it & 1; const uint32_t ph = (uint32_t)((it >> 1) & 1);
                mbar_wait(&acc1_empty[bb], ph ^ 1);
                mbar_wait(&a_full[stage], phase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_u + (uint32_t)(bb * C);
                const uint32_t a_lo = ((smem_u32(smA + stage * a_alloc) >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_f16_pred(d_tmem, ((uint64_t)DESC_HI << 32) | (a_lo + tap * tap_step1 + 2 * k),
                                 ((uint64_t)DESC_HI << 32) | (w1_lo + (uint32_t)((tap * W_BLK) >> 4) + 2 * k), idesc,
                                 (tap | k) ? 1u : 0u, issue);
                }
                umma_commit_pred(&acc1_full[bb], issue);
                if (++stage == cfg.a_stages) { stage = 0; phase ^= 1; }
                if (it > 0) conv2(it - 1);
            }
            if (it > 0) conv2(it - 1);
        }
    } else if (warp >= 4) {
        // ======================= epilogue warps =======================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int swz = (BK == 64) ? (row & 7) : ((row >> 1) & 3);         // t tile / staging slab (row-relative)
        const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
        int it = 0;
        int e2_stage = 0;                                                   // A stage of the tile epilogue2 handles next

        auto epilogue2 = [&](int j, int mt, int b) {
            const int bb = j & 1; const uint32_t ph = (uint32_t)((j >> 1) & 1);
            const int o = mt * cfg.valid + row;                             // output row of this thread
            const bool valid = row < cfg.valid && o < p.L;
            // conv2(j) complete  =>  conv1(j) complete  =>  this tile's input stage is loaded and no longer
            // needed by the tensor core; only now may it be read (residual) and handed back to the producer
            mbar_wait(&acc2_full[bb], ph);
            tc_fence_after();
            // residual: lrelu(y) sits in the input halo tile at row (row + P2 + p1d)
            uint4 rres[C / 8];
            {
                const int ra = row + P2 + p1d;
                const int swa = (BK == 64) ? (ra & 7) : ((ra >> 1) & 3);
                const uint8_t* rb = smA + e2_stage * a_alloc + ra * ROW_BYTES;
#pragma unroll
                for (int i = 0; i < C / 8; ++i) rres[i] = *reinterpret_cast<const uint4*>(rb + ((i ^ swa) << 4));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_empty[e2_stage]);
            if (++e2_stage == cfg.a_stages) e2_stage = 0;
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
            const uint32_t taddr = tmem_base + (uint32_t)(2 * C + bb * C) + lane_addr;
            uint8_t* slab = smO + q * O_SLAB + lane * ROW_BYTES;
            const int swo = (BK == 64) ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
            for (int c = 0; c < C / 16; ++c) {
                uint32_t r[16];
                tmem_ld16(taddr + c * 16, r);
                tmem_ld_wait();
                const int n = c * 16;
                float v[16];
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) v[jj] = fmaf(__uint_as_float(r[jj]), p.alpha2, p.b2[n + jj]);
                const __half2* h0 = reinterpret_cast<const __half2*>(&rres[2 * c]);
                const __half2* h1 = reinterpret_cast<const __half2*>(&rres[2 * c + 1]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 a = __half22float2(h0[i]), bq = __half22float2(h1[i]);
                    v[2 * i] += lrelu(a.x, p.res_inv_slope); v[2 * i + 1] += lrelu(a.y, p.res_inv_slope);
                    v[8 + 2 * i] += lrelu(bq.x, p.res_inv_slope); v[8 + 2 * i + 1] += lrelu(bq.y, p.res_inv_slope);
                }
                if (p.sum_h && valid) {
                    float ss[16];
                    load16h(p.sum_h + ((long long)b * p.L + o) * C + n, ss);
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) v[jj] += ss[jj];
                }
                uint4 u0, u1;
                __half2* p0 = reinterpret_cast<__half2*>(&u0);
                __half2* p1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    p0[i] = __floats2half2_rn(lrelu(v[2 * i], p.out_slope), lrelu(v[2 * i + 1], p.out_slope));
                    p1[i] = __floats2half2_rn(lrelu(v[8 + 2 * i], p.out_slope), lrelu(v[8 + 2 * i + 1], p.out_slope));
                }
                *reinterpret_cast<uint4*>(slab + (((2 * c) ^ swo) << 4)) = u0;
                *reinterpret_cast<uint4*>(slab + (((2 * c + 1) ^ swo) << 4)) = u1;
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&acc2_empty[bb]);
                // rows [q*32, q*32+32) of the tile; the last warp only owns `valid - 96` rows
                const int r0 = mt * cfg.valid + q * 32;
                if (q < 3) tma_store_3d(&tmO, smO + q * O_SLAB, 0, r0, b);
                else tma_store_3d(&tmOtail, smO + q * O_SLAB, 0, r0, b);
                tma_store_commit();
            }
        };

        const bool is_e1 = warp < 8;      // warps 4-7: epilogue 1 (conv1 -> t tile); warps 8-11: epilogue 2 (output)
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int mt = tile % cfg.m_tiles, b = tile / cfg.m_tiles;
            // ---- epilogue 1: t = lrelu(c1 + b1) -> fp16 swizzled smem tile (zeros outside the utterance)
            if (is_e1) {
                const int bb = it & 1; const uint32_t ph = (uint32_t)((it >> 1) & 1);
                mbar_wait(&acc1_full[bb], ph);
                mbar_wait(&t_empty[bb], ph ^ 1);
                tc_fence_after();
                const int trow = mt * cfg.valid - P2 + row;                 // global row of this t row
                const bool inside = trow >= 0 && trow < p.L;
                const uint32_t taddr = tmem_base + (uint32_t)(bb * C) + lane_addr;
                uint8_t* trow_ptr = smT + bb * T_ALLOC + row * ROW_BYTES;
#pragma unroll
                for (int c = 0; c < C / 16; ++c) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c * 16, r);
                    tmem_ld_wait();
                    uint4 u0, u1;
                    __half2* p0 = reinterpret_cast<__half2*>(&u0);
                    __half2* p1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float a0 = lrelu(__uint_as_float(r[2 * i]) + p.b1[c * 16 + 2 * i], p.t_slope);
                        float a1 = lrelu(__uint_as_float(r[2 * i + 1]) + p.b1[c * 16 + 2 * i + 1], p.t_slope);
                        float c0 = lrelu(__uint_as_float(r[8 + 2 * i]) + p.b1[c * 16 + 8 + 2 * i], p.t_slope);
                        float c1 = lrelu(__uint_as_float(r[8 + 2 * i + 1]) + p.b1[c * 16 + 8 + 2 * i + 1], p.t_slope);
                        if (!inside) { a0 = a1 = c0 = c1 = 0.f; }
                        p0[i] = __floats2half2_rn(a0, a1);
                        p1[i] = __floats2half2_rn(c0, c1);
                    }
                    *reinterpret_cast<uint4*>(trow_ptr + (((2 * c) ^ swz) << 4)) = u0;
                    *reinterpret_cast<uint4*>(trow_ptr + (((2 * c + 1) ^ swz) << 4)) = u1;
                }
                tc_fence_before();
                fence_proxy_async_smem();          // t tile (generic-proxy writes) -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) { mbar_arrive(&acc1_empty[bb]); mbar_arrive(&t_full[bb]); }
            }
            // ---- epilogue 2 (its own warps, running concurrently with epilogue 1 of later tiles)
            if (!is_e1) epilogue2(it, mt, b);
        }
        if (!is_e1 && lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int C, int TAPS>
int launch_rb_cfg(const UmmaResblockParams& p, cudaStream_t s) {
    constexpr int BK = C;
    constexpr int ROW_BYTES = BK * 2;
    constexpr int ROW_ALIGN = 1024 / ROW_BYTES;
    constexpr size_t W_BYTES = (size_t)2 * TAPS * C * ROW_BYTES;
    constexpr size_t T_BYTES = (size_t)2 * T_ROWS_ALLOC * ROW_BYTES;
    constexpr size_t O_BYTES = (size_t)4 * 32 * ROW_BYTES;
    constexpr size_t FIXED = (2 * RB_MAX_STAGES + 1 + 12) * 8 + 16 + 1024;
    constexpr size_t LIMIT = 227 * 1024;
    static_assert((T_ROWS_ALLOC * ROW_BYTES) % 1024 == 0, "t tile must keep the swizzle alignment");

    RbCfg cfg{};
    cfg.valid = 128 - (TAPS - 1);
    cfg.box_rows = 128 + (TAPS - 1) * p.dil;
    if (cfg.box_rows > 256) return CMTTS_ERR_UNSUPPORTED;
    cfg.rows_alloc = (cfg.box_rows + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    cfg.m_tiles = (p.L + cfg.valid - 1) / cfg.valid;
    const size_t a_alloc = (size_t)cfg.rows_alloc * ROW_BYTES;
    const size_t rest = W_BYTES + T_BYTES + O_BYTES + FIXED;
    if (rest + 2 * a_alloc > LIMIT) return CMTTS_ERR_UNSUPPORTED;
    size_t st = (LIMIT - rest) / a_alloc;
    cfg.a_stages = (int)(st > RB_MAX_STAGES ? RB_MAX_STAGES : st);
    const size_t smem = rest + (size_t)cfg.a_stages * a_alloc;

    auto kern = umma_resblock_kernel<C, TAPS>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIMIT) != cudaSuccess) {
            cmtts_set_error("umma_resblock: cannot set dynamic shared memory size", __FILE__, __LINE__);
            return CMTTS_ERR_CUDA;
        }
        attr_done = true;
    }
    CUtensorMap a_map, w1_map, w2_map, o_map, ot_map;
    const long long bs = (long long)p.L * C;
    if (!make_act_map(&a_map, p.a, C, p.L, p.B, C, bs, BK, cfg.box_rows) ||
        !make_w_map(&w1_map, p.w1, C, TAPS * C, BK, C) || !make_w_map(&w2_map, p.w2, C, TAPS * C, BK, C) ||
        !make_act_map(&o_map, p.out_h, C, p.L, p.B, C, bs, BK, 32) ||
        !make_act_map(&ot_map, p.out_h, C, p.L, p.B, C, bs, BK, cfg.valid - 96)) {
        cmtts_set_error("umma_resblock: cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    const int tiles = p.B * cfg.m_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    kern<<<grid, 384, smem, s>>>(a_map, w1_map, w2_map, o_map, ot_map, p, cfg);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

}  // namespace

// CMTTS_ERR_UNSUPPORTED (no error text) when the shape is not covered -> caller runs the two convs separately.
int launch_umma_resblock(const UmmaResblockParams& p, cudaStream_t s) {
    if (p.B == 0 || p.L == 0) return CMTTS_OK;
    if (p.dil < 1 || !p.a || !p.w1 || !p.w2 || !p.b1 || !p.b2 || !p.out_h) return CMTTS_ERR_UNSUPPORTED;
    if (((uintptr_t)p.a % 16) || ((uintptr_t)p.out_h % 16)) return CMTTS_ERR_UNSUPPORTED;
#define RB_DISPATCH(C_)                                                   \
    switch (p.taps) {                                                     \
        case 3: return launch_rb_cfg<C_, 3>(p, s);                        \
        case 7: return launch_rb_cfg<C_, 7>(p, s);                        \
        case 11: return launch_rb_cfg<C_, 11>(p, s);                      \
        default: return CMTTS_ERR_UNSUPPORTED;                            \
    }
    if (p.C == 32) { RB_DISPATCH(32) }
    if (p.C == 64) { RB_DISPATCH(64) }
#undef RB_DISPATCH
    return CMTTS_ERR_UNSUPPORTED;
}
