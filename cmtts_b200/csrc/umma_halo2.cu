// CTA-PAIR (tcgen05 cta_group::2) variant of the halo-tile conv for the HiFi-GAN ResBlock convolutions with C = 128 and
// k = 7 / 11 (level 1: 12 launches, 4.7 ms of the 29 ms C2 step).
//
// Why.  These shapes cannot keep their weights in shared memory (k * 128 * 128 * 2 B = 229 / 360 KB), so every 128-row tile
// re-streams the whole weight set from L2: 47 KB of activations + 360 KB of weights per tile.  Measured, an SM takes in
// ~43 B / cycle through TMA, i.e. 9.5 k cycles per tile against 5.6 k cycles of tensor-pipe work (88 MMAs of 128 x 128 x 16):
// the kernel is bound by bytes INTO the SM, not by HBM (traffic = algorithmic), not by the pipe (71-78 % active), and not by
// L2 reads — which is why a TMA multicast of the weights across a cluster buys nothing (every SM would still take in the
// same bytes; the L2 already serves identical requests of a few SMs from one read).
// What halves the intake is the CTA pair: ONE tcgen05.mma with cta_group::2 computes a 256 x 128 tile from both SMs' shared
// memory — each CTA holds its own 128 rows of A and only HALF of the weight block (64 of the 128 output channels).  Per tile
// an SM now takes in 47 + 180 KB = 5.3 k cycles, under the 5.6 k cycles of MMA work.
//
// Structure (one cluster = 2 CTAs on the two SMs of a TPC; rank 0 is the leader):
//   * both CTAs run a producer for their own halo tile (rows of THEIR half of the 256-row tile) and for their half of every
//     weight block; the loads are `cp.async.bulk.tensor ... cta_group::2` whose completion is signalled on the LEADER's `full`
//     barrier, which the leader arms with the byte count of both CTAs;
//   * only the leader issues MMAs (M = 256); its tcgen05.commit is multicast to the `empty` / `accumulator full` barriers of
//     both CTAs, so each CTA's producers and epilogue warps wait on their own barriers;
//   * the accumulator of rows 0-127 lives in the leader's TMEM, rows 128-255 in the peer's: each CTA runs the same 8-warp
//     epilogue as umma_halo_kernel on its own rows; `accumulator empty` is ONE barrier in the leader (16 arrivals, the peer's
//     eight arrive remotely);
//   * TMEM is allocated / freed with cta_group::2 by the same warp of both CTAs; the cluster synchronises after barrier
//     initialisation and before exit (a CTA must not leave while its peer can still signal into it).
// An odd last row tile leaves the peer's half out of range: its loads are zero-filled by TMA and its epilogue writes nothing.
#include "umma_common.cuh"
#include <stdlib.h>

namespace {

using namespace umma;

constexpr int H2_BK = 64;
constexpr int H2_ROW_BYTES = H2_BK * 2;                 // 128
constexpr int H2_MAX_A = 8, H2_MAX_W = 6;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;             // clears the CTA-rank bit of a shared::cluster address -> the leader's copy

struct Halo2Cfg { int a_stages, w_stages, rows_alloc, box_rows; };

__device__ __forceinline__ float lrelu_fwd2(float v, float slope) { return fmaxf(v, v * slope); }
__device__ __forceinline__ float lrelu_inv2(float v, float inv_slope) { return fminf(v, v * inv_slope); }

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads of a CTA pair: data lands in the EXECUTING CTA's shared memory, completion bytes go to the LEADER's barrier
__device__ __forceinline__ void tma2_load_3d_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// commit of the pair's MMAs, signalled on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the LEADER's copy of a barrier (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

template <int C, int TAPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
umma_halo2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmO, const UmmaConvParams p, const Halo2Cfg cfg) {
    constexpr int BM = 128, BN = C, BK = H2_BK, CB = C / BK;
    constexpr int ROW_BYTES = H2_ROW_BYTES;
    constexpr int WN = C / 2;                      // weight rows (output channels) held by ONE CTA
    constexpr int W_BLK = WN * ROW_BYTES;          // this CTA's half of one (channel block, tap) weight block: 8 / 16 KB
    constexpr int BNH = BN / 2;                    // output columns per epilogue warp (64 / 128)
    constexpr int NH = BNH / 64;                   // ... staged and stored 64 columns at a time
    constexpr int OROW = 128;                      // bytes per staged output row (64 fp16)
    constexpr int O_SLAB = 32 * OROW;
    constexpr int O_BYTES = 8 * O_SLAB;
    constexpr int TMEM_COLS = 2 * BN;              // double-buffered accumulator: 256 / 512 columns in EACH CTA

    const int a_alloc = cfg.rows_alloc * ROW_BYTES;        // one 64-channel halo block
    const bool has_res = p.res_h != nullptr;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* smA = smem;
    uint8_t* smO = smA + cfg.a_stages * a_alloc;
    uint8_t* smW = smO + O_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smW + (size_t)cfg.w_stages * W_BLK);
    uint64_t* a_empty = a_full + H2_MAX_A;
    uint64_t* w_full = a_empty + H2_MAX_A;
    uint64_t* w_empty = w_full + H2_MAX_W;
    uint64_t* tfull = w_empty + H2_MAX_W;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();       // 0 = leader
    const int m_tiles = (p.M + BM - 1) / BM;
    const int m_pairs = (m_tiles + 1) / 2;
    const int pairs = p.B * m_pairs;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int shift0 = p.shift[0];

    if (threadIdx.x == 0) {
        for (int i = 0; i < H2_MAX_A; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < H2_MAX_W; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }   // 8 epilogue warps of each CTA
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < BN; i += blockDim.x) s_bias[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                            // the peer's barriers exist before anything is signalled into them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    if (warp != 3) pdl_wait();                     // the weight producer reads constants only

    if (warp == 0) {
        // ======================= activation producer: this CTA's halo tile, one 64-channel block per ring stage ==========
        prefetch_tmap(&tmA);
        int stage = 0; uint32_t phase = 0;
        const uint32_t bytes_pair = (uint32_t)(2 * cfg.box_rows * ROW_BYTES);          // both CTAs' blocks
        for (int pt = cluster_id; pt < pairs; pt += n_clusters) {
            const int mp = pt % m_pairs, b = pt / m_pairs;
            const int mt = 2 * mp + rank;
#pragma unroll 1
            for (int cb = 0; cb < CB; ++cb) {
                mbar_wait(&a_empty[stage], phase ^ 1);
                if (rank == 0) mbar_expect_tx_elect(&a_full[stage], bytes_pair);
                tma2_load_3d_elect(smA + stage * a_alloc, &tmA, &a_full[stage], cb * BK, mt * BM + shift0, b);
                if (++stage == cfg.a_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ======================= weight producer: this CTA's half (C / 2 output channels) of every block ================
        prefetch_tmap(&tmW);
        int ws = 0; uint32_t wphase = 0;
        for (int pt = cluster_id; pt < pairs; pt += n_clusters) {
#pragma unroll 1
            for (int cb = 0; cb < CB; ++cb)
#pragma unroll 1
                for (int tap = 0; tap < TAPS; ++tap) {
                    mbar_wait(&w_empty[ws], wphase ^ 1);
                    if (rank == 0) mbar_expect_tx_elect(&w_full[ws], (uint32_t)(2 * W_BLK));
                    tma2_load_2d_elect(smW + ws * W_BLK, &tmW, &w_full[ws], cb * BK, tap * p.N + rank * WN);
                    if (++ws == cfg.w_stages) { ws = 0; wphase ^= 1; }
                }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer: the leader only, M = 256 across the pair, N = C =======================
        if (rank == 0) {
            const uint32_t tmem_u = make_uniform(tmem_base);
            const uint32_t idesc = make_idesc(2 * BM, BN);
            constexpr uint32_t DESC_HI = (uint32_t)((8 * ROW_BYTES) >> 4) | (1u << 14) | (2u << 29);
            constexpr uint64_t HI = (uint64_t)DESC_HI << 32;
            const uint32_t tap_step_u = make_uniform((uint32_t)(((TAPS > 1 ? p.shift[1] - p.shift[0] : 0) * ROW_BYTES) >> 4));
            const uint32_t w_lo0 = make_uniform(((smem_u32(smW) >> 4) & 0x3FFF) | (1u << 16));
            const uint32_t a_base = make_uniform(((smem_u32(smA) >> 4) & 0x3FFF) | (1u << 16));
            const uint32_t a_alloc16 = make_uniform((uint32_t)(a_alloc >> 4));
            int stage = 0; uint32_t phase = 0;
            int ws = 0; uint32_t wphase = 0;
            int it = 0;
            for (int pt = cluster_id; pt < pairs; pt += n_clusters, ++it) {
                const int abuf = it & 1; const uint32_t aphase = (uint32_t)((it >> 1) & 1);
                mbar_wait(&tempty[abuf], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = make_uniform(tmem_u + (uint32_t)(abuf * BN));
#pragma unroll 1
                for (int cb = 0; cb < CB; ++cb) {
                    mbar_wait(&a_full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo0 = make_uniform(a_base + (uint32_t)stage * a_alloc16);
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap) {
                        mbar_wait(&w_full[ws], wphase);
                        tc_fence_after();
                        const uint32_t a_lo = a_lo0 + tap * tap_step_u;
                        const uint32_t w_lo = make_uniform(w_lo0 + (uint32_t)((ws * W_BLK) >> 4));
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma2_f16(d_tmem, HI | (a_lo + 2 * k), HI | (w_lo + 2 * k), idesc, (cb | tap | k) ? 1u : 0u);
                            umma2_commit_both(&w_empty[ws]);
                            if (tap == TAPS - 1) umma2_commit_both(&a_empty[stage]);
                        }
                        __syncwarp();
                        if (++ws == cfg.w_stages) { ws = 0; wphase ^= 1; }
                    }
                    if (++stage == cfg.a_stages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma2_commit_both(&tfull[abuf]);
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ======================= epilogue (UEPI_VOC), 8 warps, this CTA's 128 rows =======================
        // warp = (TMEM lane quarter, column half); a warp's BNH columns are staged and TMA-stored 64 at a time
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const int swz_o = lane & 7;
        uint8_t* slab = smO + (warp - 4) * O_SLAB + lane * OROW;
        int abuf = 0; uint32_t aphase = 0;
        for (int pt = cluster_id; pt < pairs; pt += n_clusters) {
            const int mp = pt % m_pairs, b = pt / m_pairs;
            const int mt = 2 * mp + rank;
            const int t = mt * BM + row;
            const bool valid = t < p.M;
            {
                const int pn = pt + n_clusters;    // next tile's residual / partial-sum lines -> L2
                if (pn < pairs) {
                    const int t2 = (2 * (pn % m_pairs) + rank) * BM + row, b2 = pn / m_pairs;
                    if (t2 < p.M) {
#pragma unroll
                        for (int hh = 0; hh < NH; ++hh) {
                            const int nb = h * BNH + hh * 64;
                            if (has_res) prefetch_l2(p.res_h + (long long)b2 * p.res_bstride + (long long)t2 * p.res_ld + nb);
                            if (p.sum_h) prefetch_l2(p.sum_h + (long long)b2 * p.out_bstride + (long long)t2 * p.out_ld + nb);
                        }
                    }
                }
            }
            const bool has_sum = p.sum_h != nullptr && valid;
            uint4 rres[8], rsum[8];
            auto load_res = [&](int nb) {
                if (has_res) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.res_h + (long long)b * p.res_bstride + (long long)t * p.res_ld + nb);
#pragma unroll
                    for (int i = 0; i < 8; ++i) rres[i] = valid ? rp[i] : make_uint4(0u, 0u, 0u, 0u);
                }
                if (has_sum) {
                    const uint4* sp = reinterpret_cast<const uint4*>(p.sum_h + (long long)b * p.out_bstride + (long long)t * p.out_ld + nb);
#pragma unroll
                    for (int i = 0; i < 8; ++i) rsum[i] = sp[i];
                }
            };
            load_res(h * BNH);                      // first 64 columns: issued before the accumulator wait
            mbar_wait(&tfull[abuf], aphase);
            tc_fence_after();
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {
                const int n_base = h * BNH + hh * 64;
                if (hh > 0) load_res(n_base);
                const uint32_t taddr = tmem_base + (uint32_t)(abuf * BN + n_base) + ((uint32_t)(q * 32) << 16);
                if (lane == 0) tma_store_wait_read();           // the previous store has finished READING this warp's slab
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
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
                            v[2 * i] += lrelu_inv2(a.x, p.res_inv_slope); v[2 * i + 1] += lrelu_inv2(a.y, p.res_inv_slope);
                            v[8 + 2 * i] += lrelu_inv2(bb.x, p.res_inv_slope); v[8 + 2 * i + 1] += lrelu_inv2(bb.y, p.res_inv_slope);
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
                    uint4 u0, u1;
                    __half2* p0 = reinterpret_cast<__half2*>(&u0);
                    __half2* p1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        p0[i] = __floats2half2_rn(lrelu_fwd2(v[2 * i], p.out_slope), lrelu_fwd2(v[2 * i + 1], p.out_slope));
                        p1[i] = __floats2half2_rn(lrelu_fwd2(v[8 + 2 * i], p.out_slope), lrelu_fwd2(v[8 + 2 * i + 1], p.out_slope));
                    }
                    *reinterpret_cast<uint4*>(slab + (((2 * c) ^ swz_o) << 4)) = u0;
                    *reinterpret_cast<uint4*>(slab + (((2 * c + 1) ^ swz_o) << 4)) = u1;
                }
                if (hh == NH - 1) tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (hh == NH - 1) mbar_arrive_leader(&tempty[abuf]);      // all TMEM reads of this tile are done
                    if (mt < m_tiles) {
                        tma_store_3d(&tmO, smO + (warp - 4) * O_SLAB, n_base, mt * BM + q * 32, b);
                        tma_store_commit();
                    }
                }
            }
            abuf ^= 1; if (abuf == 0) aphase ^= 1;
        }
        if (lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                            // nobody leaves while the peer may still signal into this CTA
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int C, int TAPS>
int launch_halo2_cfg(const UmmaConvParams& p, cudaStream_t s) {
    constexpr int ROW_BYTES = H2_ROW_BYTES;
    constexpr int ROW_ALIGN = 1024 / ROW_BYTES;
    constexpr size_t LIMIT = 227 * 1024;
    constexpr size_t O_BYTES = (size_t)8 * 32 * 128;
    constexpr size_t W_BLK = (size_t)(C / 2) * ROW_BYTES;
    constexpr size_t FIXED = (2 * H2_MAX_A + 2 * H2_MAX_W + 4) * 8 + 32 + C * 4 + 1024 + O_BYTES;
    constexpr int CB = C / H2_BK;
    Halo2Cfg cfg{};
    const int span = p.shift[p.taps - 1] - p.shift[0];
    cfg.box_rows = 128 + span;
    if (cfg.box_rows > 256) return CMTTS_ERR_UNSUPPORTED;
    cfg.rows_alloc = (cfg.box_rows + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    const size_t a_alloc = (size_t)cfg.rows_alloc * ROW_BYTES;
    cfg.w_stages = C >= 256 ? 4 : H2_MAX_W;
    const size_t w_bytes = (size_t)cfg.w_stages * W_BLK;
    if (FIXED + w_bytes + (size_t)CB * a_alloc > LIMIT) return CMTTS_ERR_UNSUPPORTED;      // at least one whole tile of A blocks
    size_t a = (LIMIT - FIXED - w_bytes) / a_alloc;
    cfg.a_stages = (int)(a > H2_MAX_A ? H2_MAX_A : a);
    const size_t smem = (size_t)cfg.a_stages * a_alloc + w_bytes + FIXED;

    auto kern = umma_halo2_kernel<C, TAPS>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIMIT) != cudaSuccess) {
            cmtts_set_error("umma_halo2: cannot set dynamic shared memory size", __FILE__, __LINE__);
            return CMTTS_ERR_CUDA;
        }
        attr_done = true;
    }
    CUtensorMap a_map, w_map, o_map;
    if (!make_act_map(&a_map, p.a_hi, p.Cin, p.Lin, p.B, p.a_ld, p.a_bstride, H2_BK, cfg.box_rows) ||
        !make_w_map(&w_map, p.w_hi, p.Cin, p.taps * p.N, H2_BK, C / 2) ||
        !make_act_map(&o_map, p.out_h, p.N, p.M, p.B, p.out_ld, p.out_bstride, 64, 32)) {
        cmtts_set_error("umma_halo2: cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
        return CMTTS_ERR_CUDA;
    }
    const int m_tiles = (p.M + 127) / 128;
    const int pairs = p.B * ((m_tiles + 1) / 2);
    int grid = 2 * pairs < num_sms() ? 2 * pairs : num_sms();
    grid &= ~1;                                          // whole CTA pairs
    if (grid < 2) return CMTTS_ERR_UNSUPPORTED;
    if (g_cmtts_prof_on) {
        const double rows = (double)p.B * p.M;
        char lbl[96];
        snprintf(lbl, sizeof(lbl), "umma_halo2<%d> k%d d%d%s%s (CTA pairs)", p.N, p.taps, p.taps > 1 ? p.shift[1] - p.shift[0] : 1,
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

// CTA-pair variant for C = 128 with streamed weights (k = 7, 11) and for C = 256 (k = 3, 7, 11: HiFi-GAN level 0, which the
// one-CTA halo kernel does not cover at all); CMTTS_ERR_UNSUPPORTED otherwise (the caller falls back).  CMTTS_UMMA_DBG bit
// 512 / CMTTS_HALO2=0 switch it off, CMTTS_HALO2=1 restricts it to C = 128.
int launch_umma_halo2(const UmmaConvParams& p, cudaStream_t s) {
    static int env = -1;
    if (env < 0) { const char* e = getenv("CMTTS_HALO2"); env = e ? atoi(e) : 2; }
    if (!env || (p.dbg & 512)) return CMTTS_ERR_UNSUPPORTED;
    if (p.Cin != p.N) return CMTTS_ERR_UNSUPPORTED;
    if (p.B == 0 || p.M == 0) return CMTTS_OK;
    if (p.N == 128) {
        if (p.taps == 7) return launch_halo2_cfg<128, 7>(p, s);
        if (p.taps == 11) return launch_halo2_cfg<128, 11>(p, s);
    } else if (p.N == 256 && env >= 2) {
        if (p.taps == 3) return launch_halo2_cfg<256, 3>(p, s);
        if (p.taps == 7) return launch_halo2_cfg<256, 7>(p, s);
        if (p.taps == 11) return launch_halo2_cfg<256, 11>(p, s);
    }
    return CMTTS_ERR_UNSUPPORTED;
}
