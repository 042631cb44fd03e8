// Shared device helpers of the tcgen05 kernels: PTX wrappers (mbarrier, TMA, tcgen05.mma/ld/commit),
// UMMA descriptor builders and small vector load/store helpers.
#pragma once
#include "umma_conv.cuh"
#include <cuda.h>
#include <cuda_fp8.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdlib.h>

namespace umma {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 1024-byte alignment of the dynamic shared-memory base as POINTER ARITHMETIC on the __shared__ array: rounding the address up
// through uintptr_t turned every pointer derived from it into a generic one (LD.E / ST.E: generic address translation in front
// of each staging-slab, bias and t-tile access, and no reordering against global stores); an offset keeps the address space,
// so the same accesses compile to LDS / STS.
__device__ __forceinline__ uint8_t* smem_align1024(uint8_t* base) { return base + ((1024u - (smem_u32(base) & 1023u)) & 1023u); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// TMA store of a shared-memory box (bulk async group); OOB rows/channels are clipped by the TMA unit
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Programmatic dependent launch (PDL).  Every tcgen05 kernel here is launched with the programmatic-stream-serialization
// attribute (launch_pdl below): its CTAs may become resident — and run their prologue: barrier init, TMEM allocation,
// tensor-map prefetch, loads of constant weights — while the previous kernel of the stream is still draining its last
// tiles.  pdl_wait() blocks until that kernel has completed and its writes are visible; it precedes every access to
// memory another kernel produces.  pdl_trigger() lets the NEXT kernel's CTAs be scheduled as soon as this grid's
// CTAs are all running (they take an SM only once its current CTA has exited: shared memory does not fit twice).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// kind::f8f6f4 with e4m3 x e4m3 -> f32: the instruction descriptor has the same bit pattern as f16 x f16 -> f32 (format
// codes 0), K = 32 per instruction (32 bytes of a K-major row: the descriptor start address advances by 2 like fp16's K = 16)
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// One elected lane of a converged warp.  `if (elect_one()) { tcgen05.mma ...; tcgen05.commit ...; }` is the form ptxas
// turns into bare warp-level UTCHMMA / UTCBAR instructions — PROVIDED every operand is derived from warp-uniform values
// only (make_uniform() of the shared-memory / TMEM bases, compile-time ring indices, kernel parameters): then descriptors
// live in uniform registers and a K = 64 block is 4 back-to-back UTCHMMA plus a few UIADD3.  With operands built from
// ring indices in vector registers the same source costs 13-17 SASS instructions per MMA (ELECT / VOTEU / five R2UR),
// which made the issuing warp — not the tensor pipe — the limiter of the hi/lo kernels (profiles/ncu_r1_gate_*.txt).
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(p));
    return p != 0;
}
// Warp-convergent predicated variants (the whole warp executes them with identical operands; the instruction itself is
// predicated on elect.sync in PTX).  Still used for barrier commits outside elect blocks and by the TMA producers; the
// MMA issue loops use `if (elect_one())` + umma_f16 / umma_commit instead (see above).
__device__ __forceinline__ void umma_f16_pred(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accum, uint32_t /*issue*/) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pred(uint64_t* bar, uint32_t /*issue*/) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
// One-lane (elect.sync) variants of the producer-side instructions, for warp-convergent producer loops
__device__ __forceinline__ void mbar_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// L2 prefetch of a TMA box (no shared memory, no barrier).  Experiment knob (CMTTS_PF): the vocoder kernels turned out
// not to be bound by the latency of their input loads, so it is off by default.
__device__ __forceinline__ void tma_prefetch_3d_elect(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];\n\t}"
        ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// value known to be identical in all lanes -> move it to a uniform register (REDUX writes a UR)
__device__ __forceinline__ uint32_t make_uniform(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) | [32,46) SBO >> 4 (8 rows)
//   [46,48) version = 1 | [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
template <int BK>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    constexpr uint64_t row_bytes = BK * 2;                    // 128 or 64
    constexpr uint64_t sbo = (8 * row_bytes) >> 4;            // 8-row core-matrix group stride
    constexpr uint64_t layout = (BK == 64) ? 2ull : 4ull;
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, K-major both
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

struct H8 { __half2 a, b, c, d; };  // 16 bytes

__device__ __forceinline__ void load16h(const __half* p, float (&f)[16]) {
    const uint4 u0 = *reinterpret_cast<const uint4*>(p);
    const uint4 u1 = *reinterpret_cast<const uint4*>(p + 8);
    const __half2* h0 = reinterpret_cast<const __half2*>(&u0);
    const __half2* h1 = reinterpret_cast<const __half2*>(&u1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = __half22float2(h0[i]), b = __half22float2(h1[i]);
        f[2 * i] = a.x; f[2 * i + 1] = a.y; f[8 + 2 * i] = b.x; f[8 + 2 * i + 1] = b.y;
    }
}
__device__ __forceinline__ void store16h(__half* p, const float (&f)[16]) {
    uint4 u0, u1;
    __half2* h0 = reinterpret_cast<__half2*>(&u0);
    __half2* h1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h0[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        h1[i] = __floats2half2_rn(f[8 + 2 * i], f[8 + 2 * i + 1]);
    }
    *reinterpret_cast<uint4*>(p) = u0;
    *reinterpret_cast<uint4*>(p + 8) = u1;
}
// v = hi + lo with hi = fp16(v), lo = fp16(v - hi)
__device__ __forceinline__ void store16_hilo(__half* hi, __half* lo, const float (&f)[16]) {
    float h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const __half hh = __float2half_rn(f[i]);
        h[i] = __half2float(hh);
        l[i] = f[i] - h[i];
    }
    store16h(hi, h);
    store16h(lo, l);
}
// packed forms of the two helpers above/below (for epilogues that stage their output in shared memory for a TMA store)
__device__ __forceinline__ void pack16_hilo(const float (&f)[16], uint4& h0, uint4& h1, uint4& l0, uint4& l1) {
    __half2* ph0 = reinterpret_cast<__half2*>(&h0); __half2* ph1 = reinterpret_cast<__half2*>(&h1);
    __half2* pl0 = reinterpret_cast<__half2*>(&l0); __half2* pl1 = reinterpret_cast<__half2*>(&l1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 a = __floats2half2_rn(f[2 * i], f[2 * i + 1]), b = __floats2half2_rn(f[8 + 2 * i], f[8 + 2 * i + 1]);
        const float2 af = __half22float2(a), bf = __half22float2(b);
        ph0[i] = a; ph1[i] = b;
        pl0[i] = __floats2half2_rn(f[2 * i] - af.x, f[2 * i + 1] - af.y);
        pl1[i] = __floats2half2_rn(f[8 + 2 * i] - bf.x, f[8 + 2 * i + 1] - bf.y);
    }
}
__device__ __forceinline__ void pack16_f8pair(const float (&f)[16], uint4& h8, uint4& l8) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t hw = 0, lw = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float a = f[4 * i + 2 * j], b = f[4 * i + 2 * j + 1];
            const float ah = __half2float(__float2half_rn(a)), bh = __half2float(__float2half_rn(b));
            const uint32_t ph = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(ah, bh), __NV_SATFINITE, __NV_E4M3);
            const uint32_t pl = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2((a - ah) * 2048.f, (b - bh) * 2048.f), __NV_SATFINITE, __NV_E4M3);
            hw |= ph << (16 * j); lw |= pl << (16 * j);
        }
        h[i] = hw; l[i] = lw;
    }
    h8 = make_uint4(h[0], h[1], h[2], h[3]);
    l8 = make_uint4(l[0], l[1], l[2], l[3]);
}
// e4m3 pair of 16 values for the fp8 cross terms: hi8 = e4m3(fp16(v)), lo8 = e4m3((v - fp16(v)) * 2^11)   (16 bytes each)
__device__ __forceinline__ void store16_f8pair(unsigned char* hi8, unsigned char* lo8, const float (&f)[16]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t hw = 0, lw = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float a = f[4 * i + 2 * j], b = f[4 * i + 2 * j + 1];
            const float ah = __half2float(__float2half_rn(a)), bh = __half2float(__float2half_rn(b));
            const uint32_t ph = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(ah, bh), __NV_SATFINITE, __NV_E4M3);
            const uint32_t pl = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2((a - ah) * 2048.f, (b - bh) * 2048.f), __NV_SATFINITE, __NV_E4M3);
            hw |= ph << (16 * j); lw |= pl << (16 * j);
        }
        h[i] = hw; l[i] = lw;
    }
    *reinterpret_cast<uint4*>(hi8) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo8) = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void load16f(const float* p, float (&f)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(p + 4 * i);
        f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
    }
}
__device__ __forceinline__ void store16f(float* p, const float (&f)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(p + 4 * i) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}

constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }


// ------------------------------------------------------------------------------------------
// host side: TMA tensor maps
// ------------------------------------------------------------------------------------------
static inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// activations [B][L][C] fp16 (row stride ld, batch stride bstride): box {BK channels, 128 rows, 1}
static inline bool make_act_map(CUtensorMap* m, const __half* base, int C, int L, int B, int ld, long long bstride, int BK,
                                int box_rows = 128) {
    auto enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bstride * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// byte tensors (e4m3 operands) [B][L][C]: box {128 bytes, box_rows, 1}, 128B swizzle
static inline bool make_act_map8(CUtensorMap* m, const unsigned char* base, int C, int L, int B, int ld, long long bstride, int box_rows) {
    auto enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)bstride};
    cuuint32_t box[3] = {128u, (cuuint32_t)box_rows, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<unsigned char*>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// fp32 planes [P][L][C] as the TARGET of TMA stores: box {32 floats (128 bytes), box_rows, 1}, 128B swizzle
static inline bool make_store_map_f32(CUtensorMap* m, float* base, int C, int L, int P, int ld, long long pstride, int box_rows) {
    auto enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)P};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)pstride * 4};
    cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// byte tensors [B][L][C] as the TARGET of 64-byte-wide TMA stores: box {64 bytes, box_rows, 1}, 64B swizzle
static inline bool make_store_map8(CUtensorMap* m, unsigned char* base, int C, int L, int B, int ld, long long bstride, int box_rows) {
    auto enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)bstride};
    cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static inline bool make_w_map8(CUtensorMap* m, const unsigned char* base, int Cin, int rows, int BN) {
    auto enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Cin};
    cuuint32_t box[2] = {128u, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<unsigned char*>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// weights [taps*N][Cin] fp16: box {BK, BN}
static inline bool make_w_map(CUtensorMap* m, const __half* base, int Cin, int rows, int BK, int BN) {
    auto enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    const CUtensorMapSwizzle sw = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// kernel launch with the PDL attribute (CMTTS_PDL=0 turns the attribute off: plain stream order)
static inline bool pdl_enabled() {
    if (g_cmtts_pdl < 0) { const char* e = getenv("CMTTS_PDL"); g_cmtts_pdl = e ? (atoi(e) != 0) : 1; }
    return g_cmtts_pdl != 0;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

}  // namespace umma
