// fp32 implicit-GEMM conv1d over channels-last activations (CUDA cores, FFMA).
//
// This is the accuracy-critical path: the text encoder and the variance adaptor feed three
// quantisers (duration rounding model/modules.py:369-372, energy bucketize :326-328, f0_to_coarse
// utils/pitch_tools.py:26-35) that flip under TF32-sized perturbations (SURVEY.md §7 "Hard
// parts"), so everything upstream of them stays in fp32 FFMA with fp32 accumulation.  It is also
// the kernel the tcgen05 paths are checked against on the device.
//
// Tiling: CTA tile 128 (time rows) x BN (output channels), K chunk 16 per (tap, channel block),
// 2*BN threads, 8x8 register micro-tile per thread split into 4x4 quadrants so that every
// shared-memory fragment read is one conflict-free LDS.128.  Global->register prefetch of the next
// K chunk overlaps the FFMA block of the current one (double-buffered shared memory, one
// __syncthreads per chunk).  A time tile never crosses an utterance, so zero padding at sequence
// ends is a row predicate.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int APAD = 4;

__device__ __forceinline__ float act_apply(float v, int act, float slope) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.f);
        case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        case ACT_LRELU: return v > 0.f ? v : v * slope;
        case ACT_TANH: return tanhf(v);
        case ACT_SWISH: return v / (1.f + expf(-v));
        default: return v;
    }
}

template <int BN>
__global__ void __launch_bounds__(2 * BN) conv1d_simt_kernel(const ConvParams p) {
    constexpr int NT = 2 * BN;
    constexpr int TX = BN / 8;
    constexpr int A_LD = BM + APAD;
    constexpr int B_LD = BN + 4;
    constexpr int A_F4 = BM * BK / 4 / NT;       // float4 loads of A per thread per chunk
    constexpr int B_F4 = (BK * BN / 4 + NT - 1) / NT;  // = 2

    extern __shared__ __align__(16) float smem[];
    float* As = smem;                      // [2][BK][A_LD]
    float* Bs = smem + 2 * BK * A_LD;      // [2][BK][B_LD]

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int n0 = blockIdx.x * BN;
    const int t0 = blockIdx.y * BM;
    const int b = blockIdx.z;

    const float* xb = p.x + (long long)b * p.x_bstride;
    const int cblocks = p.Cin / BK;
    const int nchunks = p.taps * cblocks;

    float4 a_reg[A_F4];
    float4 b_reg[B_F4];

    auto load_chunk = [&](int chunk) {
        const int tap = chunk / cblocks;
        const int c0 = (chunk - tap * cblocks) * BK;
        const int sh = p.shift[tap];
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int idx = tid + i * NT;
            const int row = idx >> 2, kq = idx & 3;
            const int src = t0 + row + sh;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src >= 0 && src < p.Lin && (t0 + row) < p.M) {
                v = *reinterpret_cast<const float4*>(xb + (long long)src * p.x_ld + c0 + kq * 4);
                if (p.pre_lrelu) {
                    v.x = v.x > 0.f ? v.x : v.x * p.pre_slope;
                    v.y = v.y > 0.f ? v.y : v.y * p.pre_slope;
                    v.z = v.z > 0.f ? v.z : v.z * p.pre_slope;
                    v.w = v.w > 0.f ? v.w : v.w * p.pre_slope;
                }
            }
            a_reg[i] = v;
        }
        const float* wk = p.w + ((long long)tap * p.Cin + c0) * p.N;
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            const int idx = tid + i * NT;
            const int k = idx / (BN / 4), nq = idx % (BN / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < BK && n0 + nq * 4 < p.N)
                v = *reinterpret_cast<const float4*>(wk + (long long)k * p.N + n0 + nq * 4);
            b_reg[i] = v;
        }
    };
    auto store_chunk = [&](int buf) {
        float* as = As + buf * BK * A_LD;
        float* bs = Bs + buf * BK * B_LD;
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int idx = tid + i * NT;
            const int row = idx >> 2, kq = idx & 3;
            as[(kq * 4 + 0) * A_LD + row] = a_reg[i].x;
            as[(kq * 4 + 1) * A_LD + row] = a_reg[i].y;
            as[(kq * 4 + 2) * A_LD + row] = a_reg[i].z;
            as[(kq * 4 + 3) * A_LD + row] = a_reg[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            const int idx = tid + i * NT;
            const int k = idx / (BN / 4), nq = idx % (BN / 4);
            if (k < BK) *reinterpret_cast<float4*>(bs + k * B_LD + nq * 4) = b_reg[i];
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    load_chunk(0);
    store_chunk(0);
    __syncthreads();

    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        if (chunk + 1 < nchunks) load_chunk(chunk + 1);
        const float* as = As + buf * BK * A_LD;
        const float* bs = Bs + buf * BK * B_LD;
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(as + k * A_LD + ty * 4);
            const float4 a1 = *reinterpret_cast<const float4*>(as + k * A_LD + 64 + ty * 4);
            const float4 b0 = *reinterpret_cast<const float4*>(bs + k * B_LD + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(bs + k * B_LD + BN / 2 + tx * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (chunk + 1 < nchunks) {
            store_chunk(buf ^ 1);
            __syncthreads();
        }
    }

    // ------------------------------ epilogue ------------------------------
    const long long len_b = p.lens ? p.lens[b] : (long long)p.M;
    const int gated = (p.act == ACT_GATED);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = t0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (t >= p.M) continue;
        const bool keep = (long long)t < len_b;
        if (gated) {
            // columns [0,BN/2) of the tile are gates, [BN/2,BN) the matching filters
            const int ng = n0 + tx * 4, nf = n0 + BN / 2 + tx * 4;
            const int co = blockIdx.x * (BN / 2) + tx * 4;
            float r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float g = acc[i][j] * p.alpha + (p.bias ? p.bias[ng + j] : 0.f);
                float f = acc[i][4 + j] * p.alpha + (p.bias ? p.bias[nf + j] : 0.f);
                float v = (1.f / (1.f + expf(-g))) * tanhf(f);
                r[j] = keep ? v * p.out_scale : 0.f;
            }
            *reinterpret_cast<float4*>(p.out + (long long)b * p.out_bstride + (long long)t * p.out_ld + co) =
                make_float4(r[0], r[1], r[2], r[3]);
            continue;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = n0 + h * (BN / 2) + tx * 4;
            if (n >= p.N) continue;
            float r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = acc[i][h * 4 + j] * p.alpha + (p.bias ? p.bias[n + j] : 0.f);
            if (p.aux_out)   // raw conv output (before beta / activation), e.g. the denoiser's F
                *reinterpret_cast<float4*>(p.aux_out + (long long)b * p.aux_bstride + (long long)t * p.aux_ld + n) =
                    make_float4(r[0], r[1], r[2], r[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = act_apply(r[j] * p.beta, p.act, p.act_slope);
            if (p.addvec) {
                const float4 a = *reinterpret_cast<const float4*>(p.addvec + (long long)b * p.addvec_bstride + n);
                r[0] += a.x; r[1] += a.y; r[2] += a.z; r[3] += a.w;
            }
            if (p.res1) {
                const float4 a = *reinterpret_cast<const float4*>(
                    p.res1 + (long long)b * p.res1_bstride + (long long)t * p.res1_ld + n);
                r[0] = fmaf(a.x, p.res1_scale, r[0]); r[1] = fmaf(a.y, p.res1_scale, r[1]);
                r[2] = fmaf(a.z, p.res1_scale, r[2]); r[3] = fmaf(a.w, p.res1_scale, r[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = keep ? r[j] * p.out_scale : 0.f;
            float* o = p.out + (long long)b * p.out_bstride + (long long)t * p.out_ld + n;
            if (p.accumulate) {
                const float4 a = *reinterpret_cast<const float4*>(o);
                r[0] += a.x; r[1] += a.y; r[2] += a.z; r[3] += a.w;
            }
            *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
            if (p.out_h) {
                __half2 h01 = __floats2half2_rn(r[0] > 0.f ? r[0] : r[0] * p.out_h_slope, r[1] > 0.f ? r[1] : r[1] * p.out_h_slope);
                __half2 h23 = __floats2half2_rn(r[2] > 0.f ? r[2] : r[2] * p.out_h_slope, r[3] > 0.f ? r[3] : r[3] * p.out_h_slope);
                uint2 u;
                u.x = *reinterpret_cast<uint32_t*>(&h01);
                u.y = *reinterpret_cast<uint32_t*>(&h23);
                *reinterpret_cast<uint2*>(p.out_h + (long long)b * p.out_bstride + (long long)t * p.out_ld + n) = u;
            }
        }
    }
}

template <int BN>
int launch_bn(const ConvParams& p, cudaStream_t s) {
    constexpr int A_LD = BM + APAD, B_LD = BN + 4;
    const size_t smem = sizeof(float) * (2 * BK * A_LD + 2 * BK * B_LD);
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.B);
    conv1d_simt_kernel<BN><<<grid, 2 * BN, smem, s>>>(p);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// few-rows linear layer: out[m, n] = epi( sum_k x[m, k] w[k, n] ), M <= 32 (per-utterance vectors: step embedding MLP,
// per-layer step / speaker projections).  One CTA per 32 output columns, 8 k-slices per column, x staged in shared
// memory in 64-wide k chunks, fixed-order reduction of the slices (deterministic and independent of M).
// ------------------------------------------------------------------------------------------
namespace {
constexpr int FR_ROWS = 32, FR_COLS = 32, FR_SLICES = 8, FR_KC = 64;

__global__ void __launch_bounds__(FR_COLS * FR_SLICES)
few_rows_linear_kernel(const ConvParams p) {
    __shared__ float s_x[FR_ROWS][FR_KC + 1];
    __shared__ float s_red[FR_ROWS][FR_COLS + 1];
    const int tx = threadIdx.x & 31, ky = threadIdx.x >> 5;
    const int n = blockIdx.x * FR_COLS + tx;
    const bool n_ok = n < p.N;
    // rows are processed in chunks of FR_ROWS (blockIdx.y): a row's result does not depend on how many other rows the
    // call holds, so a 64-utterance batch and its two 32-utterance shards agree bit for bit (multi-GPU global padding)
    const int m0 = blockIdx.y * FR_ROWS;
    const int rows = min(FR_ROWS, p.M - m0);
    float acc[FR_ROWS];
#pragma unroll
    for (int m = 0; m < FR_ROWS; ++m) acc[m] = 0.f;
    for (int k0 = 0; k0 < p.Cin; k0 += FR_KC) {
        for (int i = threadIdx.x; i < FR_ROWS * FR_KC; i += blockDim.x) {
            const int m = i / FR_KC, kk = i - m * FR_KC;
            s_x[m][kk] = (m < rows && k0 + kk < p.Cin) ? p.x[(long long)(m0 + m) * p.x_ld + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < FR_KC / FR_SLICES; ++j) {
            const int kk = ky * (FR_KC / FR_SLICES) + j;
            const float wv = (n_ok && k0 + kk < p.Cin) ? p.w[(long long)(k0 + kk) * p.N + n] : 0.f;
#pragma unroll
            for (int m = 0; m < FR_ROWS; ++m) acc[m] = fmaf(s_x[m][kk], wv, acc[m]);
        }
        __syncthreads();
    }
    for (int r = 0; r < FR_SLICES; ++r) {            // fixed-order reduction of the k slices
        if (ky == r) {
#pragma unroll
            for (int m = 0; m < FR_ROWS; ++m) s_red[m][tx] = (r == 0 ? 0.f : s_red[m][tx]) + acc[m];
        }
        __syncthreads();
    }
    if (ky == 0 && n_ok) {
        for (int m = 0; m < rows; ++m) {
            float v = fmaf(s_red[m][tx], p.alpha, p.bias ? p.bias[n] : 0.f);
            v *= p.beta;
            if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
            if (p.res1) v = fmaf(p.res1[(long long)(m0 + m) * p.res1_ld + n], p.res1_scale, v);
            p.out[(long long)(m0 + m) * p.out_ld + n] = v * p.out_scale;
        }
    }
}
}  // namespace

int launch_conv1d_simt(const ConvParams& p, cudaStream_t s) {
    if (p.few_rows_ok && p.B == 1 && p.M >= 1 && p.M <= FR_ROWS * 1024 && p.taps == 1 && p.shift[0] == 0 && !p.pre_lrelu &&
        (p.act == ACT_NONE || p.act == ACT_RELU) && !p.aux_out && !p.addvec && !p.lens && !p.accumulate && !p.out_h &&
        p.x && p.w && p.out && p.N > 0) {
        few_rows_linear_kernel<<<dim3((p.N + FR_COLS - 1) / FR_COLS, (p.M + FR_ROWS - 1) / FR_ROWS), FR_COLS * FR_SLICES, 0, s>>>(p);
        CMTTS_CHECK_LAUNCH();
        return CMTTS_OK;
    }
    CMTTS_REQUIRE(p.x && p.w && p.out, "conv1d: null pointer");
    CMTTS_REQUIRE(p.Cin % BK == 0, "conv1d: Cin must be a multiple of 16");
    CMTTS_REQUIRE(p.N % 4 == 0 && p.x_ld % 4 == 0 && p.out_ld % 4 == 0, "conv1d: N, x_ld, out_ld must be multiples of 4");
    CMTTS_REQUIRE(p.taps >= 1 && p.taps <= CMTTS_MAX_TAPS, "conv1d: taps out of range");
    CMTTS_REQUIRE(((uintptr_t)p.x % 16 == 0) && ((uintptr_t)p.w % 16 == 0) && ((uintptr_t)p.out % 16 == 0), "conv1d: pointers must be 16-byte aligned");
    CMTTS_REQUIRE(p.B <= 65535 && (p.M + BM - 1) / BM <= 65535, "conv1d: grid too large");
    if (p.B == 0 || p.M == 0 || p.N == 0) return CMTTS_OK;
    if (p.act == ACT_GATED) {
        CMTTS_REQUIRE(p.N % 128 == 0, "conv1d: gated epilogue needs N % 128 == 0 (64 gates + 64 filters per tile)");
        CMTTS_REQUIRE(!p.accumulate && !p.res1 && !p.addvec && !p.aux_out, "conv1d: gated epilogue is plain");
        return launch_bn<128>(p, s);
    }
    if (p.N % 128 == 0 || p.N > 256) return launch_bn<128>(p, s);
    if (p.N % 64 == 0 || p.N > 64) return launch_bn<64>(p, s);
    return launch_bn<32>(p, s);
}
