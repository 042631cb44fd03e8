// Self-attention over phonemes for the FFT blocks (model/blocks.py:303-312 ->
// F.multi_head_attention_forward: q scaled by head_dim**-0.5, key_padding_mask -> -inf, softmax,
// PV; the head-averaged weights the reference also returns are discarded by its caller,
// blocks.py:602, and are not produced here).
//
// Input is the packed in-projection output (B, T, 3C) = [q | k | v]; head h owns channels
// [h*D, (h+1)*D) of each third.  S <= a few hundred phonemes, D = 128: the whole thing is
// ~0.4 GFLOP per layer, so this is a plain fp32 kernel: four query rows per warp, keys streamed
// through shared memory in chunks of 32 (one key per lane for QK^T, conflict-free thanks to the
// +1 row padding), online softmax, P.V accumulated with lane-strided output channels.
#include "common.cuh"
#include <math.h>

namespace {

// WARPS warps per CTA, QW = 4 queries per warp (QT = 4 * WARPS queries per CTA).  The first version gave every warp ONE query:
// per 32-key chunk a warp then issued 256 shared-memory loads for its QK^T (K[lane][d] and the broadcast Q[d]) and 128 for
// P.V — 384 LDS for 160 FMAs per lane, and the kernel sat on the shared-memory pipe (65 us per layer at C2, unchanged when the
// K / V re-reads were halved).  With four queries per warp K[lane][d] and V[j][c] are loaded once and used four times, Q comes
// as one broadcast LDS.128 per d: the same 384 LDS now feed 640 FMAs.  Every query still sees exactly the operation order of
// the one-query version (sequential fmaf over d, 32-key chunks, online softmax, sequential P.V over keys): results are bitwise equal.
constexpr int QW = 4;

template <int D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) attention_kernel(const float* __restrict__ qkv,
                                                               const long long* __restrict__ src_lens,
                                                               float* __restrict__ out, int T, int C) {
    constexpr int DPL = D / 32;
    constexpr int QT = QW * WARPS;
    __shared__ float Ks[32][D + 1];
    __shared__ float Vs[32][D + 1];
    __shared__ float4 Qs[WARPS][D];                    // Qs[w][d] = q[d] of the warp's four queries

    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tq0 = blockIdx.x * QT + warp * QW;       // first query of this warp
    const int len = (int)min((long long)T, src_lens ? src_lens[b] : (long long)T);
    const float* base = qkv + (long long)b * T * 3 * C;
    const float scale = 1.0f / sqrtf((float)D);

#pragma unroll
    for (int i = 0; i < DPL; ++i) {
        const int d = lane + 32 * i;
        float q[QW];
#pragma unroll
        for (int k = 0; k < QW; ++k)
            q[k] = (tq0 + k < T) ? __fmul_rn(base[(long long)(tq0 + k) * 3 * C + h * D + d], scale) : 0.f;
        Qs[warp][d] = make_float4(q[0], q[1], q[2], q[3]);
    }
    float m[QW], l[QW], o[QW][DPL];
#pragma unroll
    for (int k = 0; k < QW; ++k) {
        m[k] = -INFINITY; l[k] = 0.f;
#pragma unroll
        for (int i = 0; i < DPL; ++i) o[k][i] = 0.f;
    }

    for (int k0 = 0; k0 < len; k0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * D; i += WARPS * 32) {
            const int r = i / D, c = i - r * D;
            const int key = k0 + r;
            float kv = 0.f, vv = 0.f;
            if (key < T) {
                kv = base[(long long)key * 3 * C + C + h * D + c];
                vv = base[(long long)key * 3 * C + 2 * C + h * D + c];
            }
            Ks[r][c] = kv;
            Vs[r][c] = vv;
        }
        __syncthreads();
        if (tq0 >= T) continue;
        float s[QW];
#pragma unroll
        for (int k = 0; k < QW; ++k) s[k] = 0.f;
#pragma unroll 8
        for (int d = 0; d < D; ++d) {
            const float kd = Ks[lane][d];
            const float4 q4 = Qs[warp][d];
            s[0] = fmaf(q4.x, kd, s[0]);
            s[1] = fmaf(q4.y, kd, s[1]);
            s[2] = fmaf(q4.z, kd, s[2]);
            s[3] = fmaf(q4.w, kd, s[3]);
        }
        const bool valid = (k0 + lane) < len;
        float p[QW];
#pragma unroll
        for (int k = 0; k < QW; ++k) {
            const float sk = valid ? s[k] : -INFINITY;
            float mx = sk;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m[k], mx);        // finite: key k0 (< len) is valid
            p[k] = valid ? expf(sk - m_new) : 0.f;
            const float corr = expf(m[k] - m_new);      // exp(-inf) = 0 on the first chunk
            float ps = p[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
            l[k] = l[k] * corr + ps;
#pragma unroll
            for (int i = 0; i < DPL; ++i) o[k][i] *= corr;
            m[k] = m_new;
        }
        for (int j = 0; j < 32; ++j) {
            float v[DPL];
#pragma unroll
            for (int i = 0; i < DPL; ++i) v[i] = Vs[j][lane + 32 * i];
#pragma unroll
            for (int k = 0; k < QW; ++k) {
                const float pj = __shfl_sync(0xffffffffu, p[k], j);
#pragma unroll
                for (int i = 0; i < DPL; ++i) o[k][i] = fmaf(pj, v[i], o[k][i]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < QW; ++k) {
        if (tq0 + k < T) {
            const float inv = len > 0 ? 1.f / l[k] : 0.f;
#pragma unroll
            for (int i = 0; i < DPL; ++i)
                out[((long long)b * T + tq0 + k) * C + h * D + lane + 32 * i] = o[k][i] * inv;
        }
    }
}

}  // namespace

int launch_attention(const float* qkv, const long long* src_lens, float* out, int B, int T, int C,
                     int heads, cudaStream_t s) {
    if (B == 0 || T == 0) return CMTTS_OK;
    CMTTS_REQUIRE(heads > 0 && C % heads == 0, "attention: C % heads");
    const int D = C / heads;
    // 16 queries per CTA (4 warps) when that still gives every SM a CTA, 8 (2 warps) for the small batches
    const bool wide = (long long)((T + 15) / 16) * heads * B >= 148;
    const int QT = wide ? 16 : 8;
    dim3 grid((T + QT - 1) / QT, heads, B);
#define CMTTS_ATT(D_) do { if (wide) attention_kernel<D_, 4><<<grid, 4 * 32, 0, s>>>(qkv, src_lens, out, T, C); \
                           else attention_kernel<D_, 2><<<grid, 2 * 32, 0, s>>>(qkv, src_lens, out, T, C); } while (0)
    if (D == 128) CMTTS_ATT(128);
    else if (D == 64) CMTTS_ATT(64);
    else if (D == 32) CMTTS_ATT(32);
    else { cmtts_set_error("attention: head_dim must be 32, 64 or 128", __FILE__, __LINE__); return CMTTS_ERR_UNSUPPORTED; }
#undef CMTTS_ATT
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}
