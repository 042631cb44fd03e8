// Self-attention over phonemes for the FFT blocks (model/blocks.py:303-312 ->
// F.multi_head_attention_forward: q scaled by head_dim**-0.5, key_padding_mask -> -inf, softmax,
// PV; the head-averaged weights the reference also returns are discarded by its caller,
// blocks.py:602, and are not produced here).
//
// Input is the packed in-projection output (B, T, 3C) = [q | k | v]; head h owns channels
// [h*D, (h+1)*D) of each third.  S <= a few hundred phonemes, D = 128: the whole thing is
// ~0.4 GFLOP per layer, so this is a plain fp32 kernel: one warp per query row, keys streamed
// through shared memory in chunks of 32 (one key per lane for QK^T, conflict-free thanks to the
// +1 row padding), online softmax, P.V accumulated with lane-strided output channels.
#include "common.cuh"
#include <math.h>

namespace {

// QT = queries (warps) per CTA.  Every CTA streams ALL keys and values of its (utterance, head) through shared memory, so
// the K / V traffic is proportional to the number of CTAs: 16 queries per CTA when that still leaves >= one CTA per SM
// (C2: 960 -> 512 CTAs, half the re-reads), 8 otherwise.  A warp's arithmetic does not depend on QT: results are bitwise equal.
template <int D, int QT>
__global__ void __launch_bounds__(QT * 32) attention_kernel(const float* __restrict__ qkv,
                                                            const long long* __restrict__ src_lens,
                                                            float* __restrict__ out, int T, int C) {
    constexpr int DPL = D / 32;
    __shared__ float Ks[32][D + 1];
    __shared__ float Vs[32][D + 1];
    __shared__ float Qs[QT][D];

    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tq = blockIdx.x * QT + warp;
    const int len = (int)min((long long)T, src_lens ? src_lens[b] : (long long)T);
    const float* base = qkv + (long long)b * T * 3 * C;
    const float scale = 1.0f / sqrtf((float)D);

    if (tq < T) {
#pragma unroll
        for (int i = 0; i < DPL; ++i)
            Qs[warp][lane + 32 * i] = __fmul_rn(base[(long long)tq * 3 * C + h * D + lane + 32 * i], scale);
    }
    float m = -INFINITY, l = 0.f;
    float o[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[i] = 0.f;

    for (int k0 = 0; k0 < len; k0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * D; i += QT * 32) {
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
        if (tq >= T) continue;
        float s = 0.f;
#pragma unroll 8
        for (int d = 0; d < D; ++d) s = fmaf(Qs[warp][d], Ks[lane][d], s);
        const bool valid = (k0 + lane) < len;
        s = valid ? s : -INFINITY;
        float mx = s;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        const float m_new = fmaxf(m, mx);           // finite: key k0 (< len) is valid
        const float p = valid ? expf(s - m_new) : 0.f;
        const float corr = expf(m - m_new);         // exp(-inf) = 0 on the first chunk
        float ps = p;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
        l = l * corr + ps;
#pragma unroll
        for (int i = 0; i < DPL; ++i) o[i] *= corr;
        for (int j = 0; j < 32; ++j) {
            const float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
            for (int i = 0; i < DPL; ++i) o[i] = fmaf(pj, Vs[j][lane + 32 * i], o[i]);
        }
        m = m_new;
    }
    if (tq < T) {
        const float inv = len > 0 ? 1.f / l : 0.f;
#pragma unroll
        for (int i = 0; i < DPL; ++i)
            out[((long long)b * T + tq) * C + h * D + lane + 32 * i] = o[i] * inv;
    }
}

}  // namespace

int launch_attention(const float* qkv, const long long* src_lens, float* out, int B, int T, int C,
                     int heads, cudaStream_t s) {
    if (B == 0 || T == 0) return CMTTS_OK;
    CMTTS_REQUIRE(heads > 0 && C % heads == 0, "attention: C % heads");
    const int D = C / heads;
    const bool wide = (long long)((T + 15) / 16) * heads * B >= 148;
    const int QT = wide ? 16 : 8;
    dim3 grid((T + QT - 1) / QT, heads, B);
#define CMTTS_ATT(D_) do { if (wide) attention_kernel<D_, 16><<<grid, 16 * 32, 0, s>>>(qkv, src_lens, out, T, C); \
                           else attention_kernel<D_, 8><<<grid, 8 * 32, 0, s>>>(qkv, src_lens, out, T, C); } while (0)
    if (D == 128) CMTTS_ATT(128);
    else if (D == 64) CMTTS_ATT(64);
    else if (D == 32) CMTTS_ATT(32);
    else { cmtts_set_error("attention: head_dim must be 32, 64 or 128", __FILE__, __LINE__); return CMTTS_ERR_UNSUPPORTED; }
#undef CMTTS_ATT
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}
