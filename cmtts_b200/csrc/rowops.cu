// Row-wise (HBM-bound) kernels of the hot path: LayerNorm, predictor heads, embedding gathers,
// positional tables, quantisers, duration scan + length-regulator gather, inverse-CWT pitch,
// Karras re-noise, layout transposes and the HiFi-GAN output stage.  All fp32 / int64, warp-shuffle
// reductions, 128-bit vectors where the layout allows it.  Each function cites the reference code it
// reproduces (paths relative to the reference repo).
#include "common.cuh"
#include <math.h>

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the channel axis, one warp per (b, t) row.  torch.nn.LayerNorm semantics
// (biased variance, eps inside the sqrt): model/blocks.py:88-107 (eps 1e-12), modules.py:74 (1e-5).
// Rows t >= lens[b] are written as zeros (the reference multiplies by the non-padding mask right
// after every norm on the token path: modules.py:99, :502-503).
// ---------------------------------------------------------------------------------------------
template <int NPL>
__device__ __forceinline__ void ln_row(const float* __restrict__ xr, const float* __restrict__ w,
                                       const float* __restrict__ b, float eps, int lane, float (&y)[NPL]) {
    constexpr int C = NPL * 32;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) { y[i] = xr[lane + 32 * i]; s += y[i]; }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) { const float d = y[i] - mean; q = fmaf(d, d, q); }
    const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / C) + eps);
#pragma unroll
    for (int i = 0; i < NPL; ++i) y[i] = (y[i] - mean) * rstd * w[lane + 32 * i] + b[lane + 32 * i];
}

template <int NPL>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ b, float eps, float* __restrict__ out,
                                 int rows, int T, const long long* __restrict__ lens) {
    constexpr int C = NPL * 32;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* o = out + (long long)row * C;
    if (lens && (row % T) >= lens[row / T]) {
#pragma unroll
        for (int i = 0; i < NPL; ++i) o[lane + 32 * i] = 0.f;
        return;
    }
    float y[NPL];
    ln_row<NPL>(x + (long long)row * C, w, b, eps, lane, y);
#pragma unroll
    for (int i = 0; i < NPL; ++i) o[lane + 32 * i] = y[i];
}

// LayerNorm followed by a narrow Linear head (odim <= 16): the tail of DurationPredictor /
// PitchPredictor / EnergyPredictor (model/modules.py:505-509, :554-555).  With `lens`, the
// normalised row is zeroed for padded tokens before the head and the head output is zeroed too
// (DurationPredictor masks after every layer and after the Linear, modules.py:502-506), so a
// padded token yields exactly 0 (bias included), like the reference.
template <int NPL>
__global__ void ln_head_kernel(const float* __restrict__ x, const float* __restrict__ lw,
                               const float* __restrict__ lb, float eps, const float* __restrict__ hw,
                               const float* __restrict__ hb, int odim, float scale, float* __restrict__ out,
                               int rows, int T, const long long* __restrict__ lens) {
    constexpr int C = NPL * 32;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    if (lens && (row % T) >= lens[row / T]) {
        if (lane < odim) out[(long long)row * odim + lane] = 0.f;
        return;
    }
    float y[NPL];
    ln_row<NPL>(x + (long long)row * C, lw, lb, eps, lane, y);
    for (int o = 0; o < odim; ++o) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NPL; ++i) s = fmaf(y[i], hw[o * C + lane + 32 * i], s);
        s = warp_sum(s);
        if (lane == 0) out[(long long)row * odim + o] = __fmul_rn(s + hb[o], scale);
    }
}

// ---------------------------------------------------------------------------------------------
// Token embedding + sinusoidal positions: FastspeechEncoder.forward_embedding
// (model/modules.py:145-151), positions = make_positions(tokens, 0) (utils/tools.py:810-822):
// 1-based running count of non-pad tokens, 0 for pad.  One CTA per utterance; warp 0 scans.
// The positional table is built on the host exactly as model/blocks.py:44-60 does.
// ---------------------------------------------------------------------------------------------
__global__ void embed_tokens_kernel(const long long* __restrict__ tokens, const float* __restrict__ emb,
                                    const float* __restrict__ pe, int pe_rows, float emb_scale,
                                    float* __restrict__ out, int T, int C, const long long* __restrict__ lens) {
    extern __shared__ int s_pos[];
    const int b = blockIdx.x;
    const long long* tok = tokens + (long long)b * T;
    if (threadIdx.x < 32) {
        int running = 0;
        for (int base = 0; base < T; base += 32) {
            const int t = base + threadIdx.x;
            const bool f = t < T && tok[t] != 0;
            const unsigned m = __ballot_sync(0xffffffffu, f);
            if (t < T) s_pos[t] = f ? running + __popc(m & ((1u << threadIdx.x) - 1u)) + 1 : 0;
            running += __popc(m);
        }
    }
    __syncthreads();
    const long long len = lens ? lens[b] : (long long)T;
    const int c4 = C >> 2;
    for (int i = threadIdx.x; i < T * c4; i += blockDim.x) {
        const int t = i / c4, c = (i - t * c4) << 2;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < len) {
            const int pos = min(s_pos[t], pe_rows - 1);
            const float4 e = *reinterpret_cast<const float4*>(emb + tok[t] * C + c);
            const float4 p = *reinterpret_cast<const float4*>(pe + (long long)pos * C + c);
            r.x = __fadd_rn(__fmul_rn(emb_scale, e.x), p.x);
            r.y = __fadd_rn(__fmul_rn(emb_scale, e.y), p.y);
            r.z = __fadd_rn(__fmul_rn(emb_scale, e.z), p.z);
            r.w = __fadd_rn(__fmul_rn(emb_scale, e.w), p.w);
        }
        *reinterpret_cast<float4*>(out + ((long long)b * T + t) * C + c) = r;
    }
}

__global__ void add_rowvec_kernel(float* __restrict__ x, const float* __restrict__ vec, int T, int C,
                                  long long total4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c4 = C >> 2;
    const long long row = i / c4;
    const int c = (int)(i - row * c4) << 2;
    const long long b = row / T;
    float4 v = *reinterpret_cast<float4*>(x + row * C + c);
    const float4 a = *reinterpret_cast<const float4*>(vec + b * C + c);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    *reinterpret_cast<float4*>(x + row * C + c) = v;
}

// PitchPredictor.forward prologue (model/modules.py:548-549): xs + alpha * PE[pos], where pos is the
// running count of rows whose FIRST channel is non-zero (make_positions on xs[..., 0]).
__global__ void add_positional_kernel(const float* __restrict__ x, const float* __restrict__ pe, int pe_rows,
                                      const float* __restrict__ alpha, float* __restrict__ out, int T, int C) {
    extern __shared__ int s_pos[];
    const int b = blockIdx.x;
    const float* xb = x + (long long)b * T * C;
    if (threadIdx.x < 32) {
        int running = 0;
        for (int base = 0; base < T; base += 32) {
            const int t = base + threadIdx.x;
            const bool f = t < T && xb[(long long)t * C] != 0.f;
            const unsigned m = __ballot_sync(0xffffffffu, f);
            if (t < T) s_pos[t] = f ? running + __popc(m & ((1u << threadIdx.x) - 1u)) + 1 : 0;
            running += __popc(m);
        }
    }
    __syncthreads();
    const float a = alpha[0];
    const int c4 = C >> 2;
    for (int i = threadIdx.x; i < T * c4; i += blockDim.x) {
        const int t = i / c4, c = (i - t * c4) << 2;
        const int pos = min(s_pos[t], pe_rows - 1);
        const float4 v = *reinterpret_cast<const float4*>(xb + (long long)t * C + c);
        const float4 p = *reinterpret_cast<const float4*>(pe + (long long)pos * C + c);
        float4 r;
        r.x = __fadd_rn(v.x, __fmul_rn(a, p.x));
        r.y = __fadd_rn(v.y, __fmul_rn(a, p.y));
        r.z = __fadd_rn(v.z, __fmul_rn(a, p.z));
        r.w = __fadd_rn(v.w, __fmul_rn(a, p.w));
        *reinterpret_cast<float4*>(out + ((long long)b * T + t) * C + c) = r;
    }
}

// get_energy_embedding (model/modules.py:319-329): prediction * control -> torch.bucketize(.,
// bins) (right=False: number of boundaries strictly below the value) -> embedding row added to x.
__global__ void energy_embed_kernel(const float* __restrict__ x, const float* __restrict__ pred, float control,
                                    const float* __restrict__ bins, int nbins, const float* __restrict__ emb,
                                    float* __restrict__ out, long long* __restrict__ idx_out,
                                    float* __restrict__ pred_out, int rows, int C) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float v = __fmul_rn(pred[row], control);
    int lo = 0, hi = nbins;  // first index with bins[i] >= v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (bins[mid] < v) lo = mid + 1; else hi = mid;
    }
    if (lane == 0) {
        if (idx_out) idx_out[row] = lo;
        if (pred_out) pred_out[row] = v;
    }
    const float* e = emb + (long long)lo * C;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(x + (long long)row * C + c);
        const float4 g = *reinterpret_cast<const float4*>(e + c);
        *reinterpret_cast<float4*>(out + (long long)row * C + c) = make_float4(a.x + g.x, a.y + g.y, a.z + g.z, a.w + g.w);
    }
}

// Duration rounding + prefix scans, one warp per utterance.
//   d = clamp(round(exp(log_d) - 1) * d_control, min=0)           model/modules.py:369-372
//   cumsum[:,0,:] : scan of max(int(d), 0)   (LengthRegulator.expand, modules.py:439-441)
//   cumsum[:,1,:] : scan of round(d) * (t < src_len)   (dur_to_mel2ph, utils/tools.py:785-790)
//   mel_lens = cumsum[:,0,-1]
__global__ void round_durations_kernel(const float* __restrict__ log_d, float d_control,
                                       const long long* __restrict__ src_lens, float* __restrict__ d_rounded,
                                       long long* __restrict__ cumsum, long long* __restrict__ mel_lens, int T) {
    const int b = blockIdx.x, lane = threadIdx.x;
    long long run0 = 0, run1 = 0;
    const long long len = src_lens[b];
    for (int base = 0; base < T; base += 32) {
        const int t = base + lane;
        long long a = 0, m = 0;
        if (t < T) {
            float d = __fmul_rn(rintf(__fsub_rn(expf(log_d[(long long)b * T + t]), 1.f)), d_control);
            d = fmaxf(d, 0.f);
            d_rounded[(long long)b * T + t] = d;
            a = (long long)d;                         // int() truncation
            m = (t < len) ? (long long)rintf(d) : 0;  // torch.round(dur).long() * (1 - pad)
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long ua = __shfl_up_sync(0xffffffffu, a, o);
            const long long um = __shfl_up_sync(0xffffffffu, m, o);
            if (lane >= o) { a += ua; m += um; }
        }
        if (t < T) {
            cumsum[((long long)b * 2 + 0) * T + t] = run0 + a;
            cumsum[((long long)b * 2 + 1) * T + t] = run1 + m;
        }
        run0 += __shfl_sync(0xffffffffu, a, 31);
        run1 += __shfl_sync(0xffffffffu, m, 31);
    }
    if (lane == 0) mel_lens[b] = run0;
}

// LengthRegulator (model/modules.py:421-444 + pad utils/tools.py:724-742) as a gather:
// frame t of utterance b copies token row upper_bound(cumsum_b, t); frames >= mel_len are zero.
// Also emits mel2ph (utils/tools.py:768-798): 1-based token index, 0 for padding frames.
__global__ void length_regulate_kernel(const float* __restrict__ x, const long long* __restrict__ cumsum,
                                       const long long* __restrict__ mel_lens, float* __restrict__ out,
                                       long long* __restrict__ mel2ph, int T, int L, int C) {
    extern __shared__ int s_cs[];  // [2][T]
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) s_cs[i] = (int)cumsum[(long long)b * 2 * T + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int mel_len = (int)mel_lens[b];
    const int tot2 = T > 0 ? s_cs[2 * T - 1] : 0;
    const int t_end = min(L, (blockIdx.x + 1) * 64);
    for (int t = blockIdx.x * 64 + warp; t < t_end; t += nw) {
        float* o = out + ((long long)b * L + t) * C;
        if (t < mel_len) {
            int lo = 0, hi = T;  // first i with cs[i] > t
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_cs[mid] <= t) lo = mid + 1; else hi = mid; }
            const float* src = x + ((long long)b * T + lo) * C;
            for (int c = lane * 4; c < C; c += 128)
                *reinterpret_cast<float4*>(o + c) = *reinterpret_cast<const float4*>(src + c);
        } else {
            for (int c = lane * 4; c < C; c += 128) *reinterpret_cast<float4*>(o + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (mel2ph && lane == 0) {
            long long v = 0;
            if (t < tot2) {
                int lo = 0, hi = T;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_cs[T + mid] <= t) lo = mid + 1; else hi = mid; }
                v = lo + 1;
            }
            mel2ph[(long long)b * L + t] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// inverse CWT -> f0 -> coarse pitch bins -> + pitch embedding  (VarianceAdaptor.get_pitch_embedding
// cwt branch model/modules.py:273-307; cwt2f0_norm / cwt2f0 / inverse_cwt_torch / norm_f0 /
// denorm_f0 / f0_to_coarse in utils/pitch_tools.py:244-279, :38-78, :26-35).
// Kernel 1: per utterance, rec[t] = sum_k cwt[t,k] * (k+3.5)^-2.5 over ALL L (padded) frames, mean
// and unbiased std.  Kernel 2: per frame.
// ---------------------------------------------------------------------------------------------
__global__ void cwt_stats_kernel(const float* __restrict__ cwt, int cwt_dim, const float* __restrict__ wb,
                                 float* __restrict__ rec, float* __restrict__ stat, int L) {
    __shared__ double s_red[32];
    __shared__ double s_mean;
    const int b = blockIdx.x;
    const float* cb = cwt + (long long)b * L * cwt_dim;
    float* rb = rec + (long long)b * L;
    double s = 0.0;
    for (int t = threadIdx.x; t < L; t += blockDim.x) {
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < 10; ++k) r = __fadd_rn(r, __fmul_rn(cb[(long long)t * cwt_dim + k], wb[k]));
        rb[t] = r;
        s += (double)r;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double a = 0; for (int i = 0; i < nw; ++i) a += s_red[i]; s_mean = a / (double)L; }
    __syncthreads();
    const float mean = (float)s_mean;
    double q = 0.0;
    for (int t = threadIdx.x; t < L; t += blockDim.x) { const double d = (double)__fsub_rn(rb[t], mean); q += d * d; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0; for (int i = 0; i < nw; ++i) a += s_red[i];
        stat[b * 2 + 0] = mean;
        stat[b * 2 + 1] = (float)sqrt(a / (double)(L - 1));  // unbiased; L == 1 -> NaN like torch.std
    }
}

struct F0Consts { float mel_min, mel_span, bins_m2; };

__global__ void cwt_pitch_kernel(const float* __restrict__ cwt, int cwt_dim, const float* __restrict__ rec,
                                 const float* __restrict__ stat, const float* __restrict__ f0stats, int f0s_ld,
                                 float std_scale, float eps, int use_uv, F0Consts fc,
                                 const float* __restrict__ xf, const float* __restrict__ pitch_emb, int n_pitch,
                                 float* __restrict__ cond, float* __restrict__ f0_denorm,
                                 long long* __restrict__ pitch_idx, int L, int C, long long rows) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = (int)(row / L);
    const float r = __fdiv_rn(__fsub_rn(rec[row], stat[b * 2]), stat[b * 2 + 1]);
    const float mean = f0stats[b * f0s_ld], sd = __fmul_rn(f0stats[b * f0s_ld + 1], std_scale);
    float f0 = expf(__fadd_rn(__fmul_rn(r, sd), mean));   // cwt2f0: f0 * std + mean, exp
    f0 = log2f(__fadd_rn(f0, eps));                         // norm_f0 (pitch_norm == "log")
    f0 = exp2f(f0);                                         // denorm_f0: 2 ** f0
    if (use_uv && cwt[row * cwt_dim + cwt_dim - 1] > 0.f) f0 = 0.f;
    // f0_to_coarse, op by op in fp32 like the reference
    float m = __fmul_rn(1127.f, logf(__fadd_rn(1.f, __fdiv_rn(f0, 700.f))));
    if (m > 0.f) m = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(m, fc.mel_min), fc.bins_m2), fc.mel_span), 1.f);
    if (m <= 1.f) m = 1.f;
    if (m > 255.f) m = 255.f;
    long long idx = (long long)__fadd_rn(m, 0.5f);
    if (!(idx >= 0)) idx = 0;             // NaN (L == 1) -> keep the gather in range
    if (idx >= n_pitch) idx = n_pitch - 1;
    if (lane == 0) { f0_denorm[row] = f0; pitch_idx[row] = idx; }
    const float* e = pitch_emb + idx * C;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(xf + row * C + c);
        const float4 g = *reinterpret_cast<const float4*>(e + c);
        *reinterpret_cast<float4*>(cond + row * C + c) = make_float4(a.x + g.x, a.y + g.y, a.z + g.z, a.w + g.w);
    }
}

// DiffusionEmbedding.forward (model/blocks.py:633-640): [sin(t f_k) | cos(t f_k)], f_k from the host
__global__ void step_sinusoid_kernel(const float* __restrict__ t, const float* __restrict__ freq,
                                     float* __restrict__ out, int B, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = C >> 1;
    if (i >= B * half) return;
    const int b = i / half, k = i - b * half;
    const float a = __fmul_rn(t[b], freq[k]);
    out[b * C + k] = sinf(a);
    out[b * C + half + k] = cosf(a);
}

// Mish (model/blocks.py:621-623): x * tanh(softplus(x)), softplus with torch's threshold 20
__global__ void mish_kernel(float* __restrict__ x, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    const float sp = v > 20.f ? v : log1pf(expf(v));
    x[i] = v * tanhf(sp);
}

// stochastic_iterative_sampler re-noise (karras_diffusion.py:852): x0 + (noise * s1) * s2
__global__ void renoise_kernel(const float4* __restrict__ x0, const float4* __restrict__ noise, float s1, float s2,
                               float4* __restrict__ out, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 a = x0[i], z = noise[i];
    float4 r;
    r.x = __fadd_rn(a.x, __fmul_rn(__fmul_rn(z.x, s1), s2));
    r.y = __fadd_rn(a.y, __fmul_rn(__fmul_rn(z.y, s1), s2));
    r.z = __fadd_rn(a.z, __fmul_rn(__fmul_rn(z.z, s1), s2));
    r.w = __fadd_rn(a.w, __fmul_rn(__fmul_rn(z.w, s1), s2));
    out[i] = r;
}

// per-(utterance, layer) additive constants of the denoiser's y-recurrence (pipeline.cu, cmtts_denoiser_forward_tc):
//   yc[b][l][n] = r * ds[b][l][n] + dsp[b][l+1][n] - r * dsp[b][l][n],  l < layers - 1
__global__ void dn_fuse_steps_kernel(const float* __restrict__ ds, const float* __restrict__ dsp, float* __restrict__ yc,
                                     int layers, int C, float r, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int per_b = (layers - 1) * C;
    const long long b = i / per_b;
    const int j = (int)(i - b * per_b);                 // l * C + n
    const long long src = b * (long long)layers * C + j;
    yc[i] = r * ds[src] + dsp[src + C] - r * dsp[src];
}

__global__ void scale_kernel(const float4* __restrict__ x, float a, float4* __restrict__ out, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = x[i];
    out[i] = make_float4(v.x * a, v.y * a, v.z * a, v.w * a);
}

// (B, C, L) -> (B, L, C) through a padded 32x32 shared tile (coalesced both ways)
__global__ void transpose_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int L) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* xb = x + (long long)b * C * L;
    float* ob = out + (long long)b * C * L;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, l = l0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && l < L) ? xb[(long long)c * L + l] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int l = l0 + i, c = c0 + threadIdx.x;
        if (l < L && c < C) ob[(long long)l * C + c] = tile[threadIdx.x][i];
    }
}

// HiFi-GAN output stage (hifigan/models.py:160-163 + utils/model.py:195-198):
//   x = xs / num_kernels ; lrelu(x, 0.01) ; conv_post (C -> 1, k taps, SAME) ; tanh ;
//   int16 = (wav * 32768).astype(int16)   (C truncation toward zero, wrap like numpy on x86)
__global__ void conv_post_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ bias, float pre_slope, float pre_div,
                                 float* __restrict__ wav, short* __restrict__ wav_i16, float max_wav,
                                 int L, int C, int K) {
    extern __shared__ float s_w[];  // [K][C]
    for (int i = threadIdx.x; i < K * C; i += blockDim.x) s_w[i] = w[i];
    __syncthreads();
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= L) return;
    const float* xb = x + (long long)b * L * C;
    float acc = 0.f;
    const int pad = (K - 1) / 2;
    for (int k = 0; k < K; ++k) {
        const int src = n + k - pad;
        if (src < 0 || src >= L) continue;
        const float4* xr = reinterpret_cast<const float4*>(xb + (long long)src * C);
        const float* wk = s_w + k * C;
        for (int c4 = 0; c4 < (C >> 2); ++c4) {
            float4 v = xr[c4];
            v.x = __fdiv_rn(v.x, pre_div); v.y = __fdiv_rn(v.y, pre_div);
            v.z = __fdiv_rn(v.z, pre_div); v.w = __fdiv_rn(v.w, pre_div);
            v.x = v.x > 0.f ? v.x : v.x * pre_slope; v.y = v.y > 0.f ? v.y : v.y * pre_slope;
            v.z = v.z > 0.f ? v.z : v.z * pre_slope; v.w = v.w > 0.f ? v.w : v.w * pre_slope;
            acc = fmaf(v.x, wk[c4 * 4 + 0], acc); acc = fmaf(v.y, wk[c4 * 4 + 1], acc);
            acc = fmaf(v.z, wk[c4 * 4 + 2], acc); acc = fmaf(v.w, wk[c4 * 4 + 3], acc);
        }
    }
    const float y = tanhf(acc + bias[0]);
    if (wav) wav[(long long)b * L + n] = y;
    if (wav_i16) wav_i16[(long long)b * L + n] = (short)(int)__fmul_rn(y, max_wav);
}

}  // namespace

// ------------------------------------ launchers ------------------------------------
#define DISPATCH_NPL(C, CALL)                                      \
    switch (C) {                                                   \
        case 64: { constexpr int NPL = 2; CALL; } break;           \
        case 128: { constexpr int NPL = 4; CALL; } break;          \
        case 256: { constexpr int NPL = 8; CALL; } break;          \
        case 384: { constexpr int NPL = 12; CALL; } break;         \
        case 512: { constexpr int NPL = 16; CALL; } break;         \
        case 1024: { constexpr int NPL = 32; CALL; } break;        \
        default: cmtts_set_error("layernorm: unsupported channel count", __FILE__, __LINE__); return CMTTS_ERR_UNSUPPORTED; \
    }

int launch_layernorm(const float* x, const float* w, const float* b, float eps, float* out,
                     int B, int T, int C, const long long* lens, cudaStream_t s) {
    const int rows = B * T;
    if (rows == 0) return CMTTS_OK;
    const int grid = (rows + 7) / 8;
    DISPATCH_NPL(C, (layernorm_kernel<NPL><<<grid, 256, 0, s>>>(x, w, b, eps, out, rows, T, lens)));
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_ln_head(const float* x, const float* lw, const float* lb, float eps, const float* hw,
                   const float* hb, int odim, float scale, float* out, int B, int T, int C,
                   const long long* lens, cudaStream_t s) {
    const int rows = B * T;
    if (rows == 0) return CMTTS_OK;
    CMTTS_REQUIRE(odim >= 1 && odim <= 32, "ln_head: odim must be in [1, 32]");
    const int grid = (rows + 7) / 8;
    DISPATCH_NPL(C, (ln_head_kernel<NPL><<<grid, 256, 0, s>>>(x, lw, lb, eps, hw, hb, odim, scale, out, rows, T, lens)));
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_embed_tokens(const long long* tokens, const float* emb, const float* pe, int pe_rows,
                        float emb_scale, float* out, int B, int T, int C, const long long* lens,
                        cudaStream_t s) {
    if (B == 0 || T == 0) return CMTTS_OK;
    CMTTS_REQUIRE(C % 4 == 0, "embed_tokens: C % 4");
    CMTTS_REQUIRE(pe_rows > T, "embed_tokens: positional table too short");
    CMTTS_REQUIRE((size_t)T * 4 <= 200 * 1024, "embed_tokens: T too large");
    if ((size_t)T * 4 > 48 * 1024)
        cudaFuncSetAttribute(embed_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T * 4);
    embed_tokens_kernel<<<B, 256, T * sizeof(int), s>>>(tokens, emb, pe, pe_rows, emb_scale, out, T, C, lens);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_add_rowvec(float* x, const float* vec, int B, int T, int C, cudaStream_t s) {
    const long long total4 = (long long)B * T * C / 4;
    if (total4 == 0) return CMTTS_OK;
    add_rowvec_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, s>>>(x, vec, T, C, total4);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_add_positional(const float* x, const float* pe, int pe_rows, const float* alpha,
                          float* out, int B, int T, int C, cudaStream_t s) {
    if (B == 0 || T == 0) return CMTTS_OK;
    CMTTS_REQUIRE(pe_rows > T, "add_positional: positional table too short");
    CMTTS_REQUIRE((size_t)T * 4 <= 200 * 1024, "add_positional: T too large");
    if ((size_t)T * 4 > 48 * 1024)
        cudaFuncSetAttribute(add_positional_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T * 4);
    add_positional_kernel<<<B, 512, T * sizeof(int), s>>>(x, pe, pe_rows, alpha, out, T, C);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_energy_embed(const float* x, const float* pred, float control, const float* bins, int nbins,
                        const float* emb, float* out, long long* idx_out, float* pred_out,
                        int B, int T, int C, cudaStream_t s) {
    const int rows = B * T;
    if (rows == 0) return CMTTS_OK;
    energy_embed_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, pred, control, bins, nbins, emb, out, idx_out, pred_out, rows, C);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_round_durations(const float* log_d, float d_control, const long long* src_lens,
                           float* d_rounded, long long* cumsum, long long* mel_lens,
                           int B, int T, cudaStream_t s) {
    if (B == 0) return CMTTS_OK;
    round_durations_kernel<<<B, 32, 0, s>>>(log_d, d_control, src_lens, d_rounded, cumsum, mel_lens, T);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_length_regulate(const float* x, const long long* cumsum, const long long* mel_lens,
                           float* out, long long* mel2ph, int B, int T, int L, int C, cudaStream_t s) {
    if (B == 0 || L == 0) return CMTTS_OK;
    CMTTS_REQUIRE(C % 4 == 0, "length_regulate: C % 4");
    const size_t smem = (size_t)2 * T * sizeof(int);
    CMTTS_REQUIRE(smem <= 200 * 1024, "length_regulate: T too large");
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(length_regulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((L + 63) / 64, B);
    if (g_cmtts_prof_on)   // read (N_tok C + N_tok) 4 + write N_frm (C 4 + 8)
        cmtts_prof_note("length_regulate (scan + gather)", 0.0, ((double)B * T * C + (double)B * T) * 4.0 + (double)B * L * (C * 4.0 + 8.0));
    length_regulate_kernel<<<grid, 256, smem, s>>>(x, cumsum, mel_lens, out, mel2ph, T, L, C);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int cmtts_cwt_pitch_impl(const float* cwt, int cwt_dim, const float* cwt_b, const float* f0stats, int f0s_ld,
                         float std_scale, float eps, int use_uv, float mel_min, float mel_span,
                         const float* xf, const float* pitch_emb, int n_pitch, float* cond,
                         float* f0_denorm, long long* pitch_idx, float* rec_scratch, float* stat_scratch,
                         int B, int L, int C, cudaStream_t s) {
    if (B == 0 || L == 0) return CMTTS_OK;
    CMTTS_REQUIRE(cwt_dim >= 10, "cwt_pitch: cwt_dim < 10");
    cwt_stats_kernel<<<B, 256, 0, s>>>(cwt, cwt_dim, cwt_b, rec_scratch, stat_scratch, L);
    CMTTS_CHECK_LAUNCH();
    const long long rows = (long long)B * L;
    F0Consts fc{mel_min, mel_span, 254.f};
    cwt_pitch_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(cwt, cwt_dim, rec_scratch, stat_scratch, f0stats, f0s_ld,
                                                                  std_scale, eps, use_uv, fc, xf, pitch_emb, n_pitch,
                                                                  cond, f0_denorm, pitch_idx, L, C, rows);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int cmtts_step_sinusoid_impl(const float* t, const float* freq, float* out, int B, int C, cudaStream_t s) {
    const int n = B * (C / 2);
    if (n == 0) return CMTTS_OK;
    step_sinusoid_kernel<<<(n + 127) / 128, 128, 0, s>>>(t, freq, out, B, C);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_mish(float* x, long long n, cudaStream_t s) {
    if (n == 0) return CMTTS_OK;
    mish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_renoise(const float* x0, const float* noise, float s1, float s2, float* out, long long n,
                   cudaStream_t s) {
    if (n == 0) return CMTTS_OK;
    CMTTS_REQUIRE(n % 4 == 0, "renoise: n % 4");
    const long long n4 = n / 4;
    if (g_cmtts_prof_on) cmtts_prof_note("renoise (x0 + noise s)", 2.0 * n, 12.0 * n);
    renoise_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>((const float4*)x0, (const float4*)noise, s1, s2, (float4*)out, n4);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_dn_fuse_steps(const float* ds_all, const float* dsp_all, float* yc, int B, int layers, int C, float r,
                         cudaStream_t s) {
    const long long n = (long long)B * (layers - 1) * C;
    if (n <= 0) return CMTTS_OK;
    dn_fuse_steps_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ds_all, dsp_all, yc, layers, C, r, n);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_scale(const float* x, float a, float* out, long long n, cudaStream_t s) {
    if (n == 0) return CMTTS_OK;
    CMTTS_REQUIRE(n % 4 == 0, "scale: n % 4");
    const long long n4 = n / 4;
    scale_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>((const float4*)x, a, (float4*)out, n4);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_transpose_bcl_to_blc(const float* x, float* out, int B, int C, int L, cudaStream_t s) {
    if (B == 0 || C == 0 || L == 0) return CMTTS_OK;
    dim3 grid((L + 31) / 32, (C + 31) / 32, B), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(x, out, C, L);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

int launch_conv_post(const float* x, const float* w, const float* bias, float pre_slope, float pre_div,
                     float* wav, short* wav_i16, float max_wav, int B, int L, int C, int K, cudaStream_t s) {
    if (B == 0 || L == 0) return CMTTS_OK;
    CMTTS_REQUIRE(C % 4 == 0 && K * C * 4 <= 48 * 1024, "conv_post: shape");
    dim3 grid((L + 255) / 256, B);
    conv_post_kernel<<<grid, 256, K * C * sizeof(float), s>>>(x, w, bias, pre_slope, pre_div, wav, wav_i16, max_wav, L, C, K);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}
