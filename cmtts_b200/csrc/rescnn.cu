// DeepSpeaker ResCNN speaker encoder (zero-shot path): (B, T, 64, 1) normalised filter-bank frames ->
// 512-d L2-normalised embedding.  Reference: deepspeaker/conv_models.py:44-138 (TF-Keras; `predict` of
// deepspeaker/embedding.py:13-27).  Inference only: BatchNormalization folded into a per-channel scale / shift on the host.
//
// Not on the per-step hot path (one call per reference utterance, ~5.3 GFLOP per 160-frame window), so this is a plain
// fp32 FFMA implementation whose job is reference-faithful arithmetic (fp32 products, fp32 accumulation like TF's CPU /
// GPU kernels), not tensor-core throughput:
//   * direct NHWC convolution, TensorFlow "SAME" padding (the extra row / column goes AFTER: stride-2 5x5 convs pad 1 | 2);
//   * a thread owns one output channel and PIX consecutive output positions of a row, a warp 32 consecutive channels: the
//     HWIO weight read is one coalesced 128-byte line per (kh, kw, ci), the input patch of the block sits in shared
//     memory and is read as broadcast float4s (4 input channels per LDS.128);
//   * epilogue: BN scale / shift, clipped ReLU min(max(v, 0), 20), optionally + residual and the clip again — the order
//     of identity_block (conv_models.py:83-108: the ReLU comes BEFORE the add);
//   * tail kernel: Reshape((-1, 2048)) + mean over time + Dense(512) + l2_normalize (conv_models.py:52-66).
#include "common.cuh"

namespace {

constexpr int CI_CHUNK = 32;    // input channels staged per pass

// grid: (ceil(Wout / PIX) * Hout, ceil(Cout / blockDim.x), B); block: 64 or 128 threads (<= Cout)
// PIX = output positions per thread along W (8; 4 for the last stage, whose rows are 4 wide)
template <int K, int PIX>
__global__ void __launch_bounds__(128)
rescnn_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ res, float* __restrict__ out,
                   int H, int W, int Cin, int Hout, int Wout, int Cout, int stride, int pad_h, int pad_w) {
    // input patch of this block: K rows x ((PIX - 1) * stride + K) columns x CI_CHUNK channels
    extern __shared__ __align__(16) float patch[];
    const int wtiles = (Wout + PIX - 1) / PIX;
    const int oh = blockIdx.x / wtiles, ow0 = (blockIdx.x % wtiles) * PIX;
    const int co = blockIdx.y * blockDim.x + threadIdx.x;
    const int b = blockIdx.z;
    const int pw = (PIX - 1) * stride + K;                 // patch width in input columns
    const int ih0 = oh * stride - pad_h, iw0 = ow0 * stride - pad_w;
    const float* xb = x + (long long)b * H * W * Cin;

    float acc[PIX];
#pragma unroll
    for (int p = 0; p < PIX; ++p) acc[p] = 0.f;

    for (int c0 = 0; c0 < Cin; c0 += CI_CHUNK) {
        const int cn = min(CI_CHUNK, Cin - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < K * pw * CI_CHUNK; i += blockDim.x) {
            const int ci = i % CI_CHUNK, col = (i / CI_CHUNK) % pw, row = i / (CI_CHUNK * pw);
            const int ih = ih0 + row, iw = iw0 + col;
            float v = 0.f;
            if (ci < cn && ih >= 0 && ih < H && iw >= 0 && iw < W) v = xb[((long long)ih * W + iw) * Cin + c0 + ci];
            patch[i] = v;
        }
        __syncthreads();
        if (co < Cout) {
            for (int kh = 0; kh < K; ++kh) {
                for (int kw = 0; kw < K; ++kw) {
                    const float* wp = w + ((long long)(kh * K + kw) * Cin + c0) * Cout + co;
                    const float* pp = patch + (kh * pw + kw) * CI_CHUNK;
                    if (cn == CI_CHUNK) {
#pragma unroll 2
                        for (int ci = 0; ci < CI_CHUNK; ci += 4) {
                            const float w0 = wp[(long long)ci * Cout], w1 = wp[(long long)(ci + 1) * Cout];
                            const float w2 = wp[(long long)(ci + 2) * Cout], w3 = wp[(long long)(ci + 3) * Cout];
#pragma unroll
                            for (int p = 0; p < PIX; ++p) {
                                const float4 xv = *reinterpret_cast<const float4*>(pp + p * stride * CI_CHUNK + ci);
                                acc[p] = fmaf(xv.x, w0, acc[p]);
                                acc[p] = fmaf(xv.y, w1, acc[p]);
                                acc[p] = fmaf(xv.z, w2, acc[p]);
                                acc[p] = fmaf(xv.w, w3, acc[p]);
                            }
                        }
                    } else {
                        for (int ci = 0; ci < cn; ++ci) {
                            const float wv = wp[(long long)ci * Cout];
#pragma unroll
                            for (int p = 0; p < PIX; ++p) acc[p] = fmaf(pp[p * stride * CI_CHUNK + ci], wv, acc[p]);
                        }
                    }
                }
            }
        }
    }
    if (co >= Cout) return;
    const float sc = scale[co], sh = shift[co];
#pragma unroll
    for (int p = 0; p < PIX; ++p) {
        const int ow = ow0 + p;
        if (ow >= Wout) break;
        const long long o = (((long long)b * Hout + oh) * Wout + ow) * Cout + co;
        float v = fminf(fmaxf(fmaf(acc[p], sc, sh), 0.f), 20.f);
        if (res) v = fminf(fmaxf(v + res[o], 0.f), 20.f);
        out[o] = v;
    }
}

// Reshape((-1, W*C)) + mean over time + Dense + l2_normalize; one block per utterance, blockDim.x == n_out
__global__ void rescnn_tail_kernel(const float* __restrict__ x, const float* __restrict__ wd, const float* __restrict__ bd,
                                   float* __restrict__ emb, int H, int feat, int n_out) {
    extern __shared__ float m[];                            // [feat] means, then [32] partial sums
    float* red = m + feat;
    const float* xb = x + (long long)blockIdx.x * H * feat;
    for (int f = threadIdx.x; f < feat; f += blockDim.x) {
        float s = 0.f;
        for (int h = 0; h < H; ++h) s += xb[(long long)h * feat + f];
        m[f] = s / (float)H;
    }
    __syncthreads();
    const int n = threadIdx.x;
    float v = 0.f;
    if (n < n_out) {
        for (int k = 0; k < feat; ++k) v = fmaf(m[k], wd[(long long)k * n_out + n], v);
        v += bd[n];
    }
    float sq = (n < n_out) ? v * v : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = (threadIdx.x < (blockDim.x + 31) / 32) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    // K.l2_normalize: x * rsqrt(max(sum(x^2), 1e-12))
    if (n < n_out) emb[(long long)blockIdx.x * n_out + n] = v / sqrtf(fmaxf(red[0], 1e-12f));
}

inline int same_out(int in, int stride) { return (in + stride - 1) / stride; }
inline int same_pad_before(int in, int k, int stride) {
    const int out = same_out(in, stride);
    int total = (out - 1) * stride + k - in;
    if (total < 0) total = 0;
    return total / 2;                                       // TensorFlow puts the odd element after
}

template <int K, int PIX>
void launch_conv_t(const float* x, const float* w, const float* sc, const float* sh, const float* res, float* out, int B, int H,
                   int W, int Cin, int Cout, int stride, cudaStream_t s) {
    const int Hout = same_out(H, stride), Wout = same_out(W, stride);
    const int wtiles = (Wout + PIX - 1) / PIX;
    // 128 threads (4 warps of 32 output channels) unless that leaves most SMs without a block (the late, small stages)
    int threads = Cout < 128 ? ((Cout + 31) / 32) * 32 : 128;
    if (threads == 128 && (long long)wtiles * Hout * (Cout / 128) * B < 148) threads = 64;
    dim3 grid(wtiles * Hout, (Cout + threads - 1) / threads, B);
    const size_t smem = (size_t)K * ((PIX - 1) * stride + K) * CI_CHUNK * sizeof(float);
    rescnn_conv_kernel<K, PIX><<<grid, threads, smem, s>>>(x, w, sc, sh, res, out, H, W, Cin, Hout, Wout, Cout, stride,
                                                            same_pad_before(H, K, stride), same_pad_before(W, K, stride));
}

int launch_conv(const float* x, const float* w, const float* sc, const float* sh, const float* res, float* out, int B,
                int H, int W, int Cin, int Cout, int K, int stride, cudaStream_t s) {
    const int Hout = same_out(H, stride), Wout = same_out(W, stride);
    if (g_cmtts_prof_on) {
        char lbl[64];
        snprintf(lbl, sizeof(lbl), "rescnn_conv k%d s%d %d->%d", K, stride, Cin, Cout);
        cmtts_prof_note(lbl, 2.0 * B * Hout * Wout * (double)Cout * K * K * Cin,
                        4.0 * ((double)B * H * W * Cin + (double)B * Hout * Wout * Cout * (res ? 2 : 1) + (double)K * K * Cin * Cout));
    }
    const bool narrow = Wout <= 4;
    if (K == 5 && !narrow) launch_conv_t<5, 8>(x, w, sc, sh, res, out, B, H, W, Cin, Cout, stride, s);
    else if (K == 5) launch_conv_t<5, 4>(x, w, sc, sh, res, out, B, H, W, Cin, Cout, stride, s);
    else if (K == 3 && !narrow) launch_conv_t<3, 8>(x, w, sc, sh, res, out, B, H, W, Cin, Cout, stride, s);
    else if (K == 3) launch_conv_t<3, 4>(x, w, sc, sh, res, out, B, H, W, Cin, Cout, stride, s);
    else { cmtts_set_error("rescnn: kernel size must be 3 or 5", __FILE__, __LINE__); return CMTTS_ERR_ARG; }
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}

constexpr int N_STAGES = 4;
constexpr int BLOCKS_PER_STAGE = 3;
constexpr int CONVS = N_STAGES * (1 + 2 * BLOCKS_PER_STAGE);     // 28

inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

}  // namespace

// cfg = {n_fbanks, first_filters (64; doubles per stage), dense_out (512)}
extern "C" size_t cmtts_rescnn_workspace_bytes(const int32_t* cfg, int64_t B, int64_t T) {
    const int F = cfg[0], C0 = cfg[1];
    const size_t act = (size_t)B * same_out((int)T, 2) * same_out(F, 2) * C0;     // the largest activation (stage 1)
    return 3 * align256(act * sizeof(float));
}

// w: per conv {kernel HWIO fp32, scale [Cout], shift [Cout]} x 28 in network order (stage: strided 5x5 conv, then per identity
// block its 2a and 2b convs), then {dense kernel [feat][dense_out], dense bias}.  x: (B, T, n_fbanks) fp32 (one input channel).
extern "C" int cmtts_rescnn_forward(const int32_t* cfg, const void* const* w, const float* x, int64_t B_, int64_t T_,
                                    float* emb, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int B = (int)B_, T = (int)T_, F = cfg[0], C0 = cfg[1], NOUT = cfg[2];
    CMTTS_REQUIRE(cfg && w && x && emb, "rescnn: null argument");
    CMTTS_REQUIRE(T >= 1 && F >= 1 && C0 % 32 == 0 && NOUT >= 32 && NOUT <= 1024 && NOUT % 32 == 0, "rescnn: bad configuration");
    CMTTS_REQUIRE(ws_bytes >= cmtts_rescnn_workspace_bytes(cfg, B_, T_), "rescnn: workspace too small");
    if (B == 0) return CMTTS_OK;
    const size_t act = align256((size_t)B * same_out(T, 2) * same_out(F, 2) * C0 * sizeof(float));
    float* buf[3] = {(float*)ws, (float*)((char*)ws + act), (float*)((char*)ws + 2 * act)};
    const float* cur = x;
    int H = T, W = F, Cin = 1, wi = 0, ci = -1;                  // ci: index of the buffer holding `cur` (-1 = the input)
    for (int st = 0; st < N_STAGES; ++st) {
        const int Cout = C0 << st;
        // conv{Cout}-s: 5x5 stride 2 + BN + clipped ReLU                                     conv_models.py:112-124
        const int o = (ci + 1) % 3;                              // (-1 + 1) % 3 == 0 for the input
        CMTTS_TRY(launch_conv(cur, (const float*)w[wi], (const float*)w[wi + 1], (const float*)w[wi + 2], nullptr, buf[o], B, H, W,
                              Cin, Cout, 5, 2, s));
        wi += 3;
        H = same_out(H, 2); W = same_out(W, 2); Cin = Cout; cur = buf[o]; ci = o;
        for (int blk = 0; blk < BLOCKS_PER_STAGE; ++blk) {
            // identity_block: conv 3x3 + BN + clip ; conv 3x3 + BN + clip ; + input ; clip          conv_models.py:83-108
            const int t = (ci + 1) % 3, o2 = (ci + 2) % 3;
            CMTTS_TRY(launch_conv(cur, (const float*)w[wi], (const float*)w[wi + 1], (const float*)w[wi + 2], nullptr, buf[t], B, H, W,
                                  Cin, Cout, 3, 1, s));
            CMTTS_TRY(launch_conv(buf[t], (const float*)w[wi + 3], (const float*)w[wi + 4], (const float*)w[wi + 5], cur, buf[o2], B,
                                  H, W, Cin, Cout, 3, 1, s));
            wi += 6;
            cur = buf[o2]; ci = o2;
        }
    }
    const int feat = W * Cin;
    CMTTS_REQUIRE(feat <= 8192, "rescnn: feature width too large for the tail kernel");
    if (g_cmtts_prof_on) cmtts_prof_note("rescnn_tail mean + dense + l2norm", 2.0 * B * feat * NOUT, 4.0 * ((double)B * H * feat + (double)feat * NOUT));
    rescnn_tail_kernel<<<B, NOUT, (feat + 32) * sizeof(float), s>>>(cur, (const float*)w[3 * CONVS], (const float*)w[3 * CONVS + 1], emb,
                                                                     H, feat, NOUT);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}
