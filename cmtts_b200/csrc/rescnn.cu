// DeepSpeaker ResCNN speaker encoder (zero-shot path): (B, T, 64, 1) normalised filter-bank frames ->
// 512-d L2-normalised embedding.  Reference: deepspeaker/conv_models.py:44-138 (TF-Keras; `predict` of
// deepspeaker/embedding.py:13-27).  Inference only: BatchNormalization folded into a per-channel scale / shift on the host.
//
// Not on the per-step hot path (one call per reference utterance, ~5.3 GFLOP per 160-frame window), so this is a plain
// fp32 FFMA implementation whose job is reference-faithful arithmetic (fp32 products, fp32 accumulation like TF's CPU /
// GPU kernels), not tensor-core throughput:
//   * direct NHWC convolution, TensorFlow "SAME" padding (the extra row / column goes AFTER: stride-2 5x5 convs pad 1 | 2);
//   * a thread owns one output channel and PIX consecutive output positions of a row, a warp 32 consecutive channels: the
//     HWIO weight read is one coalesced 128-byte line per (kh, kw, ci), the warp's input patch sits in shared memory and
//     is read as broadcast float4s (4 input channels per LDS.128); the reduction over input channels is split over the
//     warps of a block (the late stages have 40 - 160 output positions: without the split most SMs had nothing to do);
//   * epilogue: BN scale / shift, clipped ReLU min(max(v, 0), 20), optionally + residual and the clip again — the order
//     of identity_block (conv_models.py:83-108: the ReLU comes BEFORE the add);
//   * tail: Reshape((-1, 2048)) + mean over time + Dense(512), then l2_normalize (conv_models.py:52-66).
#include "common.cuh"

namespace {

constexpr int CI_CHUNK = 32;    // input channels per pass of a warp
constexpr int MAX_WARPS = 4;

// One block = PIX consecutive output positions of one output row x 32 output channels.  The reduction over K x K x Cin is
// SPLIT OVER THE BLOCK'S WARPS by 32-channel chunk (warp w takes chunks w, w + nwarps, ...); partial sums meet in shared
// memory.  Per (chunk, tap) a thread first issues all 32 weight loads (one coalesced 128-byte line per input channel across
// the warp) and only then runs the 32 x PIX FMAs: the first version loaded 4 weights, used them, loaded the next 4 — one L2
// round trip per 16 FMAs, 873 us for a 512 -> 512 layer on a 10 x 4 map (ncu launch list, profiles/launches_r4_rescnn_before_summary.txt).
// grid: (ceil(Wout / PIX) * Hout, Cout / 32, B); block: 32 * min(4, Cin / 32 rounded up) threads
template <int K, int PIX>
__global__ void __launch_bounds__(32 * MAX_WARPS)
rescnn_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ res, float* __restrict__ out,
                   int H, int W, int Cin, int Hout, int Wout, int Cout, int stride, int pad_h, int pad_w) {
    extern __shared__ __align__(16) float smem[];          // [nwarps][K][pw][CI_CHUNK] input patches, reused for the reduction
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int wtiles = (Wout + PIX - 1) / PIX;
    const int oh = blockIdx.x / wtiles, ow0 = (blockIdx.x % wtiles) * PIX;
    const int co = blockIdx.y * 32 + lane;
    const int b = blockIdx.z;
    const int pw = (PIX - 1) * stride + K;                 // patch width in input columns
    const int ih0 = oh * stride - pad_h, iw0 = ow0 * stride - pad_w;
    const float* xb = x + (long long)b * H * W * Cin;
    float* patch = smem + warp * (K * pw * CI_CHUNK);
    const bool co_ok = co < Cout;

    float acc[PIX];
#pragma unroll
    for (int p = 0; p < PIX; ++p) acc[p] = 0.f;

    const int nchunks = (Cin + CI_CHUNK - 1) / CI_CHUNK;
    for (int c = warp; c < nchunks; c += nwarps) {
        const int c0 = c * CI_CHUNK, cn = min(CI_CHUNK, Cin - c0);
        __syncwarp();
        for (int i = lane; i < K * pw * CI_CHUNK; i += 32) {            // lane = input channel: one 128-byte line per position
            const int col = (i / CI_CHUNK) % pw, row = i / (CI_CHUNK * pw);
            const int ih = ih0 + row, iw = iw0 + col;
            float v = 0.f;
            if (lane < cn && ih >= 0 && ih < H && iw >= 0 && iw < W) v = xb[((long long)ih * W + iw) * Cin + c0 + lane];
            patch[i] = v;
        }
        __syncwarp();
        for (int kh = 0; kh < K; ++kh) {
            for (int kw = 0; kw < K; ++kw) {
                const float* wp = w + ((long long)(kh * K + kw) * Cin + c0) * Cout + co;
                float wv[CI_CHUNK];
#pragma unroll
                for (int ci = 0; ci < CI_CHUNK; ++ci) wv[ci] = (co_ok && ci < cn) ? wp[(long long)ci * Cout] : 0.f;
                const float* pp = patch + (kh * pw + kw) * CI_CHUNK;
#pragma unroll
                for (int ci = 0; ci < CI_CHUNK; ci += 4) {
#pragma unroll
                    for (int p = 0; p < PIX; ++p) {
                        const float4 xv = *reinterpret_cast<const float4*>(pp + p * stride * CI_CHUNK + ci);
                        acc[p] = fmaf(xv.x, wv[ci], acc[p]);
                        acc[p] = fmaf(xv.y, wv[ci + 1], acc[p]);
                        acc[p] = fmaf(xv.z, wv[ci + 2], acc[p]);
                        acc[p] = fmaf(xv.w, wv[ci + 3], acc[p]);
                    }
                }
            }
        }
    }
    // partial sums of the warps -> warp 0 (fixed order: the result does not depend on scheduling)
    __syncthreads();
    if (warp > 0) {
#pragma unroll
        for (int p = 0; p < PIX; ++p) smem[((warp - 1) * PIX + p) * 32 + lane] = acc[p];
    }
    __syncthreads();
    if (warp > 0 || !co_ok) return;
    for (int wq = 1; wq < nwarps; ++wq) {
#pragma unroll
        for (int p = 0; p < PIX; ++p) acc[p] += smem[((wq - 1) * PIX + p) * 32 + lane];
    }
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

// Reshape((-1, W*C)) + mean over time + Dense: grid (B, n_out / 32), 8 warps per block; warp w reduces the k range
// [w, w + 1) * feat / 8 for the block's 32 outputs (coalesced 128-byte weight lines), partial sums meet in shared memory.
// (One block per utterance spent 223 us streaming the 4 MB dense kernel through a single SM.)
__global__ void __launch_bounds__(256)
rescnn_dense_kernel(const float* __restrict__ x, const float* __restrict__ wd, const float* __restrict__ bd,
                    float* __restrict__ emb, int H, int feat, int n_out) {
    extern __shared__ float m[];                            // [feat] means, then [8][32] partial sums
    float* red = m + feat;
    const float* xb = x + (long long)blockIdx.x * H * feat;
    for (int f = threadIdx.x; f < feat; f += blockDim.x) {
        float s = 0.f;
        for (int h = 0; h < H; ++h) s += xb[(long long)h * feat + f];
        m[f] = s / (float)H;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y * 32 + lane;
    const int k0 = (int)((long long)feat * warp / 8), k1 = (int)((long long)feat * (warp + 1) / 8);
    float v = 0.f;
    if (n < n_out) {
#pragma unroll 16
        for (int k = k0; k < k1; ++k) v = fmaf(m[k], wd[(long long)k * n_out + n], v);
    }
    red[warp * 32 + lane] = v;
    __syncthreads();
    if (warp == 0 && n < n_out) {
        float t = bd[n];
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q * 32 + lane];
        emb[(long long)blockIdx.x * n_out + n] = t;
    }
}

// K.l2_normalize(axis=1): x / sqrt(max(sum(x^2), 1e-12)), in place; one block per utterance, blockDim.x == n_out (<= 1024)
__global__ void rescnn_l2norm_kernel(float* __restrict__ emb, int n_out) {
    __shared__ float red[32];
    const int n = threadIdx.x;
    const float v = emb[(long long)blockIdx.x * n_out + n];
    float sq = v * v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((n & 31) == 0) red[n >> 5] = sq;
    __syncthreads();
    if (n < 32) {
        float t = (n < (int)(blockDim.x >> 5)) ? red[n] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (n == 0) red[0] = t;
    }
    __syncthreads();
    emb[(long long)blockIdx.x * n_out + n] = v / sqrtf(fmaxf(red[0], 1e-12f));
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
    const int nchunks = (Cin + CI_CHUNK - 1) / CI_CHUNK;
    const int nwarps = nchunks < MAX_WARPS ? nchunks : MAX_WARPS;
    dim3 grid(wtiles * Hout, (Cout + 31) / 32, B);
    const size_t smem = (size_t)nwarps * K * ((PIX - 1) * stride + K) * CI_CHUNK * sizeof(float);
    rescnn_conv_kernel<K, PIX><<<grid, 32 * nwarps, smem, s>>>(x, w, sc, sh, res, out, H, W, Cin, Hout, Wout, Cout, stride,
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
    if (g_cmtts_prof_on) cmtts_prof_note("rescnn_dense mean + dense", 2.0 * B * feat * NOUT, 4.0 * ((double)B * H * feat + (double)feat * NOUT));
    rescnn_dense_kernel<<<dim3(B, NOUT / 32), 256, (feat + 256) * sizeof(float), s>>>(cur, (const float*)w[3 * CONVS],
                                                                                     (const float*)w[3 * CONVS + 1], emb, H, feat, NOUT);
    CMTTS_CHECK_LAUNCH();
    rescnn_l2norm_kernel<<<B, NOUT, 0, s>>>(emb, NOUT);
    CMTTS_CHECK_LAUNCH();
    return CMTTS_OK;
}
