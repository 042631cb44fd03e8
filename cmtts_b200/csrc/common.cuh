// Shared declarations for the cmtts_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define CMTTS_OK 0
#define CMTTS_ERR_ARG (-1)
#define CMTTS_ERR_CUDA (-2)
#define CMTTS_ERR_WORKSPACE (-3)
#define CMTTS_ERR_UNSUPPORTED (-4)

extern unsigned long long g_cmtts_launches;   // kernels launched by this library (bench evidence)
extern int g_cmtts_umma_dbg;                  // CMTTS_UMMA_DBG experiment bits (-1: not read yet)
extern int g_cmtts_pdl;                       // CMTTS_PDL: programmatic dependent launch on (1, default) / off (0)

// launch profiler (cmtts_prof_begin / cmtts_prof_end of the C ABI): when on, every launch site records a CUDA event
// on the profiled stream; a launch's duration is the distance to the previous event.  cmtts_prof_note() labels the NEXT
// launch and attaches its algorithmic FLOPs / bytes (sites without a note are labelled function:line).
extern int g_cmtts_prof_on;
void cmtts_prof_note(const char* label, double flops, double bytes);
void cmtts_prof_mark(const char* func, int line);

#define CMTTS_CHECK_LAUNCH()                                  \
    do {                                                      \
        ++g_cmtts_launches;                                   \
        if (g_cmtts_prof_on) cmtts_prof_mark(__func__, __LINE__); \
        cudaError_t e__ = cudaPeekAtLastError();              \
        if (e__ != cudaSuccess) { cmtts_set_error(cudaGetErrorString(e__), __FILE__, __LINE__); return CMTTS_ERR_CUDA; } \
    } while (0)

#define CMTTS_REQUIRE(cond, msg)                              \
    do {                                                      \
        if (!(cond)) { cmtts_set_error(msg, __FILE__, __LINE__); return CMTTS_ERR_ARG; } \
    } while (0)

#define CMTTS_TRY(expr)                                       \
    do { int rc__ = (expr); if (rc__ != CMTTS_OK) return rc__; } while (0)

void cmtts_set_error(const char* msg, const char* file, int line);

// ---------------------------------------------------------------------------------------------
// conv1d over channels-last activations, expressed as an implicit GEMM
//   out[b, t, n] = epi( sum_{tap, ci} pre(x[b, t + shift[tap], ci]) * w[tap][ci][n] )
// rows outside [0, Lin) read as zero AFTER pre() is applied to real rows only (zero padding of
// the activated tensor, which is what F.conv1d(padding=...) on lrelu(x) gives).
// ---------------------------------------------------------------------------------------------
#define CMTTS_MAX_TAPS 16

enum ConvAct { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_LRELU = 3, ACT_TANH = 4, ACT_GATED = 5, ACT_SWISH = 6 };

struct ConvParams {
    // input
    const float* x; long long x_bstride; int x_ld; int Lin; int Cin;
    // weights [taps][Cin][N] (N contiguous)
    const float* w; int taps; int shift[CMTTS_MAX_TAPS];
    // prologue on A
    int pre_lrelu; float pre_slope;
    // output [B][M][Nout] ; Nout = N (or N/2 for ACT_GATED)
    float* out; long long out_bstride; int out_ld; int M; int N; int B;
    // epilogue: v = acc*alpha + bias ; aux_out = v ; v = act(v*beta) ; v += addvec[b] ;
    //           v += res1*res1_scale ; v *= out_scale ; rows >= lens[b] -> 0 ; out = (accumulate? out : 0) + v
    const float* bias; float alpha; float beta; int act; float act_slope;
    float* aux_out; long long aux_bstride; int aux_ld;
    const float* addvec; long long addvec_bstride;
    const float* res1; long long res1_bstride; int res1_ld; float res1_scale;
    float out_scale;
    const long long* lens;
    int accumulate;
    // optional fp16 copy of the result with leaky-ReLU applied (operand of a following tensor-core conv)
    __half* out_h; float out_h_slope;
    // opt-in for the few-rows kernel (B == 1, M <= 32, one tap): the 128-row tile kernel runs such a GEMM on 1-8 CTAs.
    // The summation order differs from the tile kernel's, so only call sites that do not feed a quantiser set it.
    int few_rows_ok;
};

static inline ConvParams conv_params_default() {
    ConvParams p{};
    p.alpha = 1.f; p.beta = 1.f; p.res1_scale = 1.f; p.out_scale = 1.f; p.act = ACT_NONE; p.out_h_slope = 1.f;
    return p;
}

int launch_conv1d_simt(const ConvParams& p, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// row-wise kernels (rowops.cu / attention.cu)
// ---------------------------------------------------------------------------------------------
int launch_layernorm(const float* x, const float* w, const float* b, float eps, float* out,
                     int B, int T, int C, const long long* lens, cudaStream_t s);
int launch_ln_head(const float* x, const float* lw, const float* lb, float eps, const float* hw,
                   const float* hb, int odim, float scale, float* out, int B, int T, int C,
                   const long long* lens, cudaStream_t s);
int launch_embed_tokens(const long long* tokens, const float* emb, const float* pe, int pe_rows,
                        float emb_scale, float* out, int B, int T, int C, const long long* lens,
                        cudaStream_t s);
int launch_add_rowvec(float* x, const float* vec, int B, int T, int C, cudaStream_t s);
int launch_add_positional(const float* x, const float* pe, int pe_rows, const float* alpha,
                          float* out, int B, int T, int C, cudaStream_t s);
int launch_energy_embed(const float* x, const float* pred, float control, const float* bins, int nbins,
                        const float* emb, float* out, long long* idx_out, float* pred_out,
                        int B, int T, int C, cudaStream_t s);
int launch_round_durations(const float* log_d, float d_control, const long long* src_lens,
                           float* d_rounded, long long* cumsum, long long* mel_lens,
                           int B, int T, cudaStream_t s);
int launch_length_regulate(const float* x, const long long* cumsum, const long long* mel_lens,
                           float* out, long long* mel2ph, int B, int T, int L, int C, cudaStream_t s);
int launch_attention(const float* qkv, const long long* src_lens, float* out, int B, int T, int C,
                     int heads, cudaStream_t s);
int launch_mish(float* x, long long n, cudaStream_t s);
int launch_dn_fuse_steps(const float* ds_all, const float* dsp_all, float* yc, int B, int layers, int C, float r,
                         cudaStream_t s);
int launch_renoise(const float* x0, const float* noise, float s1, float s2, float* out, long long n,
                   cudaStream_t s);
int launch_scale(const float* x, float a, float* out, long long n, cudaStream_t s);
int launch_transpose_bcl_to_blc(const float* x, float* out, int B, int C, int L, cudaStream_t s);
int launch_conv_post(const float* x, const float* w, const float* bias, float pre_slope, float pre_scale,
                     float* wav, short* wav_i16, float max_wav, int B, int L, int C, int K, cudaStream_t s);
