// tcgen05 implicit-GEMM conv1d (declarations).  See umma_conv.cu.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

enum UmmaEpi {
    UEPI_VOC = 0,      // fp16 "activated storage" epilogue (HiFi-GAN convs)
    UEPI_DN_COND = 1,  // y = acc + bias + addvec[b] + x_f32            -> fp16 hi/lo
    UEPI_DN_GATE = 2,  // sigmoid(gate + b) * tanh(filter + b)          -> fp16 hi/lo   (BN = 128: 64 gates | 64 filters)
    UEPI_DN_OUT = 3,   // cols <  N/2: x = (acc + b + addvec[b] + x) * out_scale (fp32, in place)
                       // cols >= N/2: skip (+)= acc + b                 (fp32)
    UEPI_DN_OUTY = 5,  // y-recurrence of the residual stack (see pipeline.cu): y = acc + bias + addvec[utterance]
                       // + x_f32[row] (the layer's precomputed conditioner projection, fp32, read ahead of the
                       // accumulator wait) -> fp16 hi/lo, written in place over the y operand (out_h/out_lo)
    UEPI_F32_PLANES = 6,  // v = acc*alpha + bias as fp32 into a stack of column planes (out_f32, out32_ncols, out32_plane),
                       // staged in shared memory and TMA-stored; nothing else (the per-batch conditioner GEMM)
    UEPI_F32 = 4,      // generic: v = act((acc*alpha + bias) * beta) + addvec[b] + res*res_scale ; v *= out_scale ;
                       // rows >= lens[b] -> 0 ; written as fp32 (out_f32) and/or fp16 hi/lo ; cols >= n_valid dropped
};

struct UmmaConvParams {
    // problem: out[b, t, n] = epi( sum_{tap, ci} A[b, t + shift[tap], ci] * W[tap][n][ci] )
    int B, M, Lin, N, Cin, taps;
    int shift[CMTTS_MAX_TAPS];
    int split;            // 0: plain fp16 operands; 1: hi/lo operand pairs, 3 MMAs per K step (fp32-class)
    int epi;
    // operands (fp16, channels-last activations [B][Lin][Cin]; weights [taps*N][Cin])
    const __half* a_hi; const __half* a_lo; long long a_bstride; int a_ld;
    const __half* w_hi; const __half* w_lo;
    // optional SECOND activation operand (split mode, taps == 1): the contraction runs over [A | A2], i.e.
    // K = Cin + Cin2 with weights [N][Cin + Cin2]; output columns >= n_k2 contract over the first Cin only
    const __half* a2_hi; const __half* a2_lo; long long a2_bstride; int a2_ld; int Cin2; int n_k2;
    // the first `a2_diag` channels of A2 meet BLOCK-DIAGONAL weights (an identity-like term): a tile whose output
    // columns are [n0, n0 + BN) only loads the A2 channel blocks [n0, n0 + BN) of that range (requires a2_diag == N)
    int a2_diag;
    // taps index the THIRD coordinate of the A tensor map (a stack of `taps` tensors, e.g. one per layer) instead of
    // shifting rows: K = taps * Cin over [tap][row][channel]
    int a_tap_dim;
    // flattened-utterance layout: B == 1, M == utterances * rows_per_utt rows, the last row of every utterance is a
    // zero guard row (the conv's padding between neighbours); the epilogue never writes guard rows and indexes
    // addvec by row / rows_per_utt.  0 = off (3-D tensor maps, one utterance per batch index).
    int rows_per_utt;
    // UEPI_F32 with rows_per_utt: the fp32 residual / output tensors (x_f32, out_f32) are ordinary (B, L, *) tensors
    // WITHOUT guard rows: their row index is (flattened row - utterance index)
    int io_unguard;
    // epilogue
    const float* bias; float alpha;
    // UEPI_VOC: v = acc*alpha + bias + inv_lrelu(res) + sum ; out = lrelu(v, out_slope) as fp16
    const __half* res_h; long long res_bstride; int res_ld; float res_inv_slope;
    const __half* sum_h;
    __half* out_h; __half* out_lo; long long out_bstride; int out_ld; float out_slope;
    // denoiser epilogues
    const float* addvec; long long addvec_bstride;
    float* x_f32; long long x_bstride; int x_ld;
    float* skip_f32; int skip_accumulate; float out_scale;
    // UEPI_F32 extras (res = x_f32 with x_bstride / x_ld)
    int act; float beta; float res_scale; const long long* lens; float* out_f32; long long out32_bstride; int out32_ld; int n_valid;
    // UEPI_F32: out_f32 as a stack of column planes — column n goes to plane n / out32_ncols (plane stride out32_plane
    // elements) at column n % out32_ncols (out32_ncols % 16 == 0; 0 = one plane).  One GEMM over stacked weights
    // [layers * C][K] then writes [layer][row][C].
    int out32_ncols; long long out32_plane;
    // transposed convs packed as 3-tap convs (weights.py: pack_conv_transpose): output columns < tap_split_n only have
    // non-zero weights in taps [0, taps-1), columns >= tap_split_n only in taps [1, taps) -> the all-zero tap of a tile
    // is skipped (identical results: it would add exact zeros).  0 = off.
    int tap_split_n;
    // fp8 (e4m3) operand copies for the CROSS TERMS of the denoiser's gate conv (umma_gate.cu, umma_gate8_kernel):
    // activations [M][Cin] bytes and weights [taps*N][Cin] bytes, each as (e4m3(hi), e4m3(lo * 2^11)).  NULL = fp16 cross terms.
    const unsigned char* a8_hi; const unsigned char* a8_lo; const unsigned char* w8_hi; const unsigned char* w8_lo;
    // UEPI_DN_COND / UEPI_DN_OUTY / UEPI_F32 (with out_h + out_lo): also write that e4m3 pair of the output rows (row pitch out8_ld bytes, flattened rows)
    unsigned char* out8_hi; unsigned char* out8_lo; int out8_ld;
    int dbg;              // experiment bits (CMTTS_UMMA_DBG): timing ablations 8 = no epilogue pre-loads, 16 = no copy-out / stores, 32 = no MMAs, 64 = epilogue does nothing (results wrong; tools/ablate.py); 1 = descriptor base_offset, 2 = disable the halo kernel, 128 = disable the gate kernel, 256 = fp16 (not fp8) cross terms in the gate kernel, 512 = no CTA-pair halo kernel, 1024 = no paired-row ResBlock kernel at C = 32
};

static inline UmmaConvParams umma_params_default() {
    UmmaConvParams p{};
    p.alpha = 1.f; p.res_inv_slope = 1.f; p.out_slope = 1.f; p.out_scale = 1.f; p.beta = 1.f; p.res_scale = 1.f;
    return p;
}

int launch_umma_conv(const UmmaConvParams& p, cudaStream_t s);
// halo-tile / resident-weight variant for Cin == N in {32, 64, 128}; CMTTS_ERR_UNSUPPORTED if not applicable
int launch_umma_halo(const UmmaConvParams& p, cudaStream_t s);
// CTA-pair (tcgen05 cta_group::2) variant of the halo kernel for C = 128 with streamed weights (k = 7, 11); see umma_halo2.cu
int launch_umma_halo2(const UmmaConvParams& p, cudaStream_t s);
// halo-A variant of the split (hi/lo) kernel for the denoiser's gate conv (UEPI_DN_GATE, 3 taps, Cin == 256);
// CMTTS_ERR_UNSUPPORTED if not applicable
int launch_umma_gate(const UmmaConvParams& p, cudaStream_t s);

// Fused ResBlock iteration y' = c2(lrelu(c1(a) + b1)) + b2 + inv_lrelu(a) on fp16 activated storage
// (all tensors [B][L][C] contiguous; weights [k][C][C] fp16); see umma_resblock.cu
struct UmmaResblockParams {
    int B, L, C, taps, dil;
    const __half* a;                       // lrelu(y)
    const __half* w1; const float* b1; float t_slope;
    const __half* w2; const float* b2; float alpha2;
    // C == 32 only, optional: the same weights pair-packed for the two-time-steps-per-row kernel (weights.py:
    // pair_pack_d1 / pair_pack_taps): w1p = pair_pack_d1(w1) if dil == 1 else pair_pack_taps(w1); w2p = pair_pack_d1(w2)
    const __half* w1p; const __half* w2p;
    float res_inv_slope;
    const __half* sum_h;                   // optional raw partial sum added before the output activation
    __half* out_h; float out_slope;
};
// CMTTS_ERR_UNSUPPORTED when the shape is not covered (C not in {32, 64}, weights do not fit, ...)
int launch_umma_resblock(const UmmaResblockParams& p, cudaStream_t s);

// fp32 -> fp16 (optionally hi/lo pair, optional leaky-ReLU, optional channel zero-padding)
int launch_f32_to_f16(const float* x, __half* hi, __half* lo, long long rows, int C, int Cpad, float slope, cudaStream_t s);
// flattened-utterance layout (one guard row per utterance): (B, L, C) fp32 -> rows b*Lp + t of an fp16 hi/lo matrix;
// and the same row mapping for an existing fp16 (B, L, W) tensor, placed at a column offset
int launch_f32_to_f16_rows(const float* x, __half* hi, __half* lo, int B, int L, int Lp, int C, int Cpad, int out_ld,
                           float scale, cudaStream_t s);
int launch_pack_rows_f16(const __half* src, __half* dst, int B, int L, int Lp, int W, int out_ld, int col_off, cudaStream_t s);
// HiFi-GAN output stage on fp16 activated input
int launch_conv_post_f16(const __half* x, const float* w, const float* bias, float pre_div, float* wav, short* wav_i16,
                         float max_wav, int B, int L, int C, int K, cudaStream_t s);
