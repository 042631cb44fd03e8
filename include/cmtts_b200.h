/* cmtts_b200 — C ABI of the B200-native CM-TTS inference hot path.
 *
 * The reference (XiangLi2022/CM-TTS) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md §8b): its seams are the Python call protocols of
 *   - DurationPitchSpeakerNet.forward      model/cmtts.py:44-122        (encoder + variance adaptor)
 *   - Denoiser.forward / KarrasDenoiser.denoise   model/modules.py:600-638, karras_diffusion.py:392-407
 *   - stochastic_iterative_sampler re-noise       model/cm_tool/karras_diffusion.py:829-854
 *   - hifigan.Generator.forward / vocoder_infer   hifigan/models.py:149-165, utils/model.py:187-205
 * Each entry point below replaces the arithmetic behind one of those seams; the Python package
 * cmtts_b200 binds them with ctypes and re-exposes the reference's call protocols on top
 * (see INTEGRATION.md for the binding a maintainer would add to the reference).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named `host_*`;
 *   - activations are channels-last fp32: (B, T, C) with C contiguous; indices/lengths are int64;
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     immediately; nothing allocates, nothing synchronises, nothing throws;
 *   - return value 0 on success, negative on error (cmtts_last_error() gives the text);
 *   - `ws` is caller-owned scratch of at least the matching cmtts_*_workspace_bytes();
 *   - weight tables are arrays of device pointers in the order of the enums below, holding
 *     tensors re-packed by cmtts_b200/weights.py (conv weights as [tap][Cin][Cout]).
 */
#ifndef CMTTS_B200_H
#define CMTTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMTTS_ABI_VERSION 2

int cmtts_abi_version(void);
const char* cmtts_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t cmtts_launch_count(void);
/* experiment switches for A/B timing in one process (tools/ab_switch.py); production code never calls it.
 * umma_dbg: bit field of CMTTS_UMMA_DBG (2 = no halo kernel, 128 = no gate kernel, ...); pdl: 1/0 = programmatic
 * dependent launch on/off; -1 = take the value from the environment */
void cmtts_debug_set(int32_t umma_dbg, int32_t pdl);
/* launch profiler (bench.py's per-kernel table and roofline block; the reference has no counterpart — its closest
 * relative is the wall-clock Timer of p_rtf_cm.py:64-108).  Between begin and end every kernel launch of this library
 * on `stream` is followed by a CUDA event; cmtts_prof_end synchronises the stream and writes one line per kernel label,
 * "label\tlaunches\ttotal_us\talgorithmic_flops\talgorithmic_bytes\n", into host buffer `buf` (NUL-terminated, truncated
 * to `cap`) and returns the number of bytes needed, or a negative error code.  Timing a launch as the distance between
 * consecutive events serialises programmatic dependent launches, so a profiled step is a few % slower than a plain one:
 * the profile gives each kernel's SHARE of the step, the step time itself is measured without it. */
int cmtts_prof_begin(void* stream);
int64_t cmtts_prof_end(char* host_buf, size_t cap);

/* ---- model dimensions shared by the acoustic entry points ---- */
typedef struct cmtts_dims {
    int32_t hidden;        /* 256  encoder_hidden                                   */
    int32_t enc_layers;    /* 4                                                    */
    int32_t enc_heads;     /* 2                                                    */
    int32_t ffn_kernel;    /* 9                                                    */
    int32_t ffn_act;       /* 2 gelu, 1 relu, 6 swish (ConvAct codes)              */
    int32_t filter;        /* 256  variance_predictor.filter_size                  */
    int32_t dur_layers;    /* 2                                                    */
    int32_t dur_kernel;    /* 3                                                    */
    int32_t pred_layers;   /* 2                                                    */
    int32_t pred_kernel;   /* 5                                                    */
    int32_t cwt_hidden;    /* 128                                                  */
    int32_t cwt_out;       /* 11 (use_uv) or 10                                    */
    int32_t use_uv;        /* 1 / 0                                                */
    int32_t energy_bins;   /* 256  (boundaries = energy_bins - 1)                  */
    int32_t pitch_bins;    /* 300                                                  */
    int32_t n_mels;        /* 80                                                   */
    int32_t res_layers;    /* 20                                                   */
    int32_t res_channels;  /* 256                                                  */
    int32_t multi_speaker; /* 1 / 0                                                */
    int32_t spk_dim;       /* 512 external speaker embedding size                  */
    int32_t pe_rows;       /* rows of the sinusoid tables (>= max T, L) + 1         */
    float cwt_std_scale;   /* 0.8                                                  */
    float pitch_eps;       /* 1e-9                                                 */
    float f0_mel_min;      /* (float)(1127 ln(1 + 50/700))                         */
    float f0_mel_span;     /* (float)(f0_mel_max - f0_mel_min)                     */
} cmtts_dims;

/* ---- weight-table layouts (indices into `const void* const* w`) ---- */
enum { CMTTS_ENC_EMB = 0, CMTTS_ENC_PE = 1, CMTTS_ENC_LAYER0 = 2, CMTTS_ENC_PER_LAYER = 10 };
/* per layer: ln1_w, ln1_b, in_proj[C][3C], out_proj[C][C], ln2_w, ln2_b, ffn1[k][C][4C], ffn1_b,
 *            ffn2[4C][C], ffn2_b ; after the last layer: final_ln_w, final_ln_b */
enum {
    CMTTS_VA_SPK_W = 0, CMTTS_VA_SPK_B,                 /* [spk_dim][C], [C] (multi-speaker) */
    CMTTS_VA_DUR0,                                      /* dur_layers x {conv[k][Cin][F], b, ln_w, ln_b} then head_w[1][F], head_b */
    /* followed by: energy {alpha, layers x4, head_w, head_b}, PE(C) table, energy_bins,
     * energy_emb, stats {w0[C][h], b0, w2[h][h], b2, w4[h][4], b4[4]},
     * cwt_in {w[C][h], b}, PE(h) table, cwt {alpha, layers x4, head_w[cwt_out][F], head_b},
     * cwt_b[10], pitch_emb — exact indices are computed by cmtts_b200/weights.py and
     * mirrored in csrc/pipeline.cu (struct VaIdx). */
};
/* denoiser: in_w[1][M][C], in_b, freq[C/2], mlp0[C][4C], mlp2[4C][C], dproj_all[C][Lr*C],
 *           sproj_all[C][Lr*C] (or NULL), then per layer {cond_w[1][H][C], cond_b, k3_w[3][C][2C]
 *           (gate/filter interleaved per 64), k3_b, outx_w[1][C][C], outx_b, outs_w[1][C][C], outs_b},
 *           skip_w, skip_b, out_w[1][C][M], out_b */
enum { CMTTS_DN_IN_W = 0, CMTTS_DN_IN_B, CMTTS_DN_FREQ, CMTTS_DN_MLP0, CMTTS_DN_MLP2, CMTTS_DN_DPROJ,
       CMTTS_DN_SPROJ, CMTTS_DN_LAYER0, CMTTS_DN_PER_LAYER = 8 };

/* hifigan config: ints {n_levels, C0, n_kernels, n_dil, pre_k, post_k, rates[n_levels],
 *                       up_taps[n_levels], up_shift0[n_levels], ksize[n_kernels], dil[n_kernels*n_dil],
 *                       split_ok[n_levels]}   (split_ok[i] = 1: the packed transposed conv of level i has 3 taps, the
 *                       first feeding only the first half of its s*Cout columns, the last only the second half) */
/* hifigan weights: pre_w[k][80][C0], pre_b, then per level { up_w[taps][Cin][s*Cout], up_b[s*Cout],
 *                  per MRF resblock j of that level, per dilation m: {c1_w, c1_b, c2_w, c2_b} },
 *                  post_w[k][C], post_b */

/* ---- E1-E4: FastspeechEncoder.forward (model/modules.py:132-151, :80-105) ---- */
size_t cmtts_encoder_workspace_bytes(const cmtts_dims* d, int64_t B, int64_t T);
int cmtts_encoder_forward(const cmtts_dims* d, const void* const* w, const int64_t* tokens,
                          const int64_t* src_lens, int64_t B, int64_t T, float* enc_out,
                          void* ws, size_t ws_bytes, void* stream);

/* ---- V1-V3 + duration scan: token-rate half of VarianceAdaptor.forward (modules.py:331-376) ---- */
size_t cmtts_variance_token_workspace_bytes(const cmtts_dims* d, int64_t B, int64_t T);
int cmtts_variance_token(const cmtts_dims* d, const void* const* w, const float* enc,
                         const int64_t* src_lens, const float* spker_embeds, float e_control,
                         float d_control, int64_t B, int64_t T,
                         float* out1, float* log_d, float* d_rounded, float* e_pred, int64_t* e_idx,
                         int64_t* cumsum /* (B,2,T) */, int64_t* mel_lens, float* spk_emb /* (B,C) or NULL */,
                         float* f0_stats /* (B,4): mean, std, 0, 0 */, void* ws, size_t ws_bytes, void* stream);

/* ---- V4-V5: length regulator + frame-rate pitch path (modules.py:374-395, :273-307) ---- */
size_t cmtts_variance_frame_workspace_bytes(const cmtts_dims* d, int64_t B, int64_t L);
int cmtts_variance_frame(const cmtts_dims* d, const void* const* w, const float* out1,
                         const int64_t* cumsum, const int64_t* mel_lens, const float* f0_stats,
                         float p_control, int64_t B, int64_t T, int64_t L,
                         float* cond, int64_t* mel2ph, float* cwt, float* f0_denorm, int64_t* pitch_idx,
                         void* ws, size_t ws_bytes, void* stream);

/* ---- D2: step embedding + per-layer projections (blocks.py:633-640, :669-674; modules.py:579-583) ---- */
size_t cmtts_denoiser_prepare_workspace_bytes(const cmtts_dims* d, int64_t B);
int cmtts_denoiser_prepare(const cmtts_dims* d, const void* const* w, const float* t /* (B,) */,
                           const float* spk_emb /* (B,C) or NULL */, int64_t B,
                           float* ds_all /* (B, Lr*C) */, float* dsp_all /* (B, Lr*C) */,
                           void* ws, size_t ws_bytes, void* stream);

/* ---- D1/D3 + S4: one consistency-function evaluation
 *      out = c_out * F(c_in * x_t, t, cond) + c_skip * x_t   (karras_diffusion.py:392-407) ---- */
size_t cmtts_denoiser_workspace_bytes(const cmtts_dims* d, int64_t B, int64_t L);
int cmtts_denoiser_forward(const cmtts_dims* d, const void* const* w, const float* x_t /* (B,L,M) */,
                           const float* cond /* (B,L,H) */, const float* ds_all, const float* dsp_all,
                           float c_in, float c_out, float c_skip, int64_t B, int64_t L,
                           float* out /* (B,L,M) */, float* model_out /* (B,L,M) or NULL */,
                           void* ws, size_t ws_bytes, void* stream);

/* ---- S3: x = x0 + (noise * s1) * s2   (karras_diffusion.py:852) ---- */
int cmtts_renoise(const float* x0, const float* noise, float s1, float s2, float* out, int64_t n, void* stream);

/* ---- H1-H3: hifigan.Generator.forward + int16 conversion ---- */
size_t cmtts_hifigan_workspace_bytes(const int32_t* cfg, int64_t B, int64_t L);
int cmtts_hifigan_forward(const int32_t* cfg, const void* const* w, const float* mel /* (B,L,80) */,
                          int64_t B, int64_t L, float* wav /* (B, hop*L) or NULL */,
                          int16_t* wav_i16 /* (B, hop*L) or NULL */, float max_wav_value,
                          void* ws, size_t ws_bytes, void* stream);

/* ---- layout helper: (B, C, L) -> (B, L, C) ---- */
int cmtts_transpose_bcl_blc(const float* x, float* out, int64_t B, int64_t C, int64_t L, void* stream);

/* ---- single-op entry points (unit tests and stage-isolated parity checks) ---- */
typedef struct cmtts_conv_desc {
    int32_t B, M, Lin, Cin, N, taps;
    int32_t shift[16];
    int32_t x_ld, out_ld, res_ld;
    int64_t x_bstride, out_bstride, res_bstride, addvec_bstride;
    int32_t pre_lrelu; float pre_slope;
    float alpha, beta; int32_t act; float act_slope;
    float res_scale, out_scale;
    int32_t accumulate;
} cmtts_conv_desc;
int cmtts_conv1d(const cmtts_conv_desc* c, const float* x, const float* w, const float* bias,
                 const float* addvec, const float* res, const int64_t* lens, float* out, void* stream);
int cmtts_layernorm(const float* x, const float* w, const float* b, float eps, float* out,
                    int64_t B, int64_t T, int64_t C, const int64_t* lens, void* stream);
int cmtts_attention(const float* qkv, const int64_t* src_lens, float* out, int64_t B, int64_t T,
                    int64_t C, int64_t heads, void* stream);
int cmtts_length_regulate(const float* x, const int64_t* cumsum, const int64_t* mel_lens, float* out,
                          int64_t* mel2ph, int64_t B, int64_t T, int64_t L, int64_t C, void* stream);
int cmtts_round_durations(const float* log_d, float d_control, const int64_t* src_lens, float* d_rounded,
                          int64_t* cumsum, int64_t* mel_lens, int64_t B, int64_t T, void* stream);

/* ==== tensor-core (tcgen05) path ================================================================
 * fp16 operands in HBM, fp32 accumulation in TMEM.  The vocoder uses plain fp16 operands with
 * "activated storage" (every tensor is stored as leaky_relu(x), the raw value is recovered exactly
 * in the residual epilogue); the denoiser stack uses fp16 hi/lo operand pairs (3 MMAs per K step,
 * fp32-class products).  Weight tables: see cmtts_b200/weights.py (PackedAcoustic.dn16,
 * PackedHifiGan.table16). */
typedef struct cmtts_umma_desc {
    int32_t B, M, Lin, N, Cin, taps;
    int32_t shift[16];
    int32_t split, epi;                   /* epi: 0 VOC, 1 DN_COND, 2 DN_GATE, 3 DN_OUT, 4 F32 */
    int32_t a_ld, res_ld, out_ld, x_ld;
    int64_t a_bstride, res_bstride, out_bstride, x_bstride, addvec_bstride;
    float alpha, res_inv_slope, out_slope, out_scale;
    int32_t skip_accumulate;
} cmtts_umma_desc;
int cmtts_umma_conv1d(const cmtts_umma_desc* c, const void* a_hi, const void* a_lo, const void* w_hi,
                      const void* w_lo, const float* bias, const void* res_h, const void* sum_h,
                      void* out_h, void* out_lo, const float* addvec, float* x_f32, float* skip_f32,
                      void* stream);
/* fp32 (rows, C) -> fp16 (rows, Cpad) hi [+ lo], optional leaky-ReLU slope (1.0 = none) */
int cmtts_f32_to_f16(const float* x, void* hi, void* lo, int64_t rows, int64_t C, int64_t Cpad,
                     float slope, void* stream);

/* D1/D3 + S4 on tensor cores: same contract as cmtts_denoiser_forward; `w16` holds per layer
 * {cond_w hi, lo [C][H]; k3_w hi, lo [3*2C][C] (gate/filter interleaved per 64); out_w hi, lo [2C][C];
 *  out_b fp32 [2C]}, then {in_w hi, lo [C][128] (K zero-padded); skip_w hi, lo [C][C]}, then per layer
 * l < res_layers-1 the y-recurrence operands {y_w hi, lo [C][2C]; y_b fp32 [C]} with
 * y_w = [r Wo_l[:C] | r I], r = 1/sqrt(2), then the stacked skip projection
 * {skip_stack_w hi, lo [res_layers*C][C] (row l*C+n = Wo_l[C+n]); summed bias fp32 [C]}, then the output
 * projection {out_w hi, lo [128][C] (rows >= n_mels zero); out_b fp32 [128]}, then per layer the k3 conv's weights
 * once more as an e4m3 pair {hi8 = e4m3(hi), lo8 = e4m3(lo * 2^11), bytes [3*2C][C]}: the operands of the gate conv's
 * cross terms A_hi W_lo + A_lo W_hi, which run as kind::f8f6f4 MMAs (csrc/umma_gate.cu; CMTTS_GATE_FP8=0 keeps them
 * in fp16), then the conditioner stack {cond_stack_w hi, lo [res_layers*C][H]; bias fp32 [res_layers*C]}: row block 0
 * = Wc_0 (+ bc_0), row block l > 0 = Wc_l - r Wc_{l-1} (cmtts_b200/weights.py).
 *
 * The conditioner (B,L,H) is the same for every solver step (karras_diffusion.py:560-566 re-derives it per step from
 * the same inputs), so its projections for ALL layers are made once per batch by cmtts_denoiser_cond_tc:
 * cond_proj = fp32 [res_layers][B*(L+1)][C] (cmtts_denoiser_cond_tc_bytes), row b*(L+1)+t, one unused guard row per
 * utterance; cmtts_denoiser_forward_tc consumes it. */
size_t cmtts_denoiser_cond_tc_bytes(const cmtts_dims* d, int64_t B, int64_t L);
size_t cmtts_denoiser_cond_tc_workspace_bytes(const cmtts_dims* d, int64_t B, int64_t L);
int cmtts_denoiser_cond_tc(const cmtts_dims* d, const void* const* w, const void* const* w16,
                           const float* cond /* (B,L,H) */, int64_t B, int64_t L, float* cond_proj,
                           void* ws, size_t ws_bytes, void* stream);
size_t cmtts_denoiser_tc_workspace_bytes(const cmtts_dims* d, int64_t B, int64_t L);
int cmtts_denoiser_forward_tc(const cmtts_dims* d, const void* const* w, const void* const* w16,
                              const float* x_t, const float* cond_proj,
                              const float* ds_all, const float* dsp_all, float c_in, float c_out,
                              float c_skip, int64_t B, int64_t L, float* out, float* model_out,
                              void* ws, size_t ws_bytes, void* stream);

/* E1-E4 / V1-V5 with their GEMMs on the hi/lo tensor-core kernel (same contracts as the fp32 entry
 * points; `w16` = fp16 hi/lo weight pairs, see PackedAcoustic.enc16 / va16 in cmtts_b200/weights.py) */
size_t cmtts_encoder_tc_workspace_bytes(const cmtts_dims* d, int64_t B, int64_t T);
int cmtts_encoder_forward_tc(const cmtts_dims* d, const void* const* w, const void* const* w16,
                             const int64_t* tokens, const int64_t* src_lens, int64_t B, int64_t T,
                             float* enc_out, void* ws, size_t ws_bytes, void* stream);
size_t cmtts_variance_token_tc_workspace_bytes(const cmtts_dims* d, int64_t B, int64_t T);
int cmtts_variance_token_tc(const cmtts_dims* d, const void* const* w, const void* const* w16, const float* enc,
                            const int64_t* src_lens, const float* spker_embeds, float e_control,
                            float d_control, int64_t B, int64_t T, float* out1, float* log_d,
                            float* d_rounded, float* e_pred, int64_t* e_idx, int64_t* cumsum,
                            int64_t* mel_lens, float* spk_emb, float* f0_stats, void* ws, size_t ws_bytes,
                            void* stream);
size_t cmtts_variance_frame_tc_workspace_bytes(const cmtts_dims* d, int64_t B, int64_t L);
int cmtts_variance_frame_tc(const cmtts_dims* d, const void* const* w, const void* const* w16, const float* out1,
                            const int64_t* cumsum, const int64_t* mel_lens, const float* f0_stats,
                            float p_control, int64_t B, int64_t T, int64_t L, float* cond, int64_t* mel2ph,
                            float* cwt, float* f0_denorm, int64_t* pitch_idx, void* ws, size_t ws_bytes,
                            void* stream);

/* H1-H3 on tensor cores: `w16` = {pre_w fp32 [k][80][C0], pre_b, per level {up_w fp16 [taps*s*Cout][Cin],
 * up_b fp32, per resblock conv {w fp16 [k*C][C], b fp32}}, post_w fp32 [k][C], post_b}. */
size_t cmtts_hifigan_tc_workspace_bytes(const int32_t* cfg, int64_t B, int64_t L);
int cmtts_hifigan_forward_tc(const int32_t* cfg, const void* const* w16, const float* mel /* (B,L,80) */,
                             int64_t B, int64_t L, float* wav, int16_t* wav_i16, float max_wav_value,
                             void* ws, size_t ws_bytes, void* stream);

/* N4 (zero-shot path): DeepSpeaker ResCNN speaker encoder — replaces `DeepSpeakerModel().m.predict(mfcc[None])` of
 * deepspeaker/embedding.py:13-27 (model: deepspeaker/conv_models.py:44-138; caller: synthesize_zeroshot_lj.py:93-97).
 * cfg = {n_fbanks (64), first_filters (64, doubling over the 4 stages), dense_out (512)}.
 * x: (B, T, n_fbanks) fp32 per-frame-normalised filter-bank energies (one input channel; T = 160 in the reference).
 * w: per conv {kernel HWIO fp32, scale [Cout], shift [Cout]} (BatchNormalization(eps 1e-3) and the conv bias folded:
 *    scale = gamma / sqrt(var + eps), shift = beta + (bias - mean) * scale) for the 28 convs in network order — per stage
 *    the strided 5x5 conv, then each identity block's 2a and 2b 3x3 convs — then {dense kernel [feat][dense_out], bias}.
 * emb: (B, dense_out) fp32, L2-normalised (K.l2_normalize, eps 1e-12). */
size_t cmtts_rescnn_workspace_bytes(const int32_t* cfg, int64_t B, int64_t T);
int cmtts_rescnn_forward(const int32_t* cfg, const void* const* w, const float* x, int64_t B, int64_t T,
                         float* emb, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMTTS_B200_H */
