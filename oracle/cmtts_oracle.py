"""ORACLE — CPU restatement of the CM-TTS inference hot path (TEST INFRASTRUCTURE, not product).

Only `tests/`, `oracle/make_golden.py`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module; `cmtts_b200/` never does, and
the product path raises if its CUDA library is missing.

What it is: a functional restatement, in plain torch-CPU tensor ops (F.conv1d, matmul, gather —
the same third-party arithmetic the reference itself delegates to), of the
reference's hot path, reading weights by key from a state_dict in the reference's checkpoint
layout.  Every function cites the reference file:line it follows (paths relative to the reference
repo root).  It is written independently of the reference's nn.Module classes so that it can
travel to the GPU box (where /root/reference does not exist).

Parity status: PINNED.  The reference ships no tests, golden vectors or trained acoustic
checkpoint, so the pin is the reference itself run in the build container:
`oracle/make_golden.py` imports the unmodified reference through `oracle/ref_shim.py`, runs it on
seeded synthetic checkpoints / real HiFi-GAN weights, and commits stage-boundary tensors under
`tests/golden/`; `tests/test_oracle_golden.py` checks this restatement against those fixtures
(and, when /root/reference is mounted, against the live reference).

dtype: fp32 by default (what the reference computes in); pass dtype=torch.float64 for an
error-free yardstick when judging which of two fp32 results is closer to the truth.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

TE = "duration_pitch_energy_net.text_encoder."
VA = "duration_pitch_energy_net.variance_adaptor."
SPK = "duration_pitch_energy_net.speaker_emb."


class Weights:
    """state_dict accessor with optional dtype promotion."""

    def __init__(self, sd: Dict[str, torch.Tensor], dtype=torch.float32):
        self.sd = sd
        self.dtype = dtype
        self._cache: Dict[str, torch.Tensor] = {}

    def __call__(self, key: str) -> torch.Tensor:
        t = self._cache.get(key)
        if t is None:
            t = self.sd[key].detach().to("cpu")
            if t.is_floating_point():
                t = t.to(self.dtype)
            self._cache[key] = t
        return t

    def has(self, key: str) -> bool:
        return key in self.sd


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def get_mask_from_lengths(lengths: torch.Tensor, max_len: Optional[int] = None) -> torch.Tensor:
    """utils/tools.py:275-283 — True = padding."""
    if max_len is None:
        max_len = int(lengths.max())
    ids = torch.arange(0, max_len)[None, :]
    return ids >= lengths[:, None]


def sinusoid_table(n: int, dim: int, dtype=torch.float32) -> torch.Tensor:
    """model/blocks.py:44-60 — [sin | cos] halves, row 0 (padding_idx) zeroed.  Always built in
    fp32 like the reference, then promoted."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(n, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    tab = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(n, -1)
    tab[0, :] = 0
    return tab.to(dtype)


def make_positions(x: torch.Tensor, padding_idx: int = 0) -> torch.Tensor:
    """utils/tools.py:810-822 — 1-based running index of non-pad entries, pad -> padding_idx."""
    mask = x.ne(padding_idx).int()
    return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + padding_idx


def positional(x_first_channel: torch.Tensor, dim: int, dtype) -> torch.Tensor:
    """model/blocks.py:62-81 — SinusoidalPositionalEmbedding.forward."""
    b, t = x_first_channel.shape[:2]
    tab = sinusoid_table(t + 2, dim, dtype)
    pos = make_positions(x_first_channel, 0)
    return tab.index_select(0, pos.view(-1)).view(b, t, -1)


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


# --------------------------------------------------------------------------------------------
# E1-E4: text encoder
# --------------------------------------------------------------------------------------------
def mha(W: Weights, prefix: str, x_tbc: torch.Tensor, key_padding: torch.Tensor, heads: int):
    """model/blocks.py:303-312 -> F.multi_head_attention_forward with in_proj_weight, no biases,
    bool key_padding_mask (-> -inf), q scaled by head_dim**-0.5; out_proj without bias."""
    T, B, C = x_tbc.shape
    d = C // heads
    qkv = x_tbc @ W(prefix + "in_proj_weight").t()
    q, k, v = qkv.chunk(3, dim=-1)
    q = q.contiguous().view(T, B * heads, d).transpose(0, 1) * (1.0 / math.sqrt(d))
    k = k.contiguous().view(T, B * heads, d).transpose(0, 1)
    v = v.contiguous().view(T, B * heads, d).transpose(0, 1)
    scores = torch.bmm(q, k.transpose(1, 2)).view(B, heads, T, T)
    scores = scores.masked_fill(key_padding[:, None, None, :], float("-inf"))
    p = torch.softmax(scores, dim=-1).view(B * heads, T, T)
    o = torch.bmm(p, v).transpose(0, 1).contiguous().view(T, B, C)
    return o @ W(prefix + "out_proj.weight").t()


def ffn(W: Weights, prefix: str, x_tbc: torch.Tensor, kernel: int, act: str):
    """model/blocks.py:533-552 — Conv1d(k, SAME) * k**-0.5 -> act -> Linear."""
    h = F.conv1d(x_tbc.permute(1, 2, 0), W(prefix + "ffn_1.weight"), W(prefix + "ffn_1.bias"),
                 padding=kernel // 2).permute(2, 0, 1)
    h = h * kernel ** -0.5
    if act == "gelu":
        h = F.gelu(h)
    elif act == "relu":
        h = F.relu(h)
    elif act == "swish":
        h = h * torch.sigmoid(h)
    return F.linear(h, W(prefix + "ffn_2.weight"), W(prefix + "ffn_2.bias"))


def encoder(W: Weights, spec, texts: torch.Tensor, src_mask: torch.Tensor) -> torch.Tensor:
    """FastspeechEncoder.forward, model/modules.py:132-151 + FFTBlocks.forward :80-105 +
    EncSALayer.forward blocks.py:594-618.  Returns (B, T, C)."""
    C = spec.hidden
    x = math.sqrt(C) * F.embedding(texts, W(TE + "embed_tokens.weight"))
    x = x + positional(texts, C, W.dtype)
    nonpad = 1 - src_mask.transpose(0, 1).to(W.dtype)[:, :, None]  # (T,B,1)
    x = x.transpose(0, 1) * nonpad
    for l in range(spec.enc_layers):
        p = f"{TE}layers.{l}.op."
        r = x
        h = layer_norm(x, W(p + "layer_norm1.weight"), W(p + "layer_norm1.bias"), 1e-12)
        h = mha(W, p + "self_attn.", h, src_mask, spec.enc_heads)
        x = (r + h) * nonpad
        r = x
        h = layer_norm(x, W(p + "layer_norm2.weight"), W(p + "layer_norm2.bias"), 1e-12)
        h = ffn(W, p + "ffn.", h, spec.ffn_kernel, spec.ffn_act)
        x = (r + h) * nonpad
        x = x * nonpad  # FFTBlocks.forward :96 multiplies again
    x = layer_norm(x, W(TE + "layer_norm.weight"), W(TE + "layer_norm.bias"), 1e-5) * nonpad
    return x.transpose(0, 1)


# --------------------------------------------------------------------------------------------
# V1-V5: variance adaptor
# --------------------------------------------------------------------------------------------
def predictor_convs(W: Weights, prefix: str, x_bct: torch.Tensor, n_layers: int, k: int,
                    mask_keep: Optional[torch.Tensor]) -> torch.Tensor:
    """modules.py:477-487 / :527-537 — [ConstantPad1d -> Conv1d -> ReLU -> LayerNorm(dim=1, eps
    1e-12)] stacks; the duration predictor multiplies by (1-mask) after each (modules.py:500-503)."""
    for i in range(n_layers):
        x_bct = F.pad(x_bct, ((k - 1) // 2, (k - 1) // 2))
        x_bct = F.conv1d(x_bct, W(f"{prefix}conv.{i}.1.weight"), W(f"{prefix}conv.{i}.1.bias"))
        x_bct = F.relu(x_bct)
        x_bct = layer_norm(x_bct.transpose(1, -1), W(f"{prefix}conv.{i}.3.weight"),
                           W(f"{prefix}conv.{i}.3.bias"), 1e-12).transpose(1, -1)
        if mask_keep is not None:
            x_bct = x_bct * mask_keep[:, None, :]
    return x_bct


def duration_predictor(W: Weights, spec, x: torch.Tensor, src_mask: torch.Tensor) -> torch.Tensor:
    """DurationPredictor.forward, modules.py:498-509 -> (B, T) log-durations, 0 at pads."""
    keep = 1 - src_mask.to(W.dtype)
    h = predictor_convs(W, VA + "duration_predictor.", x.transpose(1, -1), spec.dur_layers,
                        spec.dur_kernel, keep)
    h = F.linear(h.transpose(1, -1), W(VA + "duration_predictor.linear.weight"),
                 W(VA + "duration_predictor.linear.bias"))
    h = h * keep[:, :, None]
    return h.squeeze(-1)


def pitch_style_predictor(W: Weights, prefix: str, x: torch.Tensor, n_layers: int, k: int):
    """PitchPredictor.forward (also EnergyPredictor), modules.py:542-555 — adds
    alpha * sinusoid(position of x[...,0] != 0), conv stack WITHOUT masking, Linear."""
    pos = W(prefix + "pos_embed_alpha") * positional(x[..., 0], x.shape[-1], W.dtype)
    h = x + pos
    h = predictor_convs(W, prefix, h.transpose(1, -1), n_layers, k, None)
    return F.linear(h.transpose(1, -1), W(prefix + "linear.weight"), W(prefix + "linear.bias"))


def round_durations(log_d: torch.Tensor, d_control: float = 1.0) -> torch.Tensor:
    """modules.py:369-372 — clamp(round(exp(log_d) - 1) * d_control, min=0); torch.round is
    round-half-to-even."""
    return torch.clamp(torch.round(torch.exp(log_d) - 1) * d_control, min=0)


def dur_to_mel2ph(dur: torch.Tensor, src_mask: torch.Tensor, L: Optional[int] = None) -> torch.Tensor:
    """utils/tools.py:768-798 restated as prefix-scan + binary search:
    mel2ph[b, t] = 1 + #{i : cumsum_i <= t} for t < sum(dur[b]), else 0.  Width = max_b sum(dur)
    (tools.py:795) unless L is given."""
    d = torch.round(dur.float()).long() * (1 - src_mask.long())
    cs = torch.cumsum(d, 1)
    width = int(d.sum(-1).max()) if L is None else L
    t = torch.arange(width)[None, :].expand(d.shape[0], -1).contiguous()
    idx = torch.searchsorted(cs, t, right=True) + 1
    return idx * (t < cs[:, -1:]).long()


def dur_to_mel2ph_literal(dur: torch.Tensor, src_mask: torch.Tensor) -> torch.Tensor:
    """utils/tools.py:788-798 as written (B x T x L boolean cube) — used to pin the scan form."""
    d = torch.round(dur.float()).long() * (1 - src_mask.long())
    token_idx = torch.arange(1, d.shape[1] + 1)[None, :, None]
    cs = torch.cumsum(d, 1)
    prev = F.pad(cs, [1, -1], mode="constant", value=0)
    pos = torch.arange(int(d.sum(-1).max()))[None, None]
    m = (pos >= prev[:, :, None]) & (pos < cs[:, :, None])
    return (token_idx * m.long()).sum(1)


def length_regulate(x: torch.Tensor, dur: torch.Tensor, max_len: Optional[int]
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """LengthRegulator.LR/expand, modules.py:421-444 + pad tools.py:724-742, restated as a
    gather: out[b, t] = x[b, searchsorted(cumsum(d_b), t, right)] for t < mel_len[b], else 0.
    NOTE the reference does NOT zero the durations of padded tokens here (expand uses
    `predicted[i]` as is); they are 0 anyway because log_d is masked (modules.py:505-506)."""
    d = torch.clamp(dur, min=0).to(torch.int64)  # max(int(expand_size), 0), modules.py:441
    cs = torch.cumsum(d, 1)
    mel_len = cs[:, -1].clone()
    L = int(mel_len.max()) if not max_len else int(max_len)
    t = torch.arange(L)[None, :].expand(d.shape[0], -1).contiguous()
    idx = torch.searchsorted(cs, t, right=True).clamp(max=x.shape[1] - 1)
    out = torch.gather(x, 1, idx[:, :, None].expand(-1, -1, x.shape[2]))
    out = out * (t < mel_len[:, None]).to(x.dtype)[:, :, None]
    return out, mel_len


def length_regulate_literal(x: torch.Tensor, dur: torch.Tensor, max_len: Optional[int]
                            ) -> Tuple[torch.Tensor, torch.Tensor]:
    """modules.py:421-444 as written: per-token `.item()` + expand + cat + pad.  This is the form
    the CPU baseline times (it is what the reference executes)."""
    outs, lens = [], []
    for b in range(x.shape[0]):
        rows = []
        for i in range(x.shape[1]):
            n = max(int(dur[b, i].item()), 0)
            rows.append(x[b, i].expand(n, -1))
        e = torch.cat(rows, 0)
        outs.append(e)
        lens.append(e.shape[0])
    L = max_len if max_len else max(lens)
    out = torch.stack([F.pad(e, (0, 0, 0, L - e.shape[0])) for e in outs])
    return out, torch.tensor(lens, dtype=torch.int64)


def f0_to_coarse(f0: torch.Tensor) -> torch.Tensor:
    """utils/pitch_tools.py:26-35 — the numpy-float64 module constants enter fp32 tensor ops as
    Python scalars (cast to the tensor dtype), left-to-right."""
    f0_bin = 256
    f0_mel_min = 1127 * np.log(1 + 50.0 / 700)
    f0_mel_max = 1127 * np.log(1 + 1100.0 / 700)
    f0_mel = 1127 * (1 + f0 / 700).log()
    pos = f0_mel > 0
    f0_mel[pos] = (f0_mel[pos] - f0_mel_min) * (f0_bin - 2) / (f0_mel_max - f0_mel_min) + 1
    f0_mel[f0_mel <= 1] = 1
    f0_mel[f0_mel > f0_bin - 1] = f0_bin - 1
    return (f0_mel + 0.5).long()


def cwt_to_f0_norm(cwt_spec: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, eps: float):
    """cwt2f0_norm, utils/pitch_tools.py:274-279 -> cwt2f0 :261-272 -> inverse_cwt_torch :244-250
    -> norm_f0 :38-47 (log).  The standardisation runs over the PADDED time axis, unbiased std."""
    n = cwt_spec.shape[-1]
    b = (torch.arange(0, n).float()[None, None, :] + 1 + 2.5) ** (-2.5)
    rec = (cwt_spec * b.to(cwt_spec.dtype)).sum(-1)
    rec = (rec - rec.mean(-1, keepdim=True)) / rec.std(-1, keepdim=True)
    f0 = (rec * std[:, None] + mean[:, None]).exp()
    return torch.log2(f0 + eps)


def dpen(W: Weights, spec, speakers, texts, src_lens, spker_embeds=None, max_mel_len: Optional[int] = None,
         p_control=1.0, e_control=1.0, d_control=1.0, literal_lr: bool = False) -> Dict[str, torch.Tensor]:
    """DurationPitchSpeakerNet.forward (inference branch), model/cmtts.py:44-122 +
    VarianceAdaptor.forward modules.py:331-412.  `max_mel_len` is what CMTotalTTS.forward passes
    when it re-runs this inside a solver step (tts_net.py:132-147: mels=x -> max_mel_len = L)."""
    B, T = texts.shape
    src_mask = get_mask_from_lengths(src_lens, T)
    enc = encoder(W, spec, texts, src_mask)
    spk = None
    x = enc
    if spec.multi_speaker:
        assert spker_embeds is not None, "Speaker embedding should not be None"  # cmtts.py:80
        spk = F.linear(spker_embeds.to(W.dtype), W(SPK + "weight"), W(SPK + "bias"))
        x = x + spk.unsqueeze(1).expand(-1, T, -1)
    log_d = duration_predictor(W, spec, x, src_mask)
    e_pred = pitch_style_predictor(W, VA + "energy_predictor.", x, spec.pred_layers,
                                   spec.pred_kernel).squeeze(-1) * e_control
    e_idx = torch.bucketize(e_pred, W(VA + "energy_bins"))
    out1 = x + F.embedding(e_idx, W(VA + "energy_embedding.weight"))
    d_rounded = round_durations(log_d, d_control)
    mel2ph = dur_to_mel2ph(d_rounded, src_mask)
    lr = length_regulate_literal if literal_lr else length_regulate
    xf, mel_len = lr(out1, d_rounded, max_mel_len)
    mel_mask = get_mask_from_lengths(mel_len)
    # get_pitch_embedding, cwt branch, modules.py:273-307
    h = F.linear(xf, W(VA + "cwt_predictor.0.weight"), W(VA + "cwt_predictor.0.bias"))
    cwt = pitch_style_predictor(W, VA + "cwt_predictor.1.", h, spec.pred_layers, spec.pred_kernel) * p_control
    st = out1[:, 0, :]
    st = F.relu(F.linear(st, W(VA + "cwt_stats_layers.0.weight"), W(VA + "cwt_stats_layers.0.bias")))
    st = F.relu(F.linear(st, W(VA + "cwt_stats_layers.2.weight"), W(VA + "cwt_stats_layers.2.bias")))
    st = F.linear(st, W(VA + "cwt_stats_layers.4.weight"), W(VA + "cwt_stats_layers.4.bias"))
    f0_mean, f0_std = st[:, 0], st[:, 1]
    f0 = cwt_to_f0_norm(cwt[:, :, :10], f0_mean, f0_std * spec.cwt_std_scale, spec.pitch_norm_eps)
    if mel2ph.shape[1] > f0.shape[1]:  # pitch_tools.py:276-277
        f0 = torch.cat([f0] + [f0[:, -1:]] * (mel2ph.shape[1] - f0.shape[1]), 1)
    f0_denorm = 2 ** f0  # denorm_f0, pitch_tools.py:64-78
    if spec.use_uv:
        uv = cwt[:, :, -1] > 0
        n = min(f0_denorm.shape[1], uv.shape[1])
        f0_denorm[:, :n][uv[:, :n]] = 0
    pitch = f0_to_coarse(f0_denorm.clone())
    cond = xf + F.embedding(pitch, W(VA + "pitch_embed.weight"))
    return {
        "cond": cond, "enc": enc, "log_d_predictions": log_d, "e_predictions": e_pred,
        "e_idx": e_idx, "d_rounded": d_rounded, "mel2ph": mel2ph, "mel_lens": mel_len,
        "mel_masks": mel_mask, "src_masks": src_mask, "speaker_emb": spk, "src_lens": src_lens,
        "cwt": cwt, "f0_mean": f0_mean, "f0_std": f0_std, "f0_denorm": f0_denorm,
        "pitch_idx": pitch, "frame_feats": xf,
    }


# --------------------------------------------------------------------------------------------
# D1-D3: denoiser
# --------------------------------------------------------------------------------------------
def mish(x):
    return x * torch.tanh(F.softplus(x))


def step_embedding(W: Weights, spec, t: torch.Tensor) -> torch.Tensor:
    """DiffusionEmbedding.forward blocks.py:633-640 + Denoiser.mlp modules.py:579-583."""
    half = spec.res_channels // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half) * -e)            # fp32 like the reference
    e = t.to(torch.float32)[:, None] * e[None, :]
    e = torch.cat((e.sin(), e.cos()), dim=-1).to(W.dtype)
    h = F.linear(e, W("net.mlp.0.linear.weight"))
    return F.linear(mish(h), W("net.mlp.2.linear.weight"))


def denoiser(W: Weights, spec, mel_b1lm: torch.Tensor, t: torch.Tensor, cond_blc: torch.Tensor,
             spk: Optional[torch.Tensor]) -> torch.Tensor:
    """CMTotalTTS.forward tail tts_net.py:152-157 + Denoiser.forward modules.py:600-638 +
    ResidualBlock.forward blocks.py:667-686.  In/out layout (B,1,L,M) like the sampler's x."""
    x = mel_b1lm[:, 0].transpose(1, 2)                 # (B, M, L)
    c = cond_blc.transpose(1, 2)                       # (B, 256, L)
    x = F.relu(F.relu(F.conv1d(x, W("net.input_projection.0.conv.weight"),
                               W("net.input_projection.0.conv.bias"))))
    s = step_embedding(W, spec, t)
    skip_sum = None
    for l in range(spec.res_layers):
        p = f"net.residual_layers.{l}."
        ds = F.linear(s, W(p + "diffusion_projection.linear.weight")).unsqueeze(-1)
        cc = F.conv1d(c, W(p + "conditioner_projection.conv.weight"), W(p + "conditioner_projection.conv.bias"))
        res = y = x + ds
        y = y + cc
        if spec.multi_speaker:
            y = y + F.linear(spk, W(p + "speaker_projection.linear.weight")).unsqueeze(-1)
        y = F.conv1d(y, W(p + "conv_layer.conv.weight"), W(p + "conv_layer.conv.bias"), padding=1)
        gate, filt = torch.chunk(y, 2, dim=1)
        y = torch.sigmoid(gate) * torch.tanh(filt)
        y = F.conv1d(y, W(p + "output_projection.conv.weight"), W(p + "output_projection.conv.bias"))
        xo, skip = torch.chunk(y, 2, dim=1)
        x = (xo + res) / math.sqrt(2.0)
        skip_sum = skip if skip_sum is None else skip_sum + skip
    x = skip_sum / math.sqrt(spec.res_layers)
    x = F.relu(F.conv1d(x, W("net.skip_projection.conv.weight"), W("net.skip_projection.conv.bias")))
    x = F.conv1d(x, W("net.output_projection.conv.weight"), W("net.output_projection.conv.bias"))
    return x.transpose(1, 2)[:, None]                  # (B,1,L,M)


# --------------------------------------------------------------------------------------------
# S2-S4: consistency sampler
# --------------------------------------------------------------------------------------------
def scalings_for_boundary_condition(sigma: torch.Tensor, sigma_min: float, sigma_data: float):
    """karras_diffusion.py:87-102."""
    c_skip = sigma_data ** 2 / ((sigma - sigma_min) ** 2 + sigma_data ** 2)
    c_out = (sigma - sigma_min) * sigma_data / (sigma ** 2 + sigma_data ** 2) ** 0.5
    c_in = 1 / (sigma ** 2 + sigma_data ** 2) ** 0.5
    return c_skip, c_out, c_in


def get_sigmas_karras(n: int, sigma_min: float, sigma_max: float, rho: float = 7.0) -> torch.Tensor:
    """karras_diffusion.py:580-586."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def sampler_plan(T: int) -> Tuple[str, int, Optional[Tuple[int, ...]]]:
    """synthesize.py:106-146 — T -> (sampler, steps, ts)."""
    if T == 1:
        return "onestep", 2, None
    if T == 2:
        return "multistep", 2, (0, 0, 1)
    if T == 4:
        return "multistep", 2, (0, 0, 0, 0, 1)
    raise ValueError(f"T must be 1, 2 or 4 (synthesize.py:106-146), got {T}")


def sample(W: Weights, spec, batch: Dict[str, torch.Tensor], T: int,
           randn: Callable[[Tuple[int, ...]], torch.Tensor], literal: bool = False,
           trace: Optional[dict] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """CMTotalTTSSynthesize.synthesize synthesize.py:88-153 -> karras_sample_tts
    karras_diffusion.py:480-577 -> sample_onestep :800-811 | stochastic_iterative_sampler :829-854
    -> KarrasDenoiser.denoise :392-407.

    literal=True follows the reference's schedule exactly (encoder + variance adaptor re-run
    inside every solver step via CMTotalTTS.forward, Python-loop length regulator) — used for
    the CPU baseline timing; literal=False computes the (bit-identical in eval mode)
    conditioner once.  `randn(shape)` supplies x_T then one tensor per re-noise, in order.
    Returns (mel (B,L,M), dpen dict)."""
    kw = dict(speakers=batch["speakers"], texts=batch["texts"], src_lens=batch["src_lens"],
              spker_embeds=batch.get("spker_embeds"))
    pre = dpen(W, spec, literal_lr=literal, **kw)      # synthesize.py:102 pre-pass
    B, L, _ = pre["cond"].shape
    sampler, steps, ts = sampler_plan(T)
    sigmas = get_sigmas_karras(steps, spec.sigma_min, spec.sigma_max, spec.rho)
    x = randn((B, 1, L, spec.n_mels)).to(W.dtype) * spec.sigma_max

    def distiller(x_t: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
        c_skip, c_out, c_in = [v[:, None, None, None].to(W.dtype) for v in
                               scalings_for_boundary_condition(sigma, spec.sigma_min, spec.sigma_data)]
        rescaled_t = 1000 * 0.25 * torch.log(sigma + 1e-44)
        d = dpen(W, spec, max_mel_len=x_t.shape[2], literal_lr=True, **kw) if literal else pre
        model_output = denoiser(W, spec, c_in * x_t, rescaled_t, d["cond"], d["speaker_emb"])
        if trace is not None:
            trace.setdefault("model_output", []).append(model_output)
        return c_out * model_output + c_skip * x_t

    s_in = torch.ones([B], dtype=torch.float32)
    if sampler == "onestep":
        x0 = distiller(x, sigmas[0] * s_in)
    else:
        t_max_rho = spec.sigma_max ** (1 / spec.rho)
        t_min_rho = spec.sigma_min ** (1 / spec.rho)
        for i in range(len(ts) - 1):
            t = (t_max_rho + ts[i] / (steps - 1) * (t_min_rho - t_max_rho)) ** spec.rho
            x0 = distiller(x, t * s_in)
            next_t = (t_max_rho + ts[i + 1] / (steps - 1) * (t_min_rho - t_max_rho)) ** spec.rho
            next_t = np.clip(next_t, spec.sigma_min, spec.sigma_max)
            x = x0 + randn(tuple(x.shape)).to(W.dtype) * np.sqrt(next_t ** 2 - spec.sigma_min ** 2) * 0.85
        x0 = x
    return x0[:, 0], pre


# --------------------------------------------------------------------------------------------
# H1-H3: HiFi-GAN generator
# --------------------------------------------------------------------------------------------
def hifigan(Wf: Weights, hspec, mel_bml: torch.Tensor) -> torch.Tensor:
    """hifigan.Generator.forward hifigan/models.py:149-165 + ResBlock.forward :96-103, on
    weight-norm-folded weights (remove_weight_norm :167-174).  (B,80,L) -> (B,1,hop*L)."""
    x = F.conv1d(mel_bml.to(Wf.dtype), Wf("conv_pre.weight"), Wf("conv_pre.bias"), padding=3)
    nk = len(hspec.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(hspec.upsample_rates, hspec.upsample_kernel_sizes)):
        x = F.leaky_relu(x, hspec.lrelu_slope)
        x = F.conv_transpose1d(x, Wf(f"ups.{i}.weight"), Wf(f"ups.{i}.bias"), stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, dils) in enumerate(zip(hspec.resblock_kernel_sizes, hspec.resblock_dilation_sizes)):
            r = i * nk + j
            y = x
            for m, dil in enumerate(dils):
                t = F.leaky_relu(y, hspec.lrelu_slope)
                t = F.conv1d(t, Wf(f"resblocks.{r}.convs1.{m}.weight"), Wf(f"resblocks.{r}.convs1.{m}.bias"),
                             dilation=dil, padding=(rk * dil - dil) // 2)
                t = F.leaky_relu(t, hspec.lrelu_slope)
                t = F.conv1d(t, Wf(f"resblocks.{r}.convs2.{m}.weight"), Wf(f"resblocks.{r}.convs2.{m}.bias"),
                             padding=(rk - 1) // 2)
                y = t + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01, hifigan/models.py:161
    x = F.conv1d(x, Wf("conv_post.weight"), Wf("conv_post.bias"), padding=3)
    return torch.tanh(x)


def wav_to_int16(wav_b1n: torch.Tensor, lengths: Optional[Sequence[int]], max_wav_value: float = 32768.0
                 ) -> List[np.ndarray]:
    """vocoder_infer, utils/model.py:195-203 — (wav.numpy() * 32768).astype(int16), crop."""
    w = (wav_b1n.squeeze(1).to(torch.float32).cpu().numpy() * max_wav_value).astype("int16")
    out = [w[i] for i in range(w.shape[0])]
    if lengths is not None:
        out = [o[: int(n)] for o, n in zip(out, lengths)]
    return out


def synthesize(W: Weights, Wf: Weights, spec, batch, T: int, randn, literal: bool = False):
    """Whole path as p_rtf_cm.py:174-226 strings it together: sample -> (B,80,L) -> HiFi-GAN ->
    int16 crop to mel_len * hop."""
    mel, pre = sample(W, spec, batch, T, randn, literal=literal)
    wav = hifigan(Wf, spec.hifigan, mel.transpose(1, 2))
    lens = (pre["mel_lens"] * spec.hop_length).tolist()
    return mel, wav, wav_to_int16(wav, lens, spec.max_wav_value), pre
