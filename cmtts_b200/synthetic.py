"""Seeded synthetic checkpoints and batches in the reference's layouts.

The reference ships no trained acoustic checkpoint (README.md:24), so parity and benchmarks run on
synthetic weights written in the reference's own checkpoint layout
(`<model_path>/CMDenoiserTTS/model{step:06d}.pt`, a flat `state_dict` of `CMTotalTTS`,
synthesize.py:44-48 / SURVEY.md §8 B4) and synthetic HiFi-GAN checkpoints in the layout of
`hifigan/generator_*.pth.tar` (`{"generator": sd}` with `weight_g` / `weight_v`,
utils/model.py:170-184).  Everything is drawn from `torch.Generator` on CPU so the same seed gives
the same bytes in the build container and on the GPU box.

Scales follow the reference's initialisers (blocks.py:10-23, :110-119, :188; torch defaults) with
three deliberate departures, so that the synthetic model exercises the whole path:
  * `net.output_projection.conv.weight` is random (the reference zero-initialises it,
    modules.py:598, which would make the denoiser output a constant);
  * the duration head has bias ln 7.25 and a weight scale giving durations of roughly 2..20
    frames per phoneme (L ~ 800 frames for ~115 phonemes);
  * energy / cwt / f0-stat heads are scaled so that predictions span the quantiser ranges.
"""
from __future__ import annotations

import hashlib
import math
import os
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

from .config import HifiGanSpec, ModelSpec


class _Draw:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def normal(self, shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=self.g, dtype=torch.float32) * std + mean

    def uniform(self, shape, bound):
        return (torch.rand(*shape, generator=self.g, dtype=torch.float32) * 2 - 1) * bound


def make_acoustic_state_dict(spec: ModelSpec, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Flat state_dict of the reference's CMTotalTTS (tts_net.py:40-47) for `spec`."""
    d = _Draw(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    H = spec.hidden
    F4 = 4 * H
    te = "duration_pitch_energy_net.text_encoder."
    for l in range(spec.enc_layers):
        p = f"{te}layers.{l}.op."
        sd[p + "layer_norm1.weight"] = d.normal((H,), 0.1, 1.0)
        sd[p + "layer_norm1.bias"] = d.normal((H,), 0.05)
        sd[p + "self_attn.in_proj_weight"] = d.uniform((3 * H, H), math.sqrt(6.0 / (4 * H)))
        sd[p + "self_attn.out_proj.weight"] = d.uniform((H, H), math.sqrt(6.0 / (2 * H)))
        sd[p + "layer_norm2.weight"] = d.normal((H,), 0.1, 1.0)
        sd[p + "layer_norm2.bias"] = d.normal((H,), 0.05)
        b = 1.0 / math.sqrt(H * spec.ffn_kernel)
        sd[p + "ffn.ffn_1.weight"] = d.uniform((F4, H, spec.ffn_kernel), b)
        sd[p + "ffn.ffn_1.bias"] = d.uniform((F4,), b)
        sd[p + "ffn.ffn_2.weight"] = d.uniform((H, F4), math.sqrt(6.0 / (H + F4)))
        sd[p + "ffn.ffn_2.bias"] = d.normal((H,), 0.01)
    sd[te + "layer_norm.weight"] = d.normal((H,), 0.1, 1.0)
    sd[te + "layer_norm.bias"] = d.normal((H,), 0.05)
    emb = d.normal((spec.vocab, H), H ** -0.5)
    emb[0] = 0.0
    sd[te + "embed_tokens.weight"] = emb
    sd[te + "embed_positions._float_tensor"] = torch.zeros(1)

    va = "duration_pitch_energy_net.variance_adaptor."
    sd[va + "energy_bins"] = torch.linspace(spec.energy_min, spec.energy_max, spec.energy_bins - 1)

    def predictor(prefix, idim, n_layers, k, odim, head_std, head_bias, with_pos):
        if with_pos:
            sd[prefix + "pos_embed_alpha"] = torch.tensor([0.9])
        for i in range(n_layers):
            cin = idim if i == 0 else spec.filter_size
            b = 1.0 / math.sqrt(cin * k)
            sd[f"{prefix}conv.{i}.1.weight"] = d.uniform((spec.filter_size, cin, k), b)
            sd[f"{prefix}conv.{i}.1.bias"] = d.uniform((spec.filter_size,), b)
            sd[f"{prefix}conv.{i}.3.weight"] = d.normal((spec.filter_size,), 0.1, 1.0)
            sd[f"{prefix}conv.{i}.3.bias"] = d.normal((spec.filter_size,), 0.05)
        sd[prefix + "linear.weight"] = d.normal((odim, spec.filter_size), head_std)
        sd[prefix + "linear.bias"] = head_bias.clone()
        if with_pos:
            sd[prefix + "embed_positions._float_tensor"] = torch.zeros(1)

    predictor(va + "duration_predictor.", H, spec.dur_layers, spec.dur_kernel, 1,
              0.022, torch.tensor([math.log(7.25)]), with_pos=False)
    b = 1.0 / math.sqrt(H)
    sd[va + "cwt_predictor.0.weight"] = d.uniform((spec.cwt_hidden, H), b)
    sd[va + "cwt_predictor.0.bias"] = d.uniform((spec.cwt_hidden,), b)
    cwt_bias = torch.zeros(spec.cwt_out)
    if spec.use_uv:
        cwt_bias[-1] = -0.5
    predictor(va + "cwt_predictor.1.", spec.cwt_hidden, spec.pred_layers, spec.pred_kernel,
              spec.cwt_out, 0.06, cwt_bias, with_pos=True)
    h = spec.cwt_hidden
    sd[va + "cwt_stats_layers.0.weight"] = d.uniform((h, H), 1.0 / math.sqrt(H))
    sd[va + "cwt_stats_layers.0.bias"] = d.uniform((h,), 1.0 / math.sqrt(H))
    sd[va + "cwt_stats_layers.2.weight"] = d.uniform((h, h), 1.0 / math.sqrt(h))
    sd[va + "cwt_stats_layers.2.bias"] = d.uniform((h,), 1.0 / math.sqrt(h))
    sd[va + "cwt_stats_layers.4.weight"] = d.uniform((2, h), 0.3 / math.sqrt(h))
    sd[va + "cwt_stats_layers.4.bias"] = torch.tensor([5.3, 0.35])
    pe = d.normal((spec.pitch_bins, H), H ** -0.5)
    pe[0] = 0.0
    sd[va + "pitch_embed.weight"] = pe
    predictor(va + "energy_predictor.", H, spec.pred_layers, spec.pred_kernel, 1,
              0.09, torch.tensor([2.0]), with_pos=True)
    ee = d.normal((spec.energy_bins, H), H ** -0.5)
    ee[0] = 0.0
    sd[va + "energy_embedding.weight"] = ee
    if spec.multi_speaker:
        if getattr(spec, "n_speakers", 0):            # nn.Embedding table (speaker_embedder: none)
            sd["duration_pitch_energy_net.speaker_emb.weight"] = d.normal((spec.n_speakers, H), 1.0)
        else:
            sd["duration_pitch_energy_net.speaker_emb.weight"] = d.normal((H, spec.ext_speaker_dim), 0.1)
            sd["duration_pitch_energy_net.speaker_emb.bias"] = d.normal((H,), 0.02)

    C = spec.res_channels
    M = spec.n_mels
    sd["net.input_projection.0.conv.weight"] = d.normal((C, M, 1), math.sqrt(2.0 / M))
    sd["net.input_projection.0.conv.bias"] = d.uniform((C,), 1.0 / math.sqrt(M))
    sd["net.mlp.0.linear.weight"] = d.uniform((4 * C, C), math.sqrt(6.0 / (5 * C)))
    sd["net.mlp.2.linear.weight"] = d.uniform((C, 4 * C), math.sqrt(6.0 / (5 * C)))
    for l in range(spec.res_layers):
        p = f"net.residual_layers.{l}."
        sd[p + "conv_layer.conv.weight"] = d.normal((2 * C, C, 3), math.sqrt(2.0 / (3 * C)))
        sd[p + "conv_layer.conv.bias"] = d.uniform((2 * C,), 1.0 / math.sqrt(3 * C))
        sd[p + "diffusion_projection.linear.weight"] = d.uniform((C, C), math.sqrt(6.0 / (2 * C)))
        if spec.multi_speaker:
            sd[p + "speaker_projection.linear.weight"] = d.uniform((C, H), math.sqrt(6.0 / (C + H)))
        sd[p + "conditioner_projection.conv.weight"] = d.normal((C, H, 1), math.sqrt(2.0 / H))
        sd[p + "conditioner_projection.conv.bias"] = d.uniform((C,), 1.0 / math.sqrt(H))
        sd[p + "output_projection.conv.weight"] = d.normal((2 * C, C, 1), math.sqrt(2.0 / C))
        sd[p + "output_projection.conv.bias"] = d.uniform((2 * C,), 1.0 / math.sqrt(C))
    sd["net.skip_projection.conv.weight"] = d.normal((C, C, 1), math.sqrt(2.0 / C))
    sd["net.skip_projection.conv.bias"] = d.uniform((C,), 1.0 / math.sqrt(C))
    sd["net.output_projection.conv.weight"] = d.normal((M, C, 1), 0.25)
    sd["net.output_projection.conv.bias"] = d.uniform((M,), 1.0 / math.sqrt(C))
    return sd


def make_hifigan_checkpoint(hspec: Optional[HifiGanSpec] = None, seed: int = 7) -> Dict[str, "OrderedDict[str, torch.Tensor]"]:
    """`{"generator": sd}` with weight-norm parametrisation, as hifigan/generator_*.pth.tar.
    Scales keep activations O(1) through the 4 up-sampling levels and 12 MRF ResBlocks."""
    hs = hspec or HifiGanSpec()
    d = _Draw(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def wn(prefix, shape, gain, n_bias):
        # torch's weight_norm state_dict order: bias, weight_g, weight_v.  The effective weight is
        # g * v / ||v|| (norm over dims != 0), so each dim-0 slice has L2 norm g.
        v = d.normal(shape, 0.01)
        jitter = (1.0 + 0.2 * d.normal((shape[0], 1, 1), 1.0)).abs()
        sd[prefix + "bias"] = d.normal((n_bias,), 0.05)
        sd[prefix + "weight_g"] = gain * jitter
        sd[prefix + "weight_v"] = v

    C0 = hs.upsample_initial_channel
    # Conv1d weight (Cout, Cin, k): g is (Cout,1,1); output std = g * rms(input)
    wn("conv_pre.", (C0, hs.n_mels, 7), 0.3, C0)
    for i, (u, k) in enumerate(zip(hs.upsample_rates, hs.upsample_kernel_sizes)):
        cin, cout = C0 // (2 ** i), C0 // (2 ** (i + 1))
        # ConvTranspose1d weight (Cin, Cout, k): g is (Cin,1,1), each slice spread over Cout*k
        # elements; an output sample sums Cin*(k/u) of them -> var = Cin g^2 / (u Cout)
        wn(f"ups.{i}.", (cin, cout, k), 1.2 * math.sqrt(u * cout / cin), cout)
    for i in range(len(hs.upsample_rates)):
        ch = C0 // (2 ** (i + 1))
        for j, (k, dil) in enumerate(zip(hs.resblock_kernel_sizes, hs.resblock_dilation_sizes)):
            r = i * len(hs.resblock_kernel_sizes) + j
            for m in range(len(dil)):
                wn(f"resblocks.{r}.convs1.{m}.", (ch, ch, k), 1.0, ch)
            for m in range(len(dil)):
                wn(f"resblocks.{r}.convs2.{m}.", (ch, ch, k), 0.5, ch)
    wn("conv_post.", (1, C0 // (2 ** len(hs.upsample_rates)), 7), 0.35, 1)
    return {"generator": sd}


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """`remove_weight_norm` (hifigan/models.py:167-174): w = g * v / ||v|| with the norm over all
    dims except dim 0 (torch.nn.utils.weight_norm default; dim 0 is Cin for ConvTranspose1d).
    Accepts already-folded dicts (`*.weight`) unchanged."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in sd.items():
        if k.endswith("weight_g"):
            base = k[: -len("weight_g")]
            g = v
            vv = sd[base + "weight_v"]
            nrm = torch.linalg.vector_norm(vv, ord=2, dim=tuple(range(1, vv.dim())), keepdim=True)
            out[base + "weight"] = vv * (g / nrm)
        elif k.endswith("weight_v"):
            continue
        else:
            out[k] = v
    return out


def make_batch(spec: ModelSpec, batch: int, src_lo: int, src_hi: int, seed: int = 1234
               ) -> Dict[str, Optional[torch.Tensor]]:
    """Synthetic inference batch in the layout of TextDataset.collate_fn / to_device
    (dataset.py:285-296, utils/tools.py:103-112): SURVEY.md §8(d)."""
    g = torch.Generator().manual_seed(seed)
    src_lens = torch.randint(src_lo, src_hi + 1, (batch,), generator=g, dtype=torch.int64)
    src_lens[0] = src_hi
    tmax = int(src_lens.max())
    texts = torch.randint(1, spec.vocab, (batch, tmax), generator=g, dtype=torch.int64)
    texts = texts * (torch.arange(tmax)[None, :] < src_lens[:, None]).long()
    speakers = torch.zeros(batch, dtype=torch.int64)
    spk = None
    if spec.multi_speaker:
        spk = torch.randn(batch, spec.ext_speaker_dim, generator=g, dtype=torch.float32)
        spk = spk / spk.norm(dim=1, keepdim=True)
    return {"speakers": speakers, "texts": texts, "src_lens": src_lens, "spker_embeds": spk}


def make_mels(batch: int, n_mels: int, frames: int, seed: int = 99) -> torch.Tensor:
    """C5 vocoder-only input: log-mel-like N(-5, 2^2) clipped to [-11.5, 2], (B, n_mels, L)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(batch, n_mels, frames, generator=g) * 2.0 - 5.0).clamp_(-11.5, 2.0)


def state_dict_digest(sd: Dict[str, torch.Tensor]) -> str:
    """sha256 over names, shapes and raw bytes — pins the RNG stream behind the golden fixtures."""
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(str(tuple(v.shape)).encode())
        h.update(v.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


def write_acoustic_checkpoint(root: str, spec: ModelSpec, seed: int = 0, step: int = 0) -> str:
    """Write `<root>/CMDenoiserTTS/model{step:06d}.pt` (synthesize.py:44-48)."""
    p = os.path.join(root, "CMDenoiserTTS")
    os.makedirs(p, exist_ok=True)
    f = os.path.join(p, "model{:06d}.pt".format(step))
    torch.save(make_acoustic_state_dict(spec, seed), f)
    return f


def make_deepspeaker_weights(seed: int = 0, scale: float = 1.0):
    """Random DeepSpeaker ResCNN weights in the naming / shapes of the reference's Keras checkpoint
    ({'<layer>/kernel:0', '<layer>_bn/gamma:0', ..., 'affine/kernel:0'}; deepspeaker/conv_models.py:83-131) — for tests and
    bench runs on boxes where the 97 MB checkpoint is not staged."""
    import numpy as np
    rng = np.random.default_rng(seed)
    w = {}
    cin = 1

    def conv(name, k, ci, co):
        w[f"{name}/kernel:0"] = (rng.standard_normal((k, k, ci, co)) * scale / np.sqrt(k * k * ci)).astype(np.float32)
        w[f"{name}/bias:0"] = (rng.standard_normal(co) * 0.05).astype(np.float32)
        w[f"{name}_bn/gamma:0"] = (1.0 + 0.2 * rng.standard_normal(co)).astype(np.float32)
        w[f"{name}_bn/beta:0"] = (0.3 * rng.standard_normal(co)).astype(np.float32)
        w[f"{name}_bn/moving_mean:0"] = (0.1 * rng.standard_normal(co)).astype(np.float32)
        w[f"{name}_bn/moving_variance:0"] = (0.5 + rng.random(co)).astype(np.float32)

    for stage, f in enumerate((64, 128, 256, 512), start=1):
        conv(f"conv{f}-s", 5, cin, f)
        for blk in range(3):
            conv(f"res{stage}_{blk}_branch_2a", 3, f, f)
            conv(f"res{stage}_{blk}_branch_2b", 3, f, f)
        cin = f
    w["affine/kernel:0"] = (rng.standard_normal((2048, 512)) / np.sqrt(2048)).astype(np.float32)
    w["affine/bias:0"] = (rng.standard_normal(512) * 0.05).astype(np.float32)
    return w


def make_voice_like(seconds: float = 2.5, sr: int = 22050, seed: int = 0):
    """Harmonic signal with a moving pitch, amplitude modulation, a little noise and a quiet head / tail (a stand-in for a
    reference recording of the zero-shot path) -> float32 in [-1, 1]."""
    import numpy as np
    rng = np.random.default_rng(seed)
    t = np.arange(int(seconds * sr)) / sr
    f0 = 120 + 40 * np.sin(2 * np.pi * 0.7 * t)
    ph = 2 * np.pi * np.cumsum(f0) / sr
    x = sum(np.sin(k * ph) / k for k in range(1, 12)) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t) ** 2)
    x = 0.2 * x / np.abs(x).max() + 0.002 * rng.standard_normal(t.size)
    x[: sr // 5] *= 0.01
    x[-sr // 5:] *= 0.01
    return x.astype(np.float32)
