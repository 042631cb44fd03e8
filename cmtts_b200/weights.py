"""Checkpoint layout -> device-resident, kernel-ready weight tables.

Input is exactly what the reference loads: the flat `state_dict` of `CMTotalTTS`
(`<model_path>/CMDenoiserTTS/model{step:06d}.pt`, synthesize.py:44-48, :79-83) and the HiFi-GAN
checkpoint `{"generator": state_dict}` with `weight_g`/`weight_v` (utils/model.py:170-184).
Output are fp32 tensors in HBM re-packed for the channels-last implicit-GEMM kernels
(conv weights as [tap][Cin][Cout]) plus ctypes pointer tables in the order
csrc/pipeline.cu expects (include/cmtts_b200.h).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .config import F0_MEL_MAX, F0_MEL_MIN, HifiGanSpec, ModelSpec
from .synthetic import fold_weight_norm

TE = "duration_pitch_energy_net.text_encoder."
VA = "duration_pitch_energy_net.variance_adaptor."
SPK = "duration_pitch_energy_net.speaker_emb."

ACT_CODES = {"relu": 1, "gelu": 2, "swish": 6}


def sinusoid_table(n: int, dim: int) -> torch.Tensor:
    """SinusoidalPositionalEmbedding.get_embedding (model/blocks.py:44-60), fp32 on CPU so the
    table holds the same bits the reference indexes; row 0 (padding_idx) is zero."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(n, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    tab = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(n, -1)
    if dim % 2 == 1:
        tab = torch.cat([tab, torch.zeros(n, 1)], dim=1)
    tab[0, :] = 0
    return tab


def conv_w(w: torch.Tensor) -> torch.Tensor:
    """(Cout, Cin, k) -> [k][Cin][Cout]."""
    return w.permute(2, 1, 0).contiguous()


def lin_w(w: torch.Tensor) -> torch.Tensor:
    """(out, in) -> [in][out]."""
    return w.t().contiguous()


def gate_permutation(C: int, tile: int = 64) -> torch.Tensor:
    """Column order for the gated k=3 conv: each 128-wide GEMM tile holds 64 gate columns followed
    by the 64 matching filter columns (torch.chunk(y, 2, dim=1) pairs channel c with c + C,
    model/blocks.py:680)."""
    idx = []
    for t in range(C // tile):
        idx += list(range(t * tile, (t + 1) * tile))            # gates
        idx += list(range(C + t * tile, C + (t + 1) * tile))    # filters
    return torch.tensor(idx, dtype=torch.long)


def pack_conv_transpose(w: torch.Tensor, bias: torch.Tensor, stride: int, padding: int):
    """ConvTranspose1d(Cin, Cout, k, stride u, padding p) as an ordinary conv with u*Cout output
    channels: out[q*u + r, co] = sum_{delta, ci} x[q + delta, ci] * w[ci, co, r + p - delta*u].
    Returns (packed [taps][Cin][u*Cout], bias replicated u times, first delta)."""
    cin, cout, k = w.shape
    u, p = stride, padding
    if (k - u) != 2 * p:
        raise NotImplementedError("ConvTranspose1d packing needs k - stride == 2 * padding (output length = stride * L)")
    deltas = [dl for dl in range(-k, k + 1) if any(0 <= r + p - dl * u < k for r in range(u))]
    d0, d1 = min(deltas), max(deltas)
    out = torch.zeros(d1 - d0 + 1, cin, u * cout, dtype=w.dtype)
    for dl in range(d0, d1 + 1):
        for r in range(u):
            j = r + p - dl * u
            if 0 <= j < k:
                out[dl - d0, :, r * cout:(r + 1) * cout] = w[:, :, j]
    return out.contiguous(), bias.repeat(u).contiguous(), d0


def conv_w_nk(w: torch.Tensor) -> torch.Tensor:
    """(Cout, Cin, k) -> [k][Cout][Cin]: K-major B operand of the tcgen05 path (rows = tap*Cout + n)."""
    return w.permute(2, 0, 1).contiguous()


def pair_pack_d1(w_nk: torch.Tensor) -> torch.Tensor:
    """Dilation-1 conv weights [k][32][32] (conv_w_nk) for the paired-row ResBlock kernel (csrc/umma_resblock.cu, MODE 1
    and every conv2): a row of the (L/2, 64) view holds the time steps 2r, 2r+1 and yields both outputs.  Unit u = input
    position u - h (h = (k-1)/2) meets the N = 64 block [W[u] ; W[u-1]] (rows [0,32) -> output 2r, rows [32,64) -> output
    2r+1; zero where the tap does not exist).  Two units share one 128-byte-row block: -> [(k+1)/2][64][64]."""
    k, co, ci = w_nk.shape
    if co != 32 or ci != 32 or k % 2 == 0:
        raise ValueError("pair_pack_d1: expects [k odd][32][32]")
    out = torch.zeros((k + 1) // 2, 64, 64, dtype=w_nk.dtype)
    for u in range(k + 1):
        c0 = (u & 1) * 32
        if u < k:
            out[u >> 1, :32, c0:c0 + 32] = w_nk[u]
        if u >= 1:
            out[u >> 1, 32:, c0:c0 + 32] = w_nk[u - 1]
    return out.contiguous()


def pair_pack_taps(w_nk: torch.Tensor) -> torch.Tensor:
    """Conv weights [k][32][32] for conv1 of the paired-row kernel at odd dilations > 1 (MODE 2): the plain per-tap blocks,
    two taps per 128-byte-row block -> [(k+1)/2][32][64] (tap j at block j >> 1, columns (j & 1) * 32 ...)."""
    k, co, ci = w_nk.shape
    if co != 32 or ci != 32:
        raise ValueError("pair_pack_taps: expects [k][32][32]")
    out = torch.zeros((k + 1) // 2, 32, 64, dtype=w_nk.dtype)
    for j in range(k):
        out[j >> 1, :, (j & 1) * 32:(j & 1) * 32 + 32] = w_nk[j]
    return out.contiguous()


#: power-of-two pre-scale of hi/lo weight pairs: keeps the `lo` halves (|lo| ~ 2^-11 |w|) out of the fp16
#: subnormal range, where their absolute spacing (6e-8) would cap the pair at ~18 bits for |w| ~ 0.01.
#: The kernels multiply the accumulator by 1 / TC_W_SCALE (csrc/pipeline.cu: TC_W_SCALE).
TC_W_SCALE = 1024.0


def split_f16(w: torch.Tensor, scale: float = TC_W_SCALE):
    """fp32 -> (hi, lo) fp16 with (hi + lo) / scale == w to ~22 bits (hi = fp16(w s), lo = fp16(w s - hi))."""
    wide = torch.float64 if w.dtype == torch.float64 else torch.float32   # fp64 in: weights derived on the host
    ws = w.to(wide) * scale
    if float(ws.abs().max()) >= 60000.0:
        raise ValueError("weight too large for the fp16 hi/lo split")
    hi = ws.to(torch.float16)
    lo = (ws - hi.to(wide)).to(torch.float16)
    return hi, lo


#: the fp8 cross terms of the gate conv (csrc/umma_gate.cu): `lo` halves are scaled by 2^11 before the e4m3 cast (they are
#: ~2^-11 of the `hi` halves), the cross-term accumulator is scaled back in the epilogue (csrc: G8_LO_SCALE_INV)
F8_LO_SCALE = 2048.0


def split_f8(w: torch.Tensor, scale: float = TC_W_SCALE):
    """Operands of the fp8 cross terms A_hi W_lo + A_lo W_hi: (hi8, lo8) = (e4m3(hi), e4m3(lo * 2^11)) as raw bytes, where
    (hi, lo) is the fp16 pair of split_f16.  The cross terms are 2^-11 of the product, so e4m3's 2^-4 relative rounding
    leaves ~2^-15 of it — measured on the whole residual stack: mel error 1.4e-4 instead of 4e-6 with fp16 cross terms,
    3.6e-3 with none (the contract is 1e-3)."""
    ws = w.to(torch.float64) * scale
    hi = ws.to(torch.float16)
    lo = ws - hi.to(torch.float64)
    hi8 = hi.to(torch.float32).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)
    lo8 = (lo * F8_LO_SCALE).to(torch.float32).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)
    return hi8, lo8


def fused_recurrence_weights(sd: Dict[str, torch.Tensor], l: int, C: int, H: int):
    """Operands of layer l of the denoiser's y-recurrence (fp64): weights (C, C + C + H) and bias (C,).

    With y_l = x_l + Wc_l cond + bc_l + step_l + spk_l (the k=3 conv's input, model/blocks.py:669-678) and
    r = 1/sqrt(2) (:683-686):  y_{l+1} = [r Wo_l[:C] | r I | Wc_{l+1} - r Wc_l] [g_l ; y_l ; cond] + const.
    The identity block lets y_l flow through the GEMM's operand pipeline instead of being read by the epilogue."""
    r = 1.0 / math.sqrt(2.0)
    p, pn = f"net.residual_layers.{l}.", f"net.residual_layers.{l + 1}."
    wo = sd[p + "output_projection.conv.weight"][:, :, 0].double()
    bo = sd[p + "output_projection.conv.bias"].double()
    wc = sd[p + "conditioner_projection.conv.weight"][:, :, 0].double()
    bc = sd[p + "conditioner_projection.conv.bias"].double()
    wcn = sd[pn + "conditioner_projection.conv.weight"][:, :, 0].double()
    bcn = sd[pn + "conditioner_projection.conv.bias"].double()
    wf = torch.zeros(C, 2 * C + H, dtype=torch.float64)
    wf[:, :C] = r * wo[:C]
    wf[:, C:2 * C] = r * torch.eye(C, dtype=torch.float64)
    wf[:, 2 * C:] = wcn - r * wc
    bf = r * bo[:C] + bcn - r * bc
    return wf, bf


def cond_stack_weights(sd: Dict[str, torch.Tensor], n_layers: int, C: int, H: int):
    """Everything the residual stack needs from the conditioner, as ONE GEMM per batch (csrc/pipeline.cu,
    cmtts_denoiser_cond_tc): weights (n_layers * C, H) and bias (n_layers * C,)  (fp64).  Row block 0 is layer 0's
    conditioner projection Wc_0 (+ bc_0): the conditioner term of y_0; row block l > 0 is the conditioner block
    Wc_l - r Wc_{l-1} of fused_recurrence_weights(l - 1) together with that GEMM's constant bias (the layer GEMMs of
    the solver steps then add one fp32 plane and the per-utterance step vector, nothing else)."""
    p0 = "net.residual_layers.0.conditioner_projection.conv."
    ws = [sd[p0 + "weight"][:, :, 0].double()]
    b = torch.zeros(n_layers * C, dtype=torch.float64)
    b[:C] = sd[p0 + "bias"].double()
    for l in range(n_layers - 1):
        wf, bf = fused_recurrence_weights(sd, l, C, H)
        ws.append(wf[:, 2 * C:])
        b[(l + 1) * C:(l + 2) * C] = bf
    return torch.cat(ws, dim=0), b


def skip_stack_weights(sd: Dict[str, torch.Tensor], n_layers: int, C: int):
    """sum_l (Wo_l[C:] g_l + bo_l[C:]) (model/modules.py:629-634, blocks.py:683-686) as ONE GEMM over the stacked
    gate outputs: weights (n_layers * C, C) with row l*C + n = Wo_l[C + n], and the summed bias (C,)  (fp64)."""
    ws, b = [], torch.zeros(C, dtype=torch.float64)
    for l in range(n_layers):
        p = f"net.residual_layers.{l}.output_projection.conv."
        ws.append(sd[p + "weight"][C:, :, 0].double())
        b = b + sd[p + "bias"][C:].double()
    return torch.cat(ws, dim=0), b


class _Table:
    """Keeps the device tensors alive next to the ctypes pointer array that references them."""

    def __init__(self, device):
        self.device = device
        self.tensors: List[Optional[torch.Tensor]] = []

    def add(self, t: Optional[torch.Tensor]) -> int:
        if t is not None:
            if t.dtype not in (torch.float16, torch.uint8):     # uint8: raw e4m3 bytes of the fp8 cross-term operands
                t = t.to(torch.float32)
            t = t.detach().contiguous().to(self.device)
            if t.data_ptr() % 16 != 0:
                raise _lib.CmttsError("weight tensor not 16-byte aligned")
        self.tensors.append(t)
        return len(self.tensors) - 1

    def finish(self):
        self.ptrs = _lib.pointer_table(self.tensors)
        return self


class PackedAcoustic:
    """Encoder + variance adaptor + denoiser weights of one CMTotalTTS checkpoint."""

    def __init__(self, spec: ModelSpec, sd: Dict[str, torch.Tensor], device, pe_rows: int = 4096):
        self.spec = spec
        self.device = torch.device(device)
        self.pe_rows = int(pe_rows)
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
        self._check_layout(sd)
        self.sd_cpu = sd
        self._pack(sd)

    # -- layout check ------------------------------------------------------------------------
    def _check_layout(self, sd):
        s = self.spec
        need = {
            TE + "embed_tokens.weight": (s.vocab, s.hidden),
            TE + "layers.0.op.self_attn.in_proj_weight": (3 * s.hidden, s.hidden),
            TE + "layers.0.op.ffn.ffn_1.weight": (4 * s.hidden, s.hidden, s.ffn_kernel),
            VA + "energy_bins": (s.energy_bins - 1,),
            VA + "cwt_predictor.1.linear.weight": (s.cwt_out, s.filter_size),
            VA + "pitch_embed.weight": (s.pitch_bins, s.hidden),
            "net.input_projection.0.conv.weight": (s.res_channels, s.n_mels, 1),
            f"net.residual_layers.{s.res_layers - 1}.conv_layer.conv.weight": (2 * s.res_channels, s.res_channels, 3),
            "net.output_projection.conv.weight": (s.n_mels, s.res_channels, 1),
        }
        if s.multi_speaker:
            need[SPK + "weight"] = (s.n_speakers, s.hidden) if s.n_speakers else (s.hidden, s.ext_speaker_dim)
            need["net.residual_layers.0.speaker_projection.linear.weight"] = (s.res_channels, s.hidden)
        for k, shp in need.items():
            if k not in sd:
                raise KeyError(f"checkpoint is missing {k!r} (expected the CMTotalTTS state_dict layout)")
            if tuple(sd[k].shape) != tuple(shp):
                raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != expected {shp} for spec {s.name}")

    def dims(self) -> _lib.Dims:
        s = self.spec
        return _lib.Dims(
            hidden=s.hidden, enc_layers=s.enc_layers, enc_heads=s.enc_heads, ffn_kernel=s.ffn_kernel,
            ffn_act=ACT_CODES[s.ffn_act], filter=s.filter_size, dur_layers=s.dur_layers,
            dur_kernel=s.dur_kernel, pred_layers=s.pred_layers, pred_kernel=s.pred_kernel,
            cwt_hidden=s.cwt_hidden, cwt_out=s.cwt_out, use_uv=int(s.use_uv), energy_bins=s.energy_bins,
            pitch_bins=s.pitch_bins, n_mels=s.n_mels, res_layers=s.res_layers, res_channels=s.res_channels,
            multi_speaker=int(s.multi_speaker), spk_dim=s.ext_speaker_dim, pe_rows=self.pe_rows,
            cwt_std_scale=s.cwt_std_scale, pitch_eps=s.pitch_norm_eps,
            f0_mel_min=float(np.float32(F0_MEL_MIN)), f0_mel_span=float(np.float32(F0_MEL_MAX - F0_MEL_MIN)),
        )

    # -- packing -----------------------------------------------------------------------------
    def _pack(self, sd):
        s, dev = self.spec, self.device
        H, C = s.hidden, s.res_channels
        pe_c = sinusoid_table(self.pe_rows, H)
        pe_h = sinusoid_table(self.pe_rows, s.cwt_hidden)

        enc = _Table(dev)
        enc.add(sd[TE + "embed_tokens.weight"])
        enc.add(pe_c)
        for l in range(s.enc_layers):
            p = f"{TE}layers.{l}.op."
            enc.add(sd[p + "layer_norm1.weight"]); enc.add(sd[p + "layer_norm1.bias"])
            enc.add(lin_w(sd[p + "self_attn.in_proj_weight"]))
            enc.add(lin_w(sd[p + "self_attn.out_proj.weight"]))
            enc.add(sd[p + "layer_norm2.weight"]); enc.add(sd[p + "layer_norm2.bias"])
            enc.add(conv_w(sd[p + "ffn.ffn_1.weight"])); enc.add(sd[p + "ffn.ffn_1.bias"])
            enc.add(lin_w(sd[p + "ffn.ffn_2.weight"])); enc.add(sd[p + "ffn.ffn_2.bias"])
        enc.add(sd[TE + "layer_norm.weight"]); enc.add(sd[TE + "layer_norm.bias"])
        self.enc = enc.finish()

        va = _Table(dev)
        if s.multi_speaker:
            if s.n_speakers:        # nn.Embedding table (n_speaker, H) == the [in][out] weight of a Linear over one-hot rows
                tab = torch.zeros(s.ext_speaker_dim, s.hidden, dtype=torch.float32)
                tab[: s.n_speakers] = sd[SPK + "weight"].detach().to(torch.float32).cpu()
                va.add(tab); va.add(torch.zeros(s.hidden, dtype=torch.float32))
            else:
                va.add(lin_w(sd[SPK + "weight"])); va.add(sd[SPK + "bias"])
        else:
            va.add(None); va.add(None)

        def predictor(prefix, n_layers, with_alpha):
            if with_alpha:
                va.add(sd[prefix + "pos_embed_alpha"].reshape(1).repeat(4))
            for i in range(n_layers):
                va.add(conv_w(sd[f"{prefix}conv.{i}.1.weight"])); va.add(sd[f"{prefix}conv.{i}.1.bias"])
                va.add(sd[f"{prefix}conv.{i}.3.weight"]); va.add(sd[f"{prefix}conv.{i}.3.bias"])
            va.add(sd[prefix + "linear.weight"]); va.add(sd[prefix + "linear.bias"])

        predictor(VA + "duration_predictor.", s.dur_layers, False)
        predictor(VA + "energy_predictor.", s.pred_layers, True)
        va.add(pe_c)
        va.add(sd[VA + "energy_bins"])
        va.add(sd[VA + "energy_embedding.weight"])
        va.add(lin_w(sd[VA + "cwt_stats_layers.0.weight"])); va.add(sd[VA + "cwt_stats_layers.0.bias"])
        va.add(lin_w(sd[VA + "cwt_stats_layers.2.weight"])); va.add(sd[VA + "cwt_stats_layers.2.bias"])
        w4 = torch.zeros(s.cwt_hidden, 4); w4[:, :2] = sd[VA + "cwt_stats_layers.4.weight"].t()
        b4 = torch.zeros(4); b4[:2] = sd[VA + "cwt_stats_layers.4.bias"]
        va.add(w4); va.add(b4)
        va.add(lin_w(sd[VA + "cwt_predictor.0.weight"])); va.add(sd[VA + "cwt_predictor.0.bias"])
        va.add(pe_h)
        predictor(VA + "cwt_predictor.1.", s.pred_layers, True)
        # inverse_cwt_torch, utils/pitch_tools.py:246: (arange(10) + 1 + 2.5) ** -2.5 in fp32
        cwt_b = (torch.arange(0, 10).float() + 1 + 2.5) ** (-2.5)
        va.add(torch.cat([cwt_b, torch.zeros(2)]))
        va.add(sd[VA + "pitch_embed.weight"])
        self.va = va.finish()

        dn = _Table(dev)
        dn.add(conv_w(sd["net.input_projection.0.conv.weight"])); dn.add(sd["net.input_projection.0.conv.bias"])
        half = C // 2
        # DiffusionEmbedding, model/blocks.py:635-637 (arange int64 * python float -> fp32, exp fp32)
        freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
        dn.add(freq.to(torch.float32))
        dn.add(lin_w(sd["net.mlp.0.linear.weight"])); dn.add(lin_w(sd["net.mlp.2.linear.weight"]))
        dn.add(torch.cat([lin_w(sd[f"net.residual_layers.{l}.diffusion_projection.linear.weight"])
                          for l in range(s.res_layers)], dim=1))
        if s.multi_speaker:
            dn.add(torch.cat([lin_w(sd[f"net.residual_layers.{l}.speaker_projection.linear.weight"])
                              for l in range(s.res_layers)], dim=1))
        else:
            dn.add(None)
        perm = gate_permutation(C)
        for l in range(s.res_layers):
            p = f"net.residual_layers.{l}."
            dn.add(conv_w(sd[p + "conditioner_projection.conv.weight"])); dn.add(sd[p + "conditioner_projection.conv.bias"])
            dn.add(conv_w(sd[p + "conv_layer.conv.weight"])[:, :, perm]); dn.add(sd[p + "conv_layer.conv.bias"][perm])
            wo, bo = sd[p + "output_projection.conv.weight"], sd[p + "output_projection.conv.bias"]
            dn.add(conv_w(wo[:C])); dn.add(bo[:C].clone())
            dn.add(conv_w(wo[C:])); dn.add(bo[C:].clone())
        dn.add(conv_w(sd["net.skip_projection.conv.weight"])); dn.add(sd["net.skip_projection.conv.bias"])
        dn.add(conv_w(sd["net.output_projection.conv.weight"])); dn.add(sd["net.output_projection.conv.bias"])
        self.dn = dn.finish()

        # tensor-core operands of the residual stack: fp16 hi/lo pairs, [tap][Cout][Cin] (K-major)
        dn16 = _Table(dev)
        for l in range(s.res_layers):
            p = f"net.residual_layers.{l}."
            for w in (conv_w_nk(sd[p + "conditioner_projection.conv.weight"]),
                      conv_w_nk(sd[p + "conv_layer.conv.weight"][perm]),
                      conv_w_nk(sd[p + "output_projection.conv.weight"])):
                hi, lo = split_f16(w)
                dn16.add(hi); dn16.add(lo)
            dn16.add(sd[p + "output_projection.conv.bias"])
        # projections around the stack: input (K = n_mels zero-padded to 128) and skip
        w_in = torch.zeros(C, 128)
        w_in[:, :s.n_mels] = sd["net.input_projection.0.conv.weight"][:, :, 0]
        for w in (w_in, sd["net.skip_projection.conv.weight"][:, :, 0].contiguous()):
            hi, lo = split_f16(w)
            dn16.add(hi); dn16.add(lo)
        # y-recurrence of the residual stack (csrc/pipeline.cu, cmtts_denoiser_forward_tc): per layer l < last the
        # (C, 2C + H) operand of fused_recurrence_weights, then the stacked skip projection of skip_stack_weights;
        # derived in fp64, then split into fp16 hi/lo like every other operand
        # (the GEMM of the solver steps contracts over [g_l ; y_l] only: the conditioner block goes to the per-batch
        # stack at the end of this table)
        for l in range(s.res_layers - 1):
            wf, bf = fused_recurrence_weights(sd, l, C, s.hidden)
            hi, lo = split_f16(wf[:, :2 * C].contiguous())
            dn16.add(hi); dn16.add(lo); dn16.add(bf.to(torch.float32))
        wsk, bsk = skip_stack_weights(sd, s.res_layers, C)
        hi, lo = split_f16(wsk)
        dn16.add(hi); dn16.add(lo); dn16.add(bsk.to(torch.float32))
        # output projection (C -> n_mels) on the hi/lo kernel: rows zero-padded to one 128-wide N tile
        w_out = torch.zeros(128, C)
        w_out[:s.n_mels] = sd["net.output_projection.conv.weight"][:, :, 0]
        b_out = torch.zeros(128)
        b_out[:s.n_mels] = sd["net.output_projection.conv.bias"]
        hi, lo = split_f16(w_out)
        dn16.add(hi); dn16.add(lo); dn16.add(b_out)
        # fp8 copies of the gate conv's weights for its cross terms: per layer {hi8, lo8 [tap][2C][C] bytes}
        for l in range(s.res_layers):
            hi8, lo8 = split_f8(conv_w_nk(sd[f"net.residual_layers.{l}.conv_layer.conv.weight"][perm]))
            dn16.add(hi8); dn16.add(lo8)
        # conditioner projections of all layers as one per-batch GEMM (cond_stack_weights)
        wcs, bcs = cond_stack_weights(sd, s.res_layers, C, s.hidden)
        hi, lo = split_f16(wcs)
        dn16.add(hi); dn16.add(lo); dn16.add(bcs.to(torch.float32))
        self.dn16 = dn16.finish()

        def add_pair(tab, w):
            hi, lo = split_f16(w)
            tab.add(hi); tab.add(lo)

        # encoder GEMMs on the hi/lo tensor-core kernel: per layer in_proj, out_proj, ffn1, ffn2
        enc16 = _Table(dev)
        for l in range(s.enc_layers):
            p = f"{TE}layers.{l}.op."
            add_pair(enc16, sd[p + "self_attn.in_proj_weight"].contiguous())              # [3C][C]
            add_pair(enc16, sd[p + "self_attn.out_proj.weight"].contiguous())             # [C][C]
            add_pair(enc16, conv_w_nk(sd[p + "ffn.ffn_1.weight"]))                        # [k][4C][C]
            add_pair(enc16, sd[p + "ffn.ffn_2.weight"].contiguous())                      # [C][4C]
        self.enc16 = enc16.finish()

        # variance-adaptor convs: duration, energy, cwt_in, cwt predictor
        va16 = _Table(dev)
        for i in range(s.dur_layers):
            add_pair(va16, conv_w_nk(sd[f"{VA}duration_predictor.conv.{i}.1.weight"]))
        for i in range(s.pred_layers):
            add_pair(va16, conv_w_nk(sd[f"{VA}energy_predictor.conv.{i}.1.weight"]))
        add_pair(va16, sd[VA + "cwt_predictor.0.weight"].contiguous())                    # [h][C]
        for i in range(s.pred_layers):
            add_pair(va16, conv_w_nk(sd[f"{VA}cwt_predictor.1.conv.{i}.1.weight"]))
        self.va16 = va16.finish()

    def ensure_pe_rows(self, n: int) -> None:
        """Sinusoid tables auto-grow like the reference's (model/blocks.py:65-72)."""
        if n + 2 > self.pe_rows:
            self.pe_rows = int(2 ** math.ceil(math.log2(n + 2)))
            self._pack(self.sd_cpu)


class PackedHifiGan:
    """HiFi-GAN generator weights with weight-norm folded (hifigan/models.py:167-174)."""

    def __init__(self, hspec: HifiGanSpec, generator_sd: Dict[str, torch.Tensor], device):
        self.hspec = hspec
        self.device = torch.device(device)
        sd = fold_weight_norm({k: v.detach().to("cpu", torch.float32) for k, v in generator_sd.items()})
        t = _Table(self.device)
        C0 = hspec.upsample_initial_channel
        if tuple(sd["conv_pre.weight"].shape) != (C0, hspec.n_mels, 7):
            raise ValueError("conv_pre.weight shape does not match the HiFi-GAN config")
        t.add(conv_w(sd["conv_pre.weight"])); t.add(sd["conv_pre.bias"])
        taps, shift0 = [], []
        nk = len(hspec.resblock_kernel_sizes)
        nd = len(hspec.resblock_dilation_sizes[0])
        # table order = consumption order in cmtts_hifigan_forward: per level {up, its MRF resblocks}
        for i, (u, k) in enumerate(zip(hspec.upsample_rates, hspec.upsample_kernel_sizes)):
            w, b, d0 = pack_conv_transpose(sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], u, (k - u) // 2)
            t.add(w); t.add(b)
            taps.append(w.shape[0]); shift0.append(d0)
            for j in range(nk):
                r = i * nk + j
                for m in range(nd):
                    t.add(conv_w(sd[f"resblocks.{r}.convs1.{m}.weight"])); t.add(sd[f"resblocks.{r}.convs1.{m}.bias"])
                    t.add(conv_w(sd[f"resblocks.{r}.convs2.{m}.weight"])); t.add(sd[f"resblocks.{r}.convs2.{m}.bias"])
        wp = sd["conv_post.weight"]            # (1, C, k)
        t.add(wp[0].t().contiguous())          # [k][C]
        t.add(sd["conv_post.bias"].reshape(1).repeat(4))
        self.table = t.finish()
        # tensor-core table: fp16 K-major weights, fp32 biases; conv_pre / conv_post stay fp32
        t16 = _Table(self.device)
        t16.add(conv_w(sd["conv_pre.weight"])); t16.add(sd["conv_pre.bias"])
        for i, (u, k) in enumerate(zip(hspec.upsample_rates, hspec.upsample_kernel_sizes)):
            w, b, d0 = pack_conv_transpose(sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], u, (k - u) // 2)
            t16.add(w.permute(0, 2, 1).contiguous().to(torch.float16)); t16.add(b)   # [taps][u*Cout][Cin]
            for j in range(nk):
                r = i * nk + j
                for m in range(nd):
                    t16.add(conv_w_nk(sd[f"resblocks.{r}.convs1.{m}.weight"]).to(torch.float16)); t16.add(sd[f"resblocks.{r}.convs1.{m}.bias"])
                    t16.add(conv_w_nk(sd[f"resblocks.{r}.convs2.{m}.weight"]).to(torch.float16)); t16.add(sd[f"resblocks.{r}.convs2.{m}.bias"])
        t16.add(wp[0].t().contiguous()); t16.add(sd["conv_post.bias"].reshape(1).repeat(4))
        # conv_pre on the hi/lo tensor-core kernel: [k][C0][128] (mel channels zero-padded to one 128-wide K block)
        w_pre = torch.zeros(sd["conv_pre.weight"].shape[2], C0, 128)
        w_pre[:, :, :hspec.n_mels] = conv_w_nk(sd["conv_pre.weight"])
        hi, lo = split_f16(w_pre)
        t16.add(hi); t16.add(lo)
        # the 32-channel level once more, pair-packed for the two-time-steps-per-row ResBlock kernel: per (resblock,
        # iteration) in consumption order {w1p, w2p} (csrc/pipeline.cu: cmtts_hifigan_forward_tc)
        ch = C0
        for i in range(len(hspec.upsample_rates)):
            ch //= 2
            if ch != 32:
                continue
            for j in range(nk):
                r = i * nk + j
                for m in range(nd):
                    w1 = conv_w_nk(sd[f"resblocks.{r}.convs1.{m}.weight"]).to(torch.float16)
                    w2 = conv_w_nk(sd[f"resblocks.{r}.convs2.{m}.weight"]).to(torch.float16)
                    k, d = w1.shape[0], hspec.resblock_dilation_sizes[j][m]
                    if k % 2 == 1 and d % 2 == 1:
                        t16.add(pair_pack_d1(w1) if d == 1 else pair_pack_taps(w1)); t16.add(pair_pack_d1(w2))
                    else:
                        t16.add(None); t16.add(None)
        self.table16 = t16.finish()
        dil = [d for ds in hspec.resblock_dilation_sizes for d in ds]
        # per level: may a tile skip its all-zero tap?  True when the packed transposed conv has 3 taps of which the
        # first only feeds output phases [0, u/2) and the last only phases [u/2, u) (k = 2 u, padding u/2) — checked on
        # the actual packed weights, not assumed
        split_ok = []
        for i, (u, k) in enumerate(zip(hspec.upsample_rates, hspec.upsample_kernel_sizes)):
            w, _, d0 = pack_conv_transpose(sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], u, (k - u) // 2)   # [taps][Cin][u*Cout]
            half = w.shape[2] // 2
            ok = w.shape[0] == 3 and d0 == -1 and u % 2 == 0 and float(w[0][:, half:].abs().max()) == 0.0 \
                and float(w[2][:, :half].abs().max()) == 0.0
            split_ok.append(1 if ok else 0)
        cfg = [len(hspec.upsample_rates), C0, nk, nd, 7, wp.shape[2]] + list(hspec.upsample_rates) + taps + shift0 \
            + list(hspec.resblock_kernel_sizes) + dil + split_ok
        self.cfg = (C.c_int32 * len(cfg))(*cfg)
        self.hop = hspec.hop
