"""`torch.library` custom ops over the C ABI (north_star: "kernels bound as custom ops through a thin C-ABI";
SURVEY.md §7.2).

    torch.ops.cmtts_b200.encoder_forward(handle, texts, src_lens)            -> enc (B, T, H)
    torch.ops.cmtts_b200.variance_token(handle, enc, src_lens, spk?, e, d)   -> out1, log_d, d_rounded, e_pred, e_idx,
                                                                                 cumsum, mel_lens, speaker_emb, f0_stats
    torch.ops.cmtts_b200.variance_frame(handle, out1, cumsum, mel_lens, f0_stats, p, L)
                                                                              -> cond, mel2ph, cwt, f0_denorm, pitch_idx
    torch.ops.cmtts_b200.denoiser_prepare(handle, t, spk?)                    -> ds_all, dsp_all
    torch.ops.cmtts_b200.split_f16(x)                                         -> hi, lo
    torch.ops.cmtts_b200.denoiser_cond(handle, cond)                          -> cond_proj (per-batch conditioner projections)
    torch.ops.cmtts_b200.denoiser_forward(handle, x, cond, cond_proj?, ds, dsp, c_in, c_out, c_skip, want_F)
                                                                              -> out, model_out
    torch.ops.cmtts_b200.renoise(x0, noise, s1, s2)                           -> x
    torch.ops.cmtts_b200.hifigan_forward(handle, mel_blc, want_float, want_int16, max_wav) -> wav, wav_i16
    torch.ops.cmtts_b200.transpose_bcl_blc(x)                                 -> (B, L, C)
    torch.ops.cmtts_b200.rescnn_forward(handle, fbank)                        -> (B, 512) speaker embeddings (zero-shot path)

Each op is registered for the CUDA dispatch key only and is a thin shim: allocate the outputs with torch (device
memory is PyTorch's job), hand raw pointers + the current stream to libcmtts_b200.so through ctypes, return.  There
is no CPU registration: calling an op with CPU tensors fails in the dispatcher ("could not run ... 'CPU' backend"),
and a missing library fails at import of `_lib` — no fallback of any kind.  `handle` is an integer naming a loaded
model / vocoder (their packed weight tables live on the device; `register` / `release` below).
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

_DEF = torch.library.Library("cmtts_b200", "DEF")
_OWNERS: Dict[int, "weakref.ref"] = {}
_next = [1]


def register(owner) -> int:
    """Give a loaded CMTotalTTS / Generator an integer handle the ops can take as a plain schema type."""
    h = _next[0]
    _next[0] += 1
    _OWNERS[h] = weakref.ref(owner)
    return h


def release(handle: int) -> None:
    _OWNERS.pop(int(handle), None)


def _owner(handle: int):
    ref = _OWNERS.get(int(handle))
    o = ref() if ref is not None else None
    if o is None or o.packed is None:
        raise _lib.CmttsError(f"cmtts_b200 op: handle {handle} names no loaded model (load_state_dict + .to('cuda') first)")
    return o


def _ctx(o):
    return o.lib, o.spec, o.device, _lib.stream_ptr(o.device)


# ---- schemas -----------------------------------------------------------------------------------------------------
_DEF.define("encoder_forward(int handle, Tensor texts, Tensor src_lens) -> Tensor")
_DEF.define("variance_token(int handle, Tensor enc, Tensor src_lens, Tensor? spker_embeds, float e_control, float d_control) -> Tensor[]")
_DEF.define("variance_frame(int handle, Tensor out1, Tensor cumsum, Tensor mel_lens, Tensor f0_stats, float p_control, int L) -> Tensor[]")
_DEF.define("denoiser_prepare(int handle, Tensor t, Tensor? speaker_emb) -> (Tensor, Tensor)")
_DEF.define("split_f16(Tensor x) -> (Tensor, Tensor)")
_DEF.define("denoiser_cond(int handle, Tensor cond) -> Tensor")
_DEF.define("denoiser_forward(int handle, Tensor x, Tensor cond, Tensor? cond_proj, Tensor ds_all, Tensor dsp_all, "
            "float c_in, float c_out, float c_skip, bool want_model_out) -> (Tensor, Tensor)")
_DEF.define("renoise(Tensor x0, Tensor noise, float s1, float s2) -> Tensor")
_DEF.define("hifigan_forward(int handle, Tensor mel, bool want_float, bool want_int16, float max_wav_value) -> (Tensor, Tensor)")
_DEF.define("transpose_bcl_blc(Tensor x) -> Tensor")
_DEF.define("rescnn_forward(int handle, Tensor fbank) -> Tensor")


# ---- CUDA implementations ------------------------------------------------------------------------------------------
def _encoder_forward(handle: int, texts: torch.Tensor, src_lens: torch.Tensor) -> torch.Tensor:
    o = _owner(handle)
    lib, s, dev, st = _ctx(o)
    B, T = texts.shape
    o.packed_check_rows(T)
    d = C.byref(o._dims)
    enc = torch.empty(B, T, s.hidden, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        if o.precision == "tc" and o.tc_frontend:
            ws = o._ws.get("enc", lib.cmtts_encoder_tc_workspace_bytes(d, B, T))
            _lib.check(lib.cmtts_encoder_forward_tc(d, o.packed.enc.ptrs, o.packed.enc16.ptrs, _lib.ptr(texts), _lib.ptr(src_lens),
                                                    B, T, _lib.ptr(enc), _lib.ptr(ws), ws.numel(), st), "encoder_forward_tc")
        else:
            ws = o._ws.get("enc", lib.cmtts_encoder_workspace_bytes(d, B, T))
            _lib.check(lib.cmtts_encoder_forward(d, o.packed.enc.ptrs, _lib.ptr(texts), _lib.ptr(src_lens), B, T, _lib.ptr(enc),
                                                 _lib.ptr(ws), ws.numel(), st), "encoder_forward")
    return enc


def _variance_token(handle: int, enc: torch.Tensor, src_lens: torch.Tensor, spker_embeds: Optional[torch.Tensor],
                    e_control: float, d_control: float) -> List[torch.Tensor]:
    o = _owner(handle)
    lib, s, dev, st = _ctx(o)
    B, T, H = enc.shape
    d = C.byref(o._dims)
    f32 = dict(dtype=torch.float32, device=dev)
    i64 = dict(dtype=torch.int64, device=dev)
    out1 = torch.empty(B, T, H, **f32)
    log_d = torch.empty(B, T, **f32)
    d_rounded = torch.empty(B, T, **f32)
    e_pred = torch.empty(B, T, **f32)
    e_idx = torch.empty(B, T, **i64)
    cumsum = torch.empty(B, 2, T, **i64)
    mel_lens = torch.empty(B, **i64)
    spk = torch.empty(B, H, **f32) if s.multi_speaker else torch.empty(0, **f32)
    f0_stats = torch.empty(B, 4, **f32)
    args = (_lib.ptr(enc), _lib.ptr(src_lens), _lib.ptr(spker_embeds) if s.multi_speaker else None, e_control, d_control, B, T,
            _lib.ptr(out1), _lib.ptr(log_d), _lib.ptr(d_rounded), _lib.ptr(e_pred), _lib.ptr(e_idx), _lib.ptr(cumsum),
            _lib.ptr(mel_lens), _lib.ptr(spk) if s.multi_speaker else None, _lib.ptr(f0_stats))
    with torch.cuda.device(dev):
        if o.precision == "tc" and o.tc_frontend:
            ws = o._ws.get("vat", lib.cmtts_variance_token_tc_workspace_bytes(d, B, T))
            _lib.check(lib.cmtts_variance_token_tc(d, o.packed.va.ptrs, o.packed.va16.ptrs, *args, _lib.ptr(ws), ws.numel(), st),
                       "variance_token_tc")
        else:
            ws = o._ws.get("vat", lib.cmtts_variance_token_workspace_bytes(d, B, T))
            _lib.check(lib.cmtts_variance_token(d, o.packed.va.ptrs, *args, _lib.ptr(ws), ws.numel(), st), "variance_token")
    return [out1, log_d, d_rounded, e_pred, e_idx, cumsum, mel_lens, spk, f0_stats]


def _variance_frame(handle: int, out1: torch.Tensor, cumsum: torch.Tensor, mel_lens: torch.Tensor, f0_stats: torch.Tensor,
                    p_control: float, L: int) -> List[torch.Tensor]:
    o = _owner(handle)
    lib, s, dev, st = _ctx(o)
    B, T, H = out1.shape
    o.packed_check_rows(L)
    d = C.byref(o._dims)
    f32 = dict(dtype=torch.float32, device=dev)
    cond = torch.empty(B, L, H, **f32)
    mel2ph = torch.empty(B, L, dtype=torch.int64, device=dev)
    cwt = torch.empty(B, L, s.cwt_out, **f32)
    f0_denorm = torch.empty(B, L, **f32)
    pitch_idx = torch.empty(B, L, dtype=torch.int64, device=dev)
    if L > 0 and B > 0:
        args = (_lib.ptr(out1), _lib.ptr(cumsum), _lib.ptr(mel_lens), _lib.ptr(f0_stats), p_control, B, T, L, _lib.ptr(cond),
                _lib.ptr(mel2ph), _lib.ptr(cwt), _lib.ptr(f0_denorm), _lib.ptr(pitch_idx))
        with torch.cuda.device(dev):
            if o.precision == "tc" and o.tc_frontend:
                ws = o._ws.get("vaf", lib.cmtts_variance_frame_tc_workspace_bytes(d, B, L))
                _lib.check(lib.cmtts_variance_frame_tc(d, o.packed.va.ptrs, o.packed.va16.ptrs, *args, _lib.ptr(ws), ws.numel(), st),
                           "variance_frame_tc")
            else:
                ws = o._ws.get("vaf", lib.cmtts_variance_frame_workspace_bytes(d, B, L))
                _lib.check(lib.cmtts_variance_frame(d, o.packed.va.ptrs, *args, _lib.ptr(ws), ws.numel(), st), "variance_frame")
    return [cond, mel2ph, cwt, f0_denorm, pitch_idx]


def _denoiser_prepare(handle: int, t: torch.Tensor, speaker_emb: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    o = _owner(handle)
    lib, s, dev, st = _ctx(o)
    B = t.shape[0]
    n = s.res_layers * s.res_channels
    ds_all = torch.empty(B, n, dtype=torch.float32, device=dev)
    dsp_all = torch.empty(B, n, dtype=torch.float32, device=dev)       # single speaker: the library copies ds_all into it
    d = C.byref(o._dims)
    ws = o._ws.get("dnp", lib.cmtts_denoiser_prepare_workspace_bytes(d, B))
    with torch.cuda.device(dev):
        _lib.check(lib.cmtts_denoiser_prepare(d, o.packed.dn.ptrs, _lib.ptr(t), _lib.ptr(speaker_emb) if s.multi_speaker else None,
                                              B, _lib.ptr(ds_all), _lib.ptr(dsp_all), _lib.ptr(ws), ws.numel(), st),
                   "denoiser_prepare")
    return ds_all, dsp_all


def _split_f16(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    rows = x.numel() // x.shape[-1] if x.numel() else 0
    if rows:
        with torch.cuda.device(x.device):
            _lib.check(lib.cmtts_f32_to_f16(_lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), rows, x.shape[-1], x.shape[-1], 1.0,
                                            _lib.stream_ptr(x.device)), "f32_to_f16")
    return hi, lo


def _denoiser_cond(handle: int, cond: torch.Tensor) -> torch.Tensor:
    """Conditioner projections of all residual layers, once per batch: fp32 [layers][B*(L+1)][C] on the tensor-core path
    (cmtts_denoiser_cond_tc); the fp32 path projects inside cmtts_denoiser_forward and gets an empty tensor."""
    o = _owner(handle)
    lib, s, dev, st = _ctx(o)
    B, L, H = cond.shape
    if o.precision != "tc":
        return torch.empty(0, dtype=torch.float32, device=dev)
    d = C.byref(o._dims)
    cond = cond.contiguous()
    proj = torch.empty(s.res_layers, B * (L + 1), s.res_channels, dtype=torch.float32, device=dev)
    assert proj.numel() * 4 == lib.cmtts_denoiser_cond_tc_bytes(d, B, L)
    if proj.numel():
        with torch.cuda.device(dev):
            ws = o._ws.get("dnc", lib.cmtts_denoiser_cond_tc_workspace_bytes(d, B, L))
            _lib.check(lib.cmtts_denoiser_cond_tc(d, o.packed.dn.ptrs, o.packed.dn16.ptrs, _lib.ptr(cond), B, L, _lib.ptr(proj),
                                                  _lib.ptr(ws), ws.numel(), st), "denoiser_cond_tc")
    return proj


def _denoiser_forward(handle: int, x: torch.Tensor, cond: torch.Tensor, cond_proj: Optional[torch.Tensor],
                      ds_all: torch.Tensor, dsp_all: torch.Tensor, c_in: float, c_out: float,
                      c_skip: float, want_model_out: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    o = _owner(handle)
    lib, s, dev, st = _ctx(o)
    B, L, M = x.shape
    out = torch.empty_like(x)
    mo = torch.empty_like(x) if want_model_out else torch.empty(0, dtype=x.dtype, device=dev)
    d = C.byref(o._dims)
    with torch.cuda.device(dev):
        if o.precision == "tc":
            if cond_proj is None:
                cond_proj = _denoiser_cond(handle, cond)
            if tuple(cond_proj.shape) != (s.res_layers, B * (L + 1), s.res_channels):
                raise ValueError(f"denoiser_forward: cond_proj {tuple(cond_proj.shape)} does not belong to a ({B}, {L}) batch")
            ws = o._ws.get("dn", lib.cmtts_denoiser_tc_workspace_bytes(d, B, L))
            _lib.check(lib.cmtts_denoiser_forward_tc(d, o.packed.dn.ptrs, o.packed.dn16.ptrs, _lib.ptr(x), _lib.ptr(cond_proj),
                                                     _lib.ptr(ds_all), _lib.ptr(dsp_all), c_in, c_out, c_skip, B, L,
                                                     _lib.ptr(out), _lib.ptr(mo) if want_model_out else None, _lib.ptr(ws),
                                                     ws.numel(), st), "denoiser_forward_tc")
        else:
            ws = o._ws.get("dn", lib.cmtts_denoiser_workspace_bytes(d, B, L))
            _lib.check(lib.cmtts_denoiser_forward(d, o.packed.dn.ptrs, _lib.ptr(x), _lib.ptr(cond), _lib.ptr(ds_all),
                                                  _lib.ptr(dsp_all), c_in, c_out, c_skip, B, L, _lib.ptr(out),
                                                  _lib.ptr(mo) if want_model_out else None, _lib.ptr(ws), ws.numel(), st),
                       "denoiser_forward")
    return out, mo


def _renoise(x0: torch.Tensor, noise: torch.Tensor, s1: float, s2: float) -> torch.Tensor:
    lib = _lib.load()
    x0, noise = x0.contiguous(), noise.contiguous()
    out = torch.empty_like(x0)
    with torch.cuda.device(x0.device):
        _lib.check(lib.cmtts_renoise(_lib.ptr(x0), _lib.ptr(noise), s1, s2, _lib.ptr(out), out.numel(), _lib.stream_ptr(x0.device)),
                   "renoise")
    return out


def _hifigan_forward(handle: int, mel: torch.Tensor, want_float: bool, want_int16: bool, max_wav_value: float
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    o = _owner(handle)
    lib, dev = o.lib, o.device
    B, L, M = mel.shape
    n = L * o.packed.hop
    wav = torch.empty((B, n) if want_float else (0,), dtype=torch.float32, device=dev)
    w16 = torch.empty((B, n) if want_int16 else (0,), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        if o.precision == "tc":
            ws = o._ws.get("hifi", lib.cmtts_hifigan_tc_workspace_bytes(o.packed.cfg, B, L))
            _lib.check(lib.cmtts_hifigan_forward_tc(o.packed.cfg, o.packed.table16.ptrs, _lib.ptr(mel), B, L,
                                                    _lib.ptr(wav) if want_float else None, _lib.ptr(w16) if want_int16 else None,
                                                    max_wav_value, _lib.ptr(ws), ws.numel(), st), "hifigan_forward_tc")
        else:
            ws = o._ws.get("hifi", lib.cmtts_hifigan_workspace_bytes(o.packed.cfg, B, L))
            _lib.check(lib.cmtts_hifigan_forward(o.packed.cfg, o.packed.table.ptrs, _lib.ptr(mel), B, L,
                                                 _lib.ptr(wav) if want_float else None, _lib.ptr(w16) if want_int16 else None,
                                                 max_wav_value, _lib.ptr(ws), ws.numel(), st), "hifigan_forward")
    return wav, w16


def _transpose_bcl_blc(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    x = x.contiguous()
    B, Cc, L = x.shape
    out = torch.empty(B, L, Cc, dtype=torch.float32, device=x.device)
    if out.numel():
        with torch.cuda.device(x.device):
            _lib.check(lib.cmtts_transpose_bcl_blc(_lib.ptr(x), _lib.ptr(out), B, Cc, L, _lib.stream_ptr(x.device)), "transpose")
    return out


def _rescnn_forward(handle: int, fbank: torch.Tensor) -> torch.Tensor:
    """DeepSpeaker ResCNN (speaker_encoder.DeepSpeakerModel): (B, T, 64) normalised filter-bank frames -> (B, 512)."""
    o = _owner(handle)
    lib, dev = o.lib, o.device
    B, T, _ = fbank.shape
    emb = torch.empty(B, 512, dtype=torch.float32, device=dev)
    if B:
        with torch.cuda.device(dev):
            need = lib.cmtts_rescnn_workspace_bytes(o.packed.cfg, B, T)
            if o._ws is None or o._ws.numel() < need:
                o._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            _lib.check(lib.cmtts_rescnn_forward(o.packed.cfg, o.packed.ptrs, _lib.ptr(fbank), B, T, _lib.ptr(emb), _lib.ptr(o._ws),
                                                o._ws.numel(), _lib.stream_ptr(dev)), "rescnn_forward")
    return emb


_IMPL = torch.library.Library("cmtts_b200", "IMPL", "CUDA")
for _name, _fn in (("encoder_forward", _encoder_forward), ("variance_token", _variance_token), ("variance_frame", _variance_frame),
                   ("denoiser_prepare", _denoiser_prepare), ("split_f16", _split_f16), ("denoiser_cond", _denoiser_cond),
                   ("denoiser_forward", _denoiser_forward), ("rescnn_forward", _rescnn_forward),
                   ("renoise", _renoise), ("hifigan_forward", _hifigan_forward), ("transpose_bcl_blc", _transpose_bcl_blc)):
    _IMPL.impl(_name, _fn)

OPS = ("encoder_forward", "variance_token", "variance_frame", "denoiser_prepare", "split_f16", "denoiser_cond", "denoiser_forward", "renoise",
       "hifigan_forward", "transpose_bcl_blc", "rescnn_forward")
