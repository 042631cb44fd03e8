"""ctypes binding of libcmtts_b200.so (the C ABI declared in include/cmtts_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcmtts_b200.so")

i64 = C.c_int64
f32 = C.c_float
vp = C.c_void_p
szt = C.c_size_t


class CmttsError(RuntimeError):
    pass


class Dims(C.Structure):
    """struct cmtts_dims (include/cmtts_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "hidden", "enc_layers", "enc_heads", "ffn_kernel", "ffn_act", "filter", "dur_layers",
        "dur_kernel", "pred_layers", "pred_kernel", "cwt_hidden", "cwt_out", "use_uv", "energy_bins",
        "pitch_bins", "n_mels", "res_layers", "res_channels", "multi_speaker", "spk_dim", "pe_rows")
    ] + [(n, C.c_float) for n in ("cwt_std_scale", "pitch_eps", "f0_mel_min", "f0_mel_span")]


class ConvDesc(C.Structure):
    """struct cmtts_conv_desc."""
    _fields_ = [
        ("B", C.c_int32), ("M", C.c_int32), ("Lin", C.c_int32), ("Cin", C.c_int32), ("N", C.c_int32),
        ("taps", C.c_int32), ("shift", C.c_int32 * 16),
        ("x_ld", C.c_int32), ("out_ld", C.c_int32), ("res_ld", C.c_int32),
        ("x_bstride", i64), ("out_bstride", i64), ("res_bstride", i64), ("addvec_bstride", i64),
        ("pre_lrelu", C.c_int32), ("pre_slope", f32),
        ("alpha", f32), ("beta", f32), ("act", C.c_int32), ("act_slope", f32),
        ("res_scale", f32), ("out_scale", f32), ("accumulate", C.c_int32),
    ]


class UmmaDesc(C.Structure):
    """struct cmtts_umma_desc."""
    _fields_ = [
        ("B", C.c_int32), ("M", C.c_int32), ("Lin", C.c_int32), ("N", C.c_int32), ("Cin", C.c_int32),
        ("taps", C.c_int32), ("shift", C.c_int32 * 16), ("split", C.c_int32), ("epi", C.c_int32),
        ("a_ld", C.c_int32), ("res_ld", C.c_int32), ("out_ld", C.c_int32), ("x_ld", C.c_int32),
        ("a_bstride", i64), ("res_bstride", i64), ("out_bstride", i64), ("x_bstride", i64), ("addvec_bstride", i64),
        ("alpha", f32), ("res_inv_slope", f32), ("out_slope", f32), ("out_scale", f32),
        ("skip_accumulate", C.c_int32),
    ]


PD = C.POINTER(Dims)
PV = C.POINTER(vp)
PI32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); mirrors include/cmtts_b200.h one to one
PROTOTYPES = {
    "cmtts_abi_version": (C.c_int, []),
    "cmtts_last_error": (C.c_char_p, []),
    "cmtts_launch_count": (C.c_uint64, []),
    "cmtts_debug_set": (None, [C.c_int32, C.c_int32]),
    "cmtts_prof_begin": (C.c_int, [vp]),
    "cmtts_prof_end": (C.c_int64, [C.c_char_p, szt]),
    "cmtts_encoder_workspace_bytes": (szt, [PD, i64, i64]),
    "cmtts_encoder_forward": (C.c_int, [PD, PV, vp, vp, i64, i64, vp, vp, szt, vp]),
    "cmtts_variance_token_workspace_bytes": (szt, [PD, i64, i64]),
    "cmtts_variance_token": (C.c_int, [PD, PV, vp, vp, vp, f32, f32, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                       vp, szt, vp]),
    "cmtts_variance_frame_workspace_bytes": (szt, [PD, i64, i64]),
    "cmtts_variance_frame": (C.c_int, [PD, PV, vp, vp, vp, vp, f32, i64, i64, i64, vp, vp, vp, vp, vp, vp, szt, vp]),
    "cmtts_denoiser_prepare_workspace_bytes": (szt, [PD, i64]),
    "cmtts_denoiser_prepare": (C.c_int, [PD, PV, vp, vp, i64, vp, vp, vp, szt, vp]),
    "cmtts_denoiser_workspace_bytes": (szt, [PD, i64, i64]),
    "cmtts_denoiser_forward": (C.c_int, [PD, PV, vp, vp, vp, vp, f32, f32, f32, i64, i64, vp, vp, vp, szt, vp]),
    "cmtts_renoise": (C.c_int, [vp, vp, f32, f32, vp, i64, vp]),
    "cmtts_hifigan_workspace_bytes": (szt, [PI32, i64, i64]),
    "cmtts_hifigan_forward": (C.c_int, [PI32, PV, vp, i64, i64, vp, vp, f32, vp, szt, vp]),
    "cmtts_transpose_bcl_blc": (C.c_int, [vp, vp, i64, i64, i64, vp]),
    "cmtts_conv1d": (C.c_int, [C.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp, vp]),
    "cmtts_layernorm": (C.c_int, [vp, vp, vp, f32, vp, i64, i64, i64, vp, vp]),
    "cmtts_attention": (C.c_int, [vp, vp, vp, i64, i64, i64, i64, vp]),
    "cmtts_length_regulate": (C.c_int, [vp, vp, vp, vp, vp, i64, i64, i64, i64, vp]),
    "cmtts_round_durations": (C.c_int, [vp, f32, vp, vp, vp, vp, i64, i64, vp]),
    "cmtts_umma_conv1d": (C.c_int, [C.POINTER(UmmaDesc)] + [vp] * 13),
    "cmtts_f32_to_f16": (C.c_int, [vp, vp, vp, i64, i64, i64, f32, vp]),
    "cmtts_denoiser_tc_workspace_bytes": (szt, [PD, i64, i64]),
    "cmtts_denoiser_cond_tc_bytes": (szt, [PD, i64, i64]),
    "cmtts_denoiser_cond_tc_workspace_bytes": (szt, [PD, i64, i64]),
    "cmtts_denoiser_cond_tc": (C.c_int, [PD, PV, PV, vp, i64, i64, vp, vp, szt, vp]),
    "cmtts_denoiser_forward_tc": (C.c_int, [PD, PV, PV, vp, vp, vp, vp, f32, f32, f32, i64, i64, vp, vp, vp, szt, vp]),
    "cmtts_encoder_tc_workspace_bytes": (szt, [PD, i64, i64]),
    "cmtts_encoder_forward_tc": (C.c_int, [PD, PV, PV, vp, vp, i64, i64, vp, vp, szt, vp]),
    "cmtts_variance_token_tc_workspace_bytes": (szt, [PD, i64, i64]),
    "cmtts_variance_token_tc": (C.c_int, [PD, PV, PV, vp, vp, vp, f32, f32, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                          vp, szt, vp]),
    "cmtts_variance_frame_tc_workspace_bytes": (szt, [PD, i64, i64]),
    "cmtts_variance_frame_tc": (C.c_int, [PD, PV, PV, vp, vp, vp, vp, f32, i64, i64, i64, vp, vp, vp, vp, vp, vp, szt, vp]),
    "cmtts_hifigan_tc_workspace_bytes": (szt, [PI32, i64, i64]),
    "cmtts_hifigan_forward_tc": (C.c_int, [PI32, PV, vp, i64, i64, vp, vp, f32, vp, szt, vp]),
    "cmtts_rescnn_workspace_bytes": (szt, [PI32, i64, i64]),
    "cmtts_rescnn_forward": (C.c_int, [PI32, PV, vp, i64, i64, vp, vp, szt, vp]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the library (no GPU needed to load / resolve symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CmttsError(
            f"{LIB_PATH} not found: build it with `python -m cmtts_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.cmtts_abi_version() != 2:
        raise CmttsError("libcmtts_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cmtts_last_error().decode("utf-8", "replace")
        raise CmttsError(f"{what} failed (rc={rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CmttsError("cmtts_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise CmttsError("cmtts_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def pointer_table(tensors) -> "C.Array":
    arr = (vp * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = ptr(t) if t is not None else None
    return arr
