"""HiFi-GAN V1 generator with the reference's call protocol:

    vocoder = get_vocoder(model_config, device)            # utils/model.py:155-184
    wav = vocoder(mels[B, 80, L])                          # hifigan/models.py:149-165 -> (B, 1, 256 L)
    wavs = vocoder_infer(mels, vocoder, model_config, preprocess_config, lengths)   # utils/model.py:187-205

The generator runs in libcmtts_b200.so: transposed convolutions packed as ordinary channels-last
convs, MRF ResBlocks with leaky-ReLU applied on operand load, residual / MRF accumulation in the
GEMM epilogues, tanh + int16 conversion fused into the output stage.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import ops  # noqa: F401  (registers torch.ops.cmtts_b200.*)
from .config import HifiGanSpec
from .model import _Workspace
from .weights import PackedHifiGan


class AttrDict(dict):
    """hifigan/__init__.py AttrDict."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def hspec_from_config(h) -> HifiGanSpec:
    g = (lambda k: h[k]) if isinstance(h, dict) else (lambda k: getattr(h, k))
    return HifiGanSpec(
        n_mels=int(g("num_mels")) if (isinstance(h, dict) and "num_mels" in h) or hasattr(h, "num_mels") else 80,
        upsample_rates=tuple(g("upsample_rates")), upsample_kernel_sizes=tuple(g("upsample_kernel_sizes")),
        upsample_initial_channel=int(g("upsample_initial_channel")),
        resblock_kernel_sizes=tuple(g("resblock_kernel_sizes")),
        resblock_dilation_sizes=tuple(tuple(d) for d in g("resblock_dilation_sizes")),
    )


class Generator:
    """B200-native stand-in for hifigan.Generator (inference only)."""

    #: "tc" — fp16 operands / fp32 accumulation on tcgen05 tensor cores; "fp32" — FFMA yardstick
    PRECISIONS = ("tc", "fp32")

    def __init__(self, h=None, hspec: Optional[HifiGanSpec] = None, precision: str = "tc"):
        if precision not in self.PRECISIONS:
            raise ValueError(f"precision must be one of {self.PRECISIONS}")
        self.precision = precision
        self.hspec = hspec if hspec is not None else (hspec_from_config(h) if h is not None else HifiGanSpec())
        self.device = torch.device("cpu")
        self._sd: Optional[Dict[str, torch.Tensor]] = None
        self.packed: Optional[PackedHifiGan] = None
        self._ws: Optional[_Workspace] = None
        self.handle: Optional[int] = None        # names this vocoder in torch.ops.cmtts_b200.hifigan_forward
        self.lib = _lib.load()

    def load_state_dict(self, sd, strict: bool = True):
        self._sd = {k: v.detach().to("cpu") for k, v in sd.items()}
        if self.device.type == "cuda":
            self._repack()
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.CmttsError("cmtts_b200 runs on CUDA devices only (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        if self._sd is not None:
            self._repack()
        return self

    def eval(self):
        return self

    def remove_weight_norm(self):
        """Weight norm is folded when the checkpoint is packed (hifigan/models.py:167-174)."""
        return self

    def _repack(self):
        from . import ops
        self.packed = PackedHifiGan(self.hspec, self._sd, self.device)
        self._ws = _Workspace(self.device)
        if self.handle is None:
            self.handle = ops.register(self)

    def run(self, mel_blc: torch.Tensor, want_float: bool = True, want_int16: bool = False,
            max_wav_value: float = 32768.0):
        """(B, L, 80) channels-last mels -> wav (B, hop L) fp32 and/or int16."""
        if self.packed is None:
            raise _lib.CmttsError("Generator: call load_state_dict(...) and .to('cuda') first")
        mel = mel_blc.to(self.device, torch.float32).contiguous()
        if mel.shape[2] != self.hspec.n_mels:
            raise ValueError(f"expected {self.hspec.n_mels} mel channels, got {mel.shape[2]}")
        wav, w16 = torch.ops.cmtts_b200.hifigan_forward(self.handle, mel, bool(want_float), bool(want_int16), float(max_wav_value))
        return (wav if want_float else None), (w16 if want_int16 else None)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """hifigan.Generator.forward: (B, 80, L) -> (B, 1, 256 L)."""
        mel = torch.ops.cmtts_b200.transpose_bcl_blc(x.to(self.device, torch.float32).contiguous())
        wav, _ = self.run(mel)
        return wav.unsqueeze(1)

    __call__ = forward


def get_vocoder(config, device, checkpoint_path: Optional[str] = None, hifigan_config: Optional[str] = None):
    """utils/model.py:155-184.  Looks for `hifigan/generator_{LJSpeech,universal}.pth.tar` and
    `hifigan/config.json` relative to the cwd like the reference, unless paths are given."""
    name = config["vocoder"]["model"]
    speaker = config["vocoder"]["speaker"]
    if name != "HiFi-GAN":
        raise NotImplementedError("only the HiFi-GAN vocoder is on the hot path (MelGAN needs torch.hub)")
    cfg_path = hifigan_config or "hifigan/config.json"
    if os.path.isfile(cfg_path):
        with open(cfg_path) as f:
            h = AttrDict(json.load(f))
        voc = Generator(h)
    else:
        voc = Generator(hspec=HifiGanSpec())
    path = checkpoint_path or f"hifigan/generator_{speaker}.pth.tar"
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    voc.load_state_dict(ckpt["generator"])
    voc.eval()
    voc.remove_weight_norm()
    voc.to(device)
    return voc


def vocoder_infer(mels: torch.Tensor, vocoder: Generator, model_config, preprocess_config,
                  lengths: Optional[Sequence[int]] = None) -> List[np.ndarray]:
    """utils/model.py:187-205: (B, 80, L) mels -> list of int16 arrays cropped to `lengths`.
    The x 32768 scaling, truncating cast and crop offsets are the reference's; the conversion runs
    on the device so only int16 samples cross PCIe."""
    if model_config["vocoder"]["model"] != "HiFi-GAN":
        raise NotImplementedError("only HiFi-GAN")
    max_wav = float(preprocess_config["preprocessing"]["audio"]["max_wav_value"])
    x = mels.to(vocoder.device, torch.float32).contiguous()
    B = x.shape[0]
    mel = torch.ops.cmtts_b200.transpose_bcl_blc(x)
    _, w16 = vocoder.run(mel, want_float=False, want_int16=True, max_wav_value=max_wav)
    host = w16.cpu().numpy()
    wavs = [host[i] for i in range(B)]
    if lengths is not None:
        lens = lengths.tolist() if hasattr(lengths, "tolist") else list(lengths)
        wavs = [w[: int(n)] for w, n in zip(wavs, lens)]
    return wavs
