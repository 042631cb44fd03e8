"""Helpers shared by the GPU parity tests."""
import os

import torch

from cmtts_b200 import synthetic
from cmtts_b200.config import HifiGanSpec, ModelSpec
from oracle import cmtts_oracle as O

from conftest import GOLDEN, ROOT

DEV = "cuda:0"


def load_golden(path):
    g = torch.load(path, map_location="cpu", weights_only=True)
    m = g["meta"]
    spec = ModelSpec.preset(m["dataset"])
    sd = synthetic.make_acoustic_state_dict(spec, m["weight_seed"])
    assert synthetic.state_dict_digest(sd) == m["digest"], "synthetic weight RNG stream drifted"
    batch = {"speakers": torch.zeros(m["batch"], dtype=torch.int64), "texts": g["texts"],
             "src_lens": g["src_lens"], "spker_embeds": g["spker_embeds"]}
    return g, m, spec, sd, batch


_MODELS = {}


def gpu_model(spec, sd, key, precision="tc"):
    from cmtts_b200.model import CMTotalTTS
    key = (key, precision)
    if key not in _MODELS:
        _MODELS[key] = CMTotalTTS(spec=spec, precision=precision).load_state_dict(sd).to(DEV)
    return _MODELS[key]


# mel tolerances: north_star demands 1e-3 max-abs; the fp32 FFMA path differs from torch-CPU only by
# summation order; the tensor-core path multiplies fp16 hi/lo pairs (22-bit operands)
MEL_TOL = {"fp32": 1e-4, "tc": 3e-4}
MODEL_OUT_TOL = {"fp32": 2e-4, "tc": 6e-4}
# HiFi-GAN: fp32 path ~1e-6; tensor-core path stores fp16 activations (SNR >= 50 dB vs the reference)
WAV_TOL = {"fp32": 2e-5, "tc": 4e-3}


class Replay:
    """generator seam: replays the given CPU noise tensors in order (moved to the device)."""

    def __init__(self, tensors):
        self.it = iter(tensors)

    def randn(self, *shape, device=None, **_):
        t = next(self.it)
        assert tuple(t.shape) == tuple(shape)
        return t.to(device)

    def randn_like(self, x):
        return self.randn(*x.shape, device=x.device)


def draw_noise(seed, shape, n):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*shape, generator=g) for _ in range(n)]


def real_hifigan_weights():
    for p in ("/root/reference/hifigan/generator_universal.pth.tar",
              os.path.join(ROOT, "oracle", "_ref", "hifigan", "generator_universal.pth.tar")):
        if os.path.isfile(p):
            return p
    return None
