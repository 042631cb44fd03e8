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


def gpu_model(spec, sd, key):
    from cmtts_b200.model import CMTotalTTS
    if key not in _MODELS:
        _MODELS[key] = CMTotalTTS(spec=spec).load_state_dict(sd).to(DEV)
    return _MODELS[key]


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
