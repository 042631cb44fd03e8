"""Import shim for the UNMODIFIED reference under /root/reference (test infrastructure only).

This file is part of the ORACLE: only `tests/`, `oracle/make_golden.py`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline leg may import anything under `oracle/`.  The product package
`cmtts_b200/` never does.

The reference is plain Python/PyTorch but does not import as shipped:
  * model/cm_tool/tts_net.py:7 imports `..diffgantts` (the class lives in model/cmtts.py:10);
  * utils/tools.py:10-13, utils/pitch_tools.py:4-9 import matplotlib / sklearn / librosa /
    parselmouth / pycwt at module import, none of which the hot path calls;
  * model/cm_tool/dist_util.py imports mpi4py + blobfile, model/loss.py imports piq.
The shim registers stub modules for the absent third-party packages and the one alias; no
reference source is edited or copied.  It reads /root/reference in the build container; on
the GPU box it falls back to the copy that __graft_entry__.build() stages under oracle/_ref/ (git-ignored).
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import tempfile
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
#: staged copy made by __graft_entry__.build() (git-ignored; travels to the GPU box with the snapshot)
STAGED_ROOT = os.path.join(_HERE, "_ref")
#: sub-trees of the reference the hot path imports (+ its YAML configs): ~0.4 MB of Python + the HiFi-GAN weights
STAGED_PARTS = ("model", "utils", "text", "config", "hifigan/__init__.py", "hifigan/models.py", "hifigan/config.json",
                "hifigan/generator_universal.pth.tar",
                # DeepSpeaker ResCNN weights of the zero-shot path (97 MB Keras HDF5; read by cmtts_b200.h5lite)
                "deepspeaker/pretrained_models/ResCNN_triplet_training_checkpoint_265.h5")


def _pick_root() -> str:
    env = os.environ.get("CMTTS_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "model", "cmtts.py")):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _pick_root()


def stage_reference(src: str = "/root/reference", dst: str = STAGED_ROOT) -> bool:
    """Copy the UNMODIFIED hot-path sub-trees of the mounted reference into oracle/_ref/ (build container only), so
    that `bench.py --impl reference` and the boundary tests can run the reference's own code and read its own YAML
    configs on the GPU box, where /root/reference does not exist.  oracle/_ref/ is git-ignored: nothing of the
    reference enters the history."""
    import shutil

    if not os.path.isfile(os.path.join(src, "model", "cmtts.py")):
        return False
    for part in STAGED_PARTS:
        s, d = os.path.join(src, part), os.path.join(dst, part)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            if not (os.path.isfile(d) and os.path.getsize(d) == os.path.getsize(s)):
                shutil.copyfile(s, d)
    return True


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "cmtts.py"))


class _Anything:
    """Attribute sink: any attribute access / call returns another sink."""

    def __init__(self, name="stub"):
        self._name = name

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _Anything(self._name + "." + item)

    def __call__(self, *a, **k):
        return _Anything(self._name + "()")

    def __iter__(self):
        return iter(())


def _stub_module(name: str, **attrs) -> types.ModuleType:
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so `import a.b` works

    def _getattr(item, _n=name):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _Anything(_n + "." + item)

    m.__getattr__ = _getattr  # PEP 562
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub_module(parent), child, m)
    return m


_INSTALLED = False


def install() -> None:
    """Make `import model.cm_tool.tts_net`, `import hifigan`, ... work from /root/reference."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in (
        "matplotlib", "matplotlib.pyplot", "librosa", "parselmouth", "pycwt", "pycwt.wavelet",
        "mpi4py", "mpi4py.MPI", "blobfile", "piq", "unidecode", "inflect", "g2p_en", "tgt",
        "pyworld", "soundfile", "torchcrf",
    ):
        try:
            importlib.import_module(name)
        except Exception:
            _stub_module(name)
    # model/__init__.py:4 -> model/speaker_embedder.py:8 -> deepspeaker (TensorFlow-Keras, offline only)
    for name in ("deepspeaker", "deepspeaker.embedding", "tensorflow", "python_speech_features"):
        _stub_module(name)
    # tts_net.py:7 — `from ..diffgantts import DurationPitchSpeakerNet`
    cmtts = importlib.import_module("model.cmtts")
    sys.modules["model.diffgantts"] = cmtts
    import model as _model_pkg

    _model_pkg.diffgantts = cmtts
    # The reference picks its device once, at import: `device = cuda if available` module globals (utils/tools.py:22,
    # model/modules.py:29) that get_mask_from_lengths (:280) and LengthRegulator.LR (:434) move tensors to.  The oracle
    # and the CPU baseline run the reference on the HOST, also on a box that has a GPU: point those globals at the CPU
    # (run-time state of the imported modules; no source is touched).
    import torch

    import model.modules as _mm
    import utils.tools as _ut

    _ut.device = torch.device("cpu")
    _mm.device = torch.device("cpu")
    _INSTALLED = True


def load_configs(dataset: str):
    """The reference's three YAMLs, read the way utils/tools.py:25-33 does but cwd-independent."""
    import yaml

    cfg_dir = os.path.join(REFERENCE_ROOT, "config", dataset)
    out = []
    for f in ("preprocess.yaml", "model.yaml", "train.yaml"):
        with open(os.path.join(cfg_dir, f), "r") as fh:
            out.append(yaml.load(fh, Loader=yaml.FullLoader))
    return tuple(out)


def make_preprocessed_dir(energy_min=-1.5, energy_max=8.0, n_speakers=4) -> str:
    """stats.json / speakers.json that VarianceAdaptor.__init__ (modules.py:233-237) and
    DurationPitchSpeakerNet.__init__ (cmtts.py:27-37) read at construction."""
    d = tempfile.mkdtemp(prefix="cmtts_pre_")
    with open(os.path.join(d, "stats.json"), "w") as f:
        json.dump({"f0": [200.0, 50.0], "energy": [energy_min, energy_max, 0.0, 1.0]}, f)
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({f"spk{i}": i for i in range(n_speakers)}, f)
    return d


def build_reference_model(dataset: str, energy_min=-1.5, energy_max=8.0):
    """(model, diffusion, configs) via the reference's own factory, script_util.py:56-75,
    mirroring synthesize.py:58-78."""
    import argparse

    import numpy as np

    install()
    from model.cm_tool.script_util import (args_to_dict, create_model_and_diffusion_tts,
                                           model_and_diffusion_defaults)

    preprocess_config, model_config, train_config = load_configs(dataset)
    preprocess_config["path"]["preprocessed_path"] = make_preprocessed_dir(energy_min, energy_max)
    # synthesize.py:333-336 computes this with pycwt; only len() is used (pitch_tools.py:246)
    preprocess_config["preprocessing"]["pitch"]["cwt_scales"] = np.arange(10, dtype=np.float64)
    args = argparse.Namespace(model="naive", T=1, restore_step=0)
    args_cm = argparse.Namespace(**train_config["cm"])
    distillation = "consistency" in args_cm.training_mode
    kw = args_to_dict(args_cm, model_and_diffusion_defaults().keys())
    kw["distillation"] = distillation
    kw["tts_model_config"] = {
        "args": args, "train_config": train_config,
        "preprocess_config": preprocess_config, "model_config": model_config,
    }
    model, diffusion = create_model_and_diffusion_tts(**kw)
    model.eval()
    return model, diffusion, (preprocess_config, model_config, train_config)


def build_reference_vocoder(state_dict=None, speaker="universal"):
    """hifigan.Generator built as utils/model.py:170-184 does (weights folded, eval)."""
    import torch

    install()
    import hifigan

    with open(os.path.join(REFERENCE_ROOT, "hifigan", "config.json")) as f:
        cfg = hifigan.AttrDict(json.load(f))
    voc = hifigan.Generator(cfg)
    if state_dict is None:
        ck = torch.load(os.path.join(REFERENCE_ROOT, "hifigan", f"generator_{speaker}.pth.tar"),
                        map_location="cpu", weights_only=True)
        state_dict = ck["generator"]
    voc.load_state_dict(state_dict)
    voc.eval()
    voc.remove_weight_norm()
    return voc


class ReplayGenerator:
    """`generator` seam of karras_sample_tts (karras_diffusion.py:498, :523, :852): replays
    pre-drawn CPU noise so the reference and the CUDA path consume identical tensors."""

    def __init__(self, seed: int = 1):
        import torch

        self.g = torch.Generator().manual_seed(seed)
        self.drawn = []

    def randn(self, *shape, device=None, **_):
        import torch

        t = torch.randn(*shape, generator=self.g)
        self.drawn.append(t)
        return t.to(device) if device is not None else t

    def randn_like(self, x):
        return self.randn(*x.shape, device=x.device)

    def randint(self, *a, **k):
        raise NotImplementedError
