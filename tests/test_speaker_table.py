"""`speaker_embedder: none` (cmtts.py:27-37, :76-78): the nn.Embedding speaker table indexed by the batch's speaker ids.

The device path feeds one-hot rows through the (padded) table as a Linear's weight, which is exact; the oracle is driven
the same way (Linear with weight = table^T, zero bias), and a live test against the unmodified reference (in the build
container) pins that equivalence to `self.speaker_emb(speakers)`."""
import dataclasses
import json
import os

import pytest
import torch

from cmtts_b200 import synthetic
from cmtts_b200.config import ModelSpec
from oracle import cmtts_oracle as O
from oracle import ref_shim

SPK = "duration_pitch_energy_net.speaker_emb."
N_SPK = 11


def table_spec():
    return dataclasses.replace(ModelSpec.preset("VCTK"), n_speakers=N_SPK, ext_speaker_dim=16)


def oracle_equivalent(sd, spec, speakers):
    """state_dict + speaker input under which the oracle's Linear reproduces the table lookup exactly"""
    sd2 = dict(sd)
    sd2[SPK + "weight"] = sd[SPK + "weight"].t().contiguous()            # (H, n_speakers)
    sd2[SPK + "bias"] = torch.zeros(spec.hidden)
    onehot = torch.nn.functional.one_hot(speakers, N_SPK).float()
    return sd2, dataclasses.replace(spec, ext_speaker_dim=N_SPK, n_speakers=0), onehot


def test_spec_from_reference_configs_with_speaker_table(tmp_path):
    root = os.path.join(ref_shim.REFERENCE_ROOT, "config")
    if not os.path.isdir(root):
        pytest.skip("the reference's YAML configs are not staged")
    pre, model, train = ref_shim.load_configs("VCTK")
    pp = tmp_path / "pre"
    pp.mkdir()
    (pp / "stats.json").write_text(json.dumps({"f0": [200.0, 50.0], "energy": [-1.5, 8.0, 0.0, 1.0]}))
    (pp / "speakers.json").write_text(json.dumps({f"p{i}": i for i in range(109)}))
    pre["path"]["preprocessed_path"] = str(pp)
    pre["preprocessing"]["speaker_embedder"] = "none"
    spec = ModelSpec.from_reference_configs(pre, model, train)
    assert spec.multi_speaker and spec.n_speakers == 109 and spec.ext_speaker_dim == 112
    sd = synthetic.make_acoustic_state_dict(spec, seed=0)
    assert tuple(sd[SPK + "weight"].shape) == (109, spec.hidden) and (SPK + "bias") not in sd


def test_speaker_input_is_one_hot_and_checks_ids():
    from cmtts_b200.model import CMTotalTTS
    m = CMTotalTTS.__new__(CMTotalTTS)                                    # host logic only: no device, no library
    m.spec = table_spec()
    x = m.speaker_input(torch.tensor([3, 0, 10]), None)
    assert x.shape == (3, 16) and x.sum().item() == 3.0 and x[0, 3] == 1 and x[2, 10] == 1
    with pytest.raises(IndexError):
        m.speaker_input(torch.tensor([11]), None)
    with pytest.raises(AssertionError):
        m.speaker_input(None, None)
    m.spec = ModelSpec.preset("VCTK")
    e = torch.randn(2, 512)
    assert m.speaker_input(torch.tensor([0, 1]), e) is e


@pytest.mark.reference
def test_one_hot_linear_is_the_references_embedding_lookup():
    """unmodified reference built with speaker_embedder 'none' vs the oracle driven through the one-hot equivalence"""
    spec = table_spec()
    sd = synthetic.make_acoustic_state_dict(spec, seed=4)
    ref_shim.install()
    import argparse
    import numpy as np
    from model.cm_tool.script_util import args_to_dict, create_model_and_diffusion_tts, model_and_diffusion_defaults
    pre, model_cfg, train = ref_shim.load_configs("VCTK")
    pre["path"]["preprocessed_path"] = ref_shim.make_preprocessed_dir(spec.energy_min, spec.energy_max, n_speakers=N_SPK)
    pre["preprocessing"]["pitch"]["cwt_scales"] = np.arange(10, dtype=np.float64)
    pre["preprocessing"]["speaker_embedder"] = "none"
    args_cm = argparse.Namespace(**train["cm"])
    kw = args_to_dict(args_cm, model_and_diffusion_defaults().keys())
    kw["distillation"] = "consistency" in args_cm.training_mode
    kw["tts_model_config"] = {"args": argparse.Namespace(model="naive", T=1, restore_step=0), "train_config": train,
                              "preprocess_config": pre, "model_config": model_cfg}
    model, _ = create_model_and_diffusion_tts(**kw)
    model.eval()
    model.load_state_dict(sd)
    batch = synthetic.make_batch(spec, 3, 5, 13, seed=5)
    batch["speakers"] = torch.tensor([7, 0, 10])
    batch["spker_embeds"] = None
    dp, _ = model.get_segmentation_model()
    sd2, spec2, onehot = oracle_equivalent(sd, spec, batch["speakers"])
    with torch.no_grad():
        ref = dp(**batch)
        mine = O.dpen(O.Weights(sd2), spec2, batch["speakers"], batch["texts"], batch["src_lens"], onehot)
    assert torch.equal(ref["d_rounded"], mine["d_rounded"])
    assert (ref["speaker_emb"] - sd[SPK + "weight"][batch["speakers"]]).abs().max() == 0
    assert (ref["speaker_emb"] - mine["speaker_emb"]).abs().max() == 0
    assert (ref["cond"] - mine["cond"]).abs().max() <= 5e-6


@pytest.mark.gpu
def test_speaker_table_path_vs_oracle():
    from cmtts_b200.model import CMTotalTTS
    from gpu_util import DEV
    spec = table_spec()
    sd = synthetic.make_acoustic_state_dict(spec, seed=4)
    batch = synthetic.make_batch(spec, 4, 6, 17, seed=8)
    batch["speakers"] = torch.tensor([7, 0, 10, 7])
    model = CMTotalTTS(spec=spec).load_state_dict(sd).to(DEV)
    dp, _ = model.get_segmentation_model()
    out = dp(speakers=batch["speakers"], texts=batch["texts"], src_lens=batch["src_lens"], spker_embeds=None)
    sd2, spec2, onehot = oracle_equivalent(sd, spec, batch["speakers"])
    with torch.no_grad():
        want = O.dpen(O.Weights(sd2), spec2, batch["speakers"], batch["texts"], batch["src_lens"], onehot)
    # the embedding rows come through the one-hot Linear bit-exactly
    assert torch.equal(out["speaker_emb"].cpu().reshape(4, -1), sd[SPK + "weight"][batch["speakers"]])
    assert torch.equal(out["d_rounded"].cpu(), want["d_rounded"]) and torch.equal(out["mel_lens"].cpu(), want["mel_lens"])
    assert (out["cond"].cpu() - want["cond"]).abs().max() <= 2e-5
    with pytest.raises(IndexError):
        dp(speakers=torch.tensor([7, 0, 10, N_SPK]), texts=batch["texts"], src_lens=batch["src_lens"], spker_embeds=None)
