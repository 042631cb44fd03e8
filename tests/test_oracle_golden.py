"""Pin the oracle restatement (oracle/cmtts_oracle.py) to the reference's own outputs.

The fixtures under tests/golden/ were produced by oracle/make_golden.py from the unmodified
reference.  Tolerances: the restatement runs the same torch-CPU arithmetic in a
slightly different association order, so floating-point stages agree to a few fp32 ulps of their
magnitude (<= 5e-6 here); integer stages (durations, lengths) must be bit-exact.
"""
import glob
import os

import pytest
import torch

from cmtts_b200 import synthetic
from cmtts_b200.config import HifiGanSpec, ModelSpec
from oracle import cmtts_oracle as O
from oracle import ref_shim

from conftest import GOLDEN

ACOUSTIC = sorted(glob.glob(os.path.join(GOLDEN, "acoustic_*.pt")))


def _load(path):
    g = torch.load(path, map_location="cpu", weights_only=True)
    m = g["meta"]
    spec = ModelSpec.preset(m["dataset"])
    sd = synthetic.make_acoustic_state_dict(spec, m["weight_seed"])
    assert synthetic.state_dict_digest(sd) == m["digest"], "synthetic weight RNG stream drifted"
    batch = {"speakers": torch.zeros(m["batch"], dtype=torch.int64), "texts": g["texts"],
             "src_lens": g["src_lens"], "spker_embeds": g["spker_embeds"]}
    return g, m, spec, sd, batch


def test_fixtures_present():
    assert len(ACOUSTIC) >= 4
    for f in ("hifigan_synthetic.pt", "hifigan_universal.pt"):
        assert os.path.isfile(os.path.join(GOLDEN, f))


@pytest.mark.parametrize("path", ACOUSTIC, ids=[os.path.basename(p) for p in ACOUSTIC])
def test_dpen_matches_reference(path):
    g, m, spec, sd, batch = _load(path)
    # the synthetic batch generator must reproduce the stored inputs (except the hand-edited row)
    if m["single_phoneme_row"] < 0:
        b2 = synthetic.make_batch(spec, m["batch"], m["src_lo"], m["src_hi"], seed=m["batch_seed"])
        assert torch.equal(b2["texts"], g["texts"]) and torch.equal(b2["src_lens"], g["src_lens"])
    with torch.no_grad():
        d = O.dpen(O.Weights(sd), spec, **batch)
    assert torch.equal(d["d_rounded"], g["d_rounded"])            # integer-exact
    assert torch.equal(d["mel_lens"], g["mel_lens"])
    assert (d["enc"] - g["enc"]).abs().max() <= 5e-6
    assert (d["log_d_predictions"] - g["log_d"]).abs().max() <= 5e-6
    assert (d["e_predictions"] - g["e_pred"]).abs().max() <= 5e-6
    assert (d["cwt"] - g["cwt"]).abs().max() <= 1e-5
    assert (d["cond"] - g["cond"]).abs().max() <= 5e-6
    # f0 in Hz (values of several hundred): a few ulps
    assert (d["f0_denorm"] - g["f0_denorm"]).abs().max() <= 2e-3
    if g["speaker_emb"] is not None:
        assert (d["speaker_emb"] - g["speaker_emb"]).abs().max() <= 1e-6
    # scan form of the length regulator == the reference's B x T x L cube and python loop
    assert torch.equal(d["mel2ph"], O.dur_to_mel2ph_literal(d["d_rounded"], d["src_masks"]))
    lit, lens = O.length_regulate_literal(d["enc"], d["d_rounded"], None)
    fast, lens2 = O.length_regulate(d["enc"], d["d_rounded"], None)
    assert torch.equal(lit, fast) and torch.equal(lens, lens2)


@pytest.mark.parametrize("T", [1, 2, 4])
@pytest.mark.parametrize("path", ACOUSTIC, ids=[os.path.basename(p) for p in ACOUSTIC])
def test_sampler_matches_reference(path, T):
    g, m, spec, sd, batch = _load(path)
    gen = ref_shim.ReplayGenerator(m["noise_seed"])
    trace = {}
    with torch.no_grad():
        mel, pre = O.sample(O.Weights(sd), spec, batch, T, lambda shp: gen.randn(*shp), trace=trace)
    assert len(gen.drawn) == g[f"n_noise_T{T}"] == (1 if T == 1 else T + 1)  # x_T + one re-noise per iteration
    assert mel.shape == g[f"mel_T{T}"].shape
    # north_star tolerance for mels is 1e-3; the restatement is ~1000x tighter
    assert (mel - g[f"mel_T{T}"]).abs().max() <= 1e-5
    mo = trace["model_output"][0][:, 0].transpose(1, 2)  # (B,M,L) like Denoiser returns
    assert (mo - g[f"model_output0_T{T}"][:, 0]).abs().max() <= 2e-5


@pytest.mark.parametrize("path", ACOUSTIC[:1])
def test_literal_schedule_is_identical(path):
    """The reference re-runs encoder + variance adaptor in every solver step;
    computing the conditioner once gives the same bits."""
    g, m, spec, sd, batch = _load(path)
    W = O.Weights(sd)
    g1, g2 = ref_shim.ReplayGenerator(3), ref_shim.ReplayGenerator(3)
    with torch.no_grad():
        a, _ = O.sample(W, spec, batch, 2, lambda s: g1.randn(*s), literal=False)
        b, _ = O.sample(W, spec, batch, 2, lambda s: g2.randn(*s), literal=True)
    assert torch.equal(a, b)


def test_hifigan_synthetic_matches_reference():
    g = torch.load(os.path.join(GOLDEN, "hifigan_synthetic.pt"), weights_only=True)
    m = g["meta"]
    ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=m["weight_seed"])
    assert synthetic.state_dict_digest(ck["generator"]) == m["digest"]
    Wf = O.Weights(synthetic.fold_weight_norm(ck["generator"]))
    mel = synthetic.make_mels(m["batch"], 80, m["frames"], seed=m["mel_seed"])
    with torch.no_grad():
        wav = O.hifigan(Wf, HifiGanSpec(), mel)
    assert wav.shape == g["wav"].shape == (m["batch"], 1, 256 * m["frames"])
    assert (wav - g["wav"]).abs().max() <= 2e-6
    i16 = O.wav_to_int16(wav, None)
    diff = (torch.from_numpy(__import__("numpy").stack(i16)).int() - g["int16"].int()).abs()
    assert diff.max() <= 1  # a 1-ulp float difference may cross an integer boundary


def _real_weights():
    for p in (os.path.join(ref_shim.REFERENCE_ROOT, "hifigan", "generator_universal.pth.tar"),
              os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "hifigan", "generator_universal.pth.tar")):
        if os.path.isfile(p):
            return p
    return None


@pytest.mark.skipif(_real_weights() is None, reason="real HiFi-GAN weights not available")
def test_hifigan_real_weights_match_reference():
    g = torch.load(os.path.join(GOLDEN, "hifigan_universal.pt"), weights_only=True)
    m = g["meta"]
    sd = torch.load(_real_weights(), map_location="cpu", weights_only=True)["generator"]
    Wf = O.Weights(synthetic.fold_weight_norm(sd))
    mel = synthetic.make_mels(m["batch"], 80, m["frames"], seed=m["mel_seed"])
    with torch.no_grad():
        wav = O.hifigan(Wf, HifiGanSpec(), mel)
    assert (wav - g["wav"]).abs().max() <= 2e-6


@pytest.mark.reference
def test_oracle_matches_live_reference_vctk():
    """Live cross-check in the build container: unmodified reference vs restatement, new seeds."""
    spec = ModelSpec.preset("VCTK")
    sd = synthetic.make_acoustic_state_dict(spec, 11)
    model, diffusion, _ = ref_shim.build_reference_model("VCTK", spec.energy_min, spec.energy_max)
    model.load_state_dict(sd)
    batch = synthetic.make_batch(spec, 3, 5, 13, seed=5)
    dp, _ = model.get_segmentation_model()
    with torch.no_grad():
        ref = dp(**batch)
        mine = O.dpen(O.Weights(sd), spec, **batch)
    assert torch.equal(ref["d_rounded"], mine["d_rounded"])
    assert (ref["cond"] - mine["cond"]).abs().max() <= 5e-6


def test_edge_cases():
    # zero-duration phonemes are legal (modules.py:369-372) and emit no frames
    x = torch.arange(12, dtype=torch.float32).view(1, 4, 3)
    d = torch.tensor([[2.0, 0.0, 3.0, 0.0]])
    out, lens = O.length_regulate(x, d, None)
    lit, lens2 = O.length_regulate_literal(x, d, None)
    assert torch.equal(out, lit) and lens.tolist() == lens2.tolist() == [5]
    assert out[0, :, 0].tolist() == [0, 0, 6, 6, 6]
    m2p = O.dur_to_mel2ph(d, torch.zeros(1, 4, dtype=torch.bool))
    assert m2p.tolist() == [[1, 1, 3, 3, 3]]
    # padding to a longer max_len zero-fills
    out2, _ = O.length_regulate(x, d, 7)
    assert out2.shape == (1, 7, 3) and float(out2[0, 5:].abs().sum()) == 0.0
    # round-half-even + clamp
    ld = torch.log(torch.tensor([[1.5, 2.5, 3.5, 0.2]]))
    assert O.round_durations(ld).tolist() == [[0.0, 2.0, 2.0, 0.0]]
    with pytest.raises(ValueError):
        O.sampler_plan(3)


# ---- bench-size fixtures (oracle/make_fullsize_batches.py): BASELINE.json C2 / C3 / C4 at their per-GPU sizes ----
FULLSIZE = sorted(glob.glob(os.path.join(GOLDEN, "fullsize_*.pt")))


def test_fullsize_fixtures_present():
    assert [os.path.basename(p) for p in FULLSIZE] == ["fullsize_C2.pt", "fullsize_C3.pt", "fullsize_C4.pt"]


@pytest.mark.parametrize("path", FULLSIZE, ids=[os.path.basename(p) for p in FULLSIZE])
def test_oracle_matches_reference_at_bench_size(path):
    """The oracle against the reference's outputs at the sizes bench.py runs: integer stages bit-exact over the whole
    batch (the inputs are cliff-free by construction), log_d / energy within a few ulps, and the mels the reference's
    own sampler produced for the first and last utterance (T as the config names it, replayed noise)."""
    g = torch.load(path, map_location="cpu", weights_only=True)
    m = g["meta"]
    spec = ModelSpec.preset(m["dataset"])
    sd = synthetic.make_acoustic_state_dict(spec, m["weight_seed"])
    assert synthetic.state_dict_digest(sd) == m["digest"], "synthetic weight RNG stream drifted"
    batch = {"speakers": torch.zeros(m["batch"], dtype=torch.int64), "texts": g["texts"], "src_lens": g["src_lens"],
             "spker_embeds": g["spker_embeds"]}
    W = O.Weights(sd)
    gen = torch.Generator().manual_seed(m["noise_seed"])
    with torch.no_grad():
        mel, d = O.sample(W, spec, batch, m["T"], lambda s: torch.randn(*s, generator=gen))
    assert d["cond"].shape[1] == m["L"]
    assert torch.equal(d["d_rounded"].to(torch.int16), g["d_rounded"]) and torch.equal(d["mel_lens"], g["mel_lens"])
    assert torch.equal(d["e_idx"].to(torch.int16), g["e_idx"]) and torch.equal(d["pitch_idx"].to(torch.int16), g["pitch_idx"])
    assert torch.equal(d["mel2ph"].to(torch.int16), g["mel2ph"])
    assert (d["log_d_predictions"] - g["log_d"]).abs().max() <= 5e-6
    assert (d["e_predictions"] - g["e_pred"]).abs().max() <= 5e-6
    assert (mel[m["mel_rows"]] - g["mel_rows"]).abs().max() <= 2e-5
