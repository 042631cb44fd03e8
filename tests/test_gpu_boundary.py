"""The boundary as a user of the reference would drive it (synthesize.py:35-153, :195-227):

    YAML configs (the reference's own files) -> checkpoint FILE in the reference layout
      -> create_model_and_diffusion_tts -> CMTotalTTSSynthesize.synthesize(7-tuple) -> get_vocoder -> synth_samples

checked against the oracle; plus the p/e/d controls forwarded to the variance adaptor, the RTF of p_rtf_cm.py, the
checkpoint-file variants of train_util.py:890-917 and the loud failures the reference has (IndexError on a bad token id,
the speaker-embedding assert)."""
import argparse
import json
import os

import numpy as np
import pytest
import torch

from cmtts_b200 import synthetic
from cmtts_b200 import synthesize as S
from cmtts_b200.config import ModelSpec
from oracle import cmtts_oracle as O
from oracle import ref_shim

from gpu_util import DEV, Replay, draw_noise

pytestmark = pytest.mark.gpu
CONFIG_ROOT = os.path.join(ref_shim.REFERENCE_ROOT, "config")
needs_configs = pytest.mark.skipif(not os.path.isdir(CONFIG_ROOT), reason="the reference's YAML configs are not staged")


def _reference_configs(dataset, tmp_path):
    """The reference's three YAMLs, unchanged, plus the two files its constructors read from `preprocessed_path`
    (stats.json modules.py:233-237, speakers.json cmtts.py:27-37)."""
    pre, model, train = S.get_configs_of(dataset, CONFIG_ROOT)
    pp = tmp_path / "preprocessed"
    pp.mkdir()
    (pp / "stats.json").write_text(json.dumps({"f0": [200.0, 50.0], "energy": [-1.5, 8.0, 0.0, 1.0]}))
    (pp / "speakers.json").write_text(json.dumps({"spk0": 0}))
    pre["path"]["preprocessed_path"] = str(pp)
    pre["preprocessing"]["pitch"]["cwt_scales"] = [0.0] * 10          # synthesize.py:333-336 (only len() is used)
    return pre, model, train


@needs_configs
@pytest.mark.parametrize("dataset,T", [("VCTK", 2), ("LJSpeech", 1)])
def test_reference_protocol_end_to_end(tmp_path, dataset, T):
    from scipy.io import wavfile

    from cmtts_b200 import output
    from cmtts_b200.vocoder import get_vocoder

    pre, model_cfg, train = _reference_configs(dataset, tmp_path)
    spec = ModelSpec.from_reference_configs(pre, model_cfg, train)
    preset = ModelSpec.preset(dataset)
    for f in ("hidden", "enc_layers", "enc_heads", "ffn_kernel", "filter_size", "res_layers", "res_channels", "n_mels",
              "multi_speaker", "use_uv", "energy_bins", "pitch_bins", "sigma_min", "sigma_max", "hop_length", "sampling_rate"):
        assert getattr(spec, f) == getattr(preset, f), f
    # checkpoint FILE in the reference layout: <model_path>/CMDenoiserTTS/model000000.pt (synthesize.py:44-48)
    model_path = str(tmp_path / "ckpt")
    ck_file = synthetic.write_acoustic_checkpoint(model_path, spec, seed=4, step=0)
    assert ck_file.endswith(os.path.join("CMDenoiserTTS", "model000000.pt"))
    sd = torch.load(ck_file, map_location="cpu", weights_only=True)
    hck = synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)
    hdir = tmp_path / "hifigan"
    hdir.mkdir()
    torch.save(hck, str(hdir / f"generator_{model_cfg['vocoder']['speaker']}.pth.tar"))
    cfg_json = os.path.join(ref_shim.REFERENCE_ROOT, "hifigan", "config.json")

    args = argparse.Namespace(T=T, mode="batch", speaker_id="p225", teacher_forced=False, restore_step=0, model="naive")
    tool = S.CMTotalTTSSynthesize(model_path, 0, args, pre, model_cfg, train, device=DEV)   # create_model_and_diffusion_tts path
    assert tool.diffusion.distillation is True and tool.model.spec.multi_speaker == preset.multi_speaker

    b = synthetic.make_batch(spec, 3, 7, 15, seed=21)
    names = ["u0", "u1", "u2"]
    raw = ["t0", "t1", "t2"]
    batch7 = (names, raw, b["speakers"].numpy(), b["texts"].numpy(), b["src_lens"].numpy(), int(b["src_lens"].max()),
              None if b["spker_embeds"] is None else b["spker_embeds"].numpy())
    batch7 = S.to_device(batch7, DEV)
    W, Wf = O.Weights(sd), O.Weights(synthetic.fold_weight_norm(hck["generator"]))
    with torch.no_grad():
        pre_o = O.dpen(W, spec, **b)
    L = pre_o["cond"].shape[1]
    noise = draw_noise(9, (3, 1, L, spec.n_mels), T + 1)
    it = iter(noise)
    with torch.no_grad():
        mel_o, _, i16_o, _ = O.synthesize(W, Wf, spec, b, T, lambda s: next(it))
    out_put = tool.synthesize(batch7, generator=Replay(noise))
    torch.cuda.synchronize()
    assert len(out_put) == 12 and out_put[0].shape == (3, L, spec.n_mels)                   # synthesize.py:148-151
    assert torch.equal(out_put[11].cpu(), pre_o["mel_lens"]) and torch.equal(out_put[10].cpu(), b["src_lens"])
    assert float((out_put[0].cpu() - mel_o).abs().max()) <= 1e-3

    vocoder = get_vocoder(model_cfg, DEV, checkpoint_path=str(hdir / f"generator_{model_cfg['vocoder']['speaker']}.pth.tar"),
                          hifigan_config=cfg_json if os.path.isfile(cfg_json) else None)
    with output.AsyncWavWriter(2) as w:
        paths = output.synth_samples(args, batch7, out_put, vocoder, model_cfg, pre, str(tmp_path / "result"), tool.diffusion,
                                     writer=w)
    assert len(paths) == 3
    for p, ref in zip(paths, i16_o):
        rate, got = wavfile.read(p)
        assert rate == spec.sampling_rate and got.dtype == np.int16 and got.shape == ref.shape
        assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= 132               # fp16 vocoder: 4e-3 * 32768


def test_controls_are_forwarded_like_the_variance_adaptor_defines_them():
    """p/e/d_control != 1 (modules.py:331-412: e_control scales the energy prediction before bucketize, d_control the
    durations before rounding, p_control the cwt prediction) against the oracle with the same controls."""
    spec = ModelSpec.preset("LJSpeech")
    sd = synthetic.make_acoustic_state_dict(spec, seed=6)
    from cmtts_b200.model import CMTotalTTS
    model = CMTotalTTS(spec=spec).load_state_dict(sd).to(DEV)
    W = O.Weights(sd)
    checked = 0
    for seed in range(40, 60):
        b = synthetic.make_batch(spec, 3, 8, 16, seed=seed)
        # d_control = 2: the reference does not re-round `round(exp(log_d) - 1) * d_control` (modules.py:369-372), and with a
        # fractional product its LengthRegulator (int() truncation per token) and dur_to_mel2ph (cumsum of the floats)
        # disagree on the length, so the reference itself fails in the pitch-embedding add; integer products are the
        # well-defined case
        ctl = dict(p_control=1.15, e_control=0.85, d_control=2.0)
        with torch.no_grad():
            ref = O.dpen(W, spec, **b, **ctl)
            base = O.dpen(W, spec, **b)
        dp, _ = model.get_segmentation_model()
        out = dp(speakers=b["speakers"], texts=b["texts"], src_lens=b["src_lens"], spker_embeds=None, **ctl)
        torch.cuda.synchronize()
        if not torch.equal(out["d_rounded"].cpu(), ref["d_rounded"]) or not torch.equal(out["e_idx"].cpu(), ref["e_idx"]) \
                or not torch.equal(out["pitch_idx"].cpu(), ref["pitch_idx"]):
            continue                      # a quantiser input of this random batch sits on a cliff: try the next seed
        checked += 1
        assert not torch.equal(ref["d_rounded"], base["d_rounded"])            # the controls do something
        assert torch.equal(out["mel_lens"].cpu(), ref["mel_lens"]) and torch.equal(out["mel2ph"].cpu(), ref["mel2ph"])
        assert float((out["cond"].cpu() - ref["cond"]).abs().max()) <= 2e-5
        assert float((out["p_predictions"]["cwt"].cpu() - ref["cwt"]).abs().max()) <= 5e-5
        if checked == 3:
            break
    assert checked >= 2, "controls: too few cliff-free batches matched the oracle exactly"


def test_synthesize_forward_controls_switch_and_checkpoint_variants(tmp_path):
    """`forward_controls` (N3: the reference parses the controls, synthesize.py:275-292, but never forwards them,
    :96-102) and the trainer's other checkpoint files (train_util.py:890-917)."""
    spec = ModelSpec.preset("LJSpeech")
    root = str(tmp_path / "m")
    synthetic.write_acoustic_checkpoint(root, spec, seed=4, step=12)
    sd = synthetic.make_acoustic_state_dict(spec, seed=5)
    torch.save(sd, os.path.join(root, "CMDenoiserTTS", "ema_0.999_000012.pt"))
    args = argparse.Namespace(T=1)
    b = synthetic.make_batch(spec, 2, 6, 9, seed=3)
    batch7 = S.to_device((["a", "b"], ["x", "y"], b["speakers"].numpy(), b["texts"].numpy(), b["src_lens"].numpy(), 9, None), DEV)
    inert = S.CMTotalTTSSynthesize(root, 12, args, None, None, {"cm": {}}, d_control=2.0, device=DEV, spec=spec)
    live = S.CMTotalTTSSynthesize(root, 12, args, None, None, {"cm": {}}, d_control=2.0, device=DEV, spec=spec,
                                  forward_controls=True)
    ema = S.CMTotalTTSSynthesize(root, 12, args, None, None, {"cm": {}}, device=DEV, spec=spec, checkpoint="ema_0.999")
    o0, o1, o2 = inert.synthesize(batch7), live.synthesize(batch7), ema.synthesize(batch7)
    torch.cuda.synchronize()
    with torch.no_grad():
        r0 = O.dpen(O.Weights(inert.model.state_dict()), spec, **b)
        r1 = O.dpen(O.Weights(live.model.state_dict()), spec, **b, d_control=2.0)
        r2 = O.dpen(O.Weights(sd), spec, **b)
    assert torch.equal(o0[11].cpu(), r0["mel_lens"])                  # like the reference: the control is inert
    assert torch.equal(o1[11].cpu(), r1["mel_lens"]) and not torch.equal(r0["mel_lens"], r1["mel_lens"])
    assert torch.equal(o2[11].cpu(), r2["mel_lens"])                  # the EMA file was the one loaded
    with pytest.raises(ValueError):
        S.checkpoint_path(root, 12, "optimizer")


def test_rtf_like_reference_is_consistent():
    """R1: RTF as p_rtf_cm.py:190-230 defines it, for T = 1 and 4 (BASELINE metric 'at T=1/4')."""
    spec = ModelSpec.preset("LJSpeech")
    sd = synthetic.make_acoustic_state_dict(spec, seed=0)
    pipe = S.Pipeline(spec, sd, synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)["generator"], DEV)
    b = synthetic.make_batch(spec, 4, 20, 30, seed=2)
    t, l = b["texts"].to(DEV), b["src_lens"].to(DEV)
    for T in (1, 4):
        S.rtf_like_reference(pipe, t, l, None, T)                      # warm-up
        rtf_ref, rtf_total, elapsed = S.rtf_like_reference(pipe, t, l, None, T)
        out = pipe(t, l, None, T=T)
        lens = out["mel_lens"].cpu().tolist()
        dur0 = lens[0] * spec.hop_length / spec.sampling_rate
        assert elapsed > 0 and np.isfinite(rtf_ref) and abs(rtf_ref - elapsed / dur0) < 1e-9
        assert abs(rtf_total - elapsed / (sum(lens) * spec.hop_length / spec.sampling_rate)) < 1e-9
        assert rtf_total <= rtf_ref < 1.0                              # faster than real time, by a wide margin


def test_loud_failures_match_the_reference():
    spec = ModelSpec.preset("VCTK")
    from cmtts_b200.model import CMTotalTTS
    model = CMTotalTTS(spec=spec).load_state_dict(synthetic.make_acoustic_state_dict(spec, seed=1)).to(DEV)
    b = synthetic.make_batch(spec, 2, 5, 8, seed=1)
    with pytest.raises(AssertionError, match="Speaker embedding"):     # cmtts.py:80
        model.dpen(b["texts"], b["src_lens"], None)
    bad = b["texts"].clone()
    bad[0, 0] = spec.vocab                                             # nn.Embedding raises IndexError (modules.py:145)
    with pytest.raises(IndexError):
        model.dpen(bad, b["src_lens"], b["spker_embeds"])
    with pytest.raises(IndexError):
        model.dpen(bad.to(DEV), b["src_lens"].to(DEV), b["spker_embeds"].to(DEV))
    out = model.dpen(b["texts"], b["src_lens"], b["spker_embeds"])
    with pytest.raises(ValueError, match="shorter than the predicted length"):
        model.dpen(b["texts"], b["src_lens"], b["spker_embeds"], max_mel_len=int(out["mel_lens"].max()) - 1)


def test_cuda_graph_path_matches_eager():
    """Pipeline(graphs=True): first call of a shape eager, second captured, then replayed.  Everything upstream of the
    noise is deterministic and must be bit-identical to the eager pipeline; with the same seed the replayed device RNG
    is expected to give the same noise as eager launches (torch registers the generator with the graph) — if a torch
    build does not, the mels must at least agree in distribution."""
    spec = ModelSpec.preset("VCTK")
    sd = synthetic.make_acoustic_state_dict(spec, seed=0)
    hsd = synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)["generator"]
    eager = S.Pipeline(spec, sd, hsd, DEV)
    graphed = S.Pipeline(spec, sd, hsd, DEV, graphs=True)
    batches = [synthetic.make_batch(spec, 3, 8, 20, seed=s) for s in (31, 32)]
    batches[1]["texts"] = batches[1]["texts"][:, : batches[0]["texts"].shape[1]]      # same (B, Tsrc): same head graph
    if batches[1]["texts"].shape[1] < batches[0]["texts"].shape[1]:
        pytest.skip("seeds gave different widths")
    batches[1]["src_lens"] = batches[1]["src_lens"].clamp(max=batches[1]["texts"].shape[1])
    for T in (1, 4):
        for rep in range(4):
            b = batches[rep % 2] if rep >= 2 else batches[0]
            args = (b["texts"].to(DEV), b["src_lens"].to(DEV), b["spker_embeds"].to(DEV))
            torch.manual_seed(100 + rep)
            ref = eager(*args, T=T)
            torch.manual_seed(100 + rep)
            out = graphed(*args, T=T)
            torch.cuda.synchronize()
            assert torch.equal(out["mel_lens"], ref["mel_lens"])
            assert torch.equal(out["dpen"]["cond"], ref["dpen"]["cond"]) and torch.equal(out["dpen"]["mel2ph"], ref["dpen"]["mel2ph"])
            assert out["mel"].shape == ref["mel"].shape and torch.isfinite(out["mel"]).all()
            if not torch.equal(out["mel"], ref["mel"]):
                print(f"T={T} rep={rep}: replayed RNG stream differs from eager; comparing distributions")
                assert abs(float(out["mel"].mean() - ref["mel"].mean())) < 0.05 * float(ref["mel"].std())
                assert abs(float(out["mel"].std() / ref["mel"].std()) - 1) < 0.05
            else:
                assert torch.equal(out["wav_i16"], ref["wav_i16"])
    assert graphed.graph_replays >= 4 and graphed.graph_kernel_launches > 100
    # results are fresh tensors: a later call must not overwrite an earlier result
    a = graphed(*args, T=4)
    keep = a["wav_i16"].clone()
    graphed(*args, T=4)
    torch.cuda.synchronize()
    assert torch.equal(a["wav_i16"], keep)


@needs_configs
def test_zero_shot_protocol_ref_audio(tmp_path):
    """synthesize_zeroshot_lj.py:89-102: the batch's speaker embeddings are replaced by the DeepSpeaker embedding of ONE
    reference recording.  `CMTotalTTSSynthesize.synthesize(batch, ref_audio=...)` must equal `synthesize` on a batch that
    already carries that embedding, and the embedding must be the speaker encoder's (oracle-checked in
    tests/test_speaker_encoder.py) for the window the reference's random draw picks."""
    import random

    from scipy.io import wavfile

    from cmtts_b200 import speaker_encoder as SE

    ck = None
    for root in (os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref"), "/root/reference"):
        p = os.path.join(root, "deepspeaker", "pretrained_models", "ResCNN_triplet_training_checkpoint_265.h5")
        if os.path.isfile(p):
            ck = p
            break
    if ck is None:
        pytest.skip("reference DeepSpeaker checkpoint not staged")
    pre, model_cfg, train = _reference_configs("VCTK", tmp_path)
    spec = ModelSpec.from_reference_configs(pre, model_cfg, train)
    model_path = str(tmp_path / "ckpt")
    synthetic.write_acoustic_checkpoint(model_path, spec, seed=4, step=0)
    args = argparse.Namespace(T=2, mode="batch", speaker_id="p225", teacher_forced=False, restore_step=0, model="naive")
    tool = S.CMTotalTTSSynthesize(model_path, 0, args, pre, model_cfg, train, device=DEV)
    wav = str(tmp_path / "ref.wav")
    wavfile.write(wav, SE.SAMPLE_RATE, (synthetic.make_voice_like(2.5, SE.SAMPLE_RATE, seed=5) * 32767).astype(np.int16))

    b = synthetic.make_batch(spec, 3, 7, 15, seed=21)
    as7 = lambda emb: S.to_device((["u0", "u1", "u2"], ["t0", "t1", "t2"], b["speakers"].numpy(), b["texts"].numpy(),
                                   b["src_lens"].numpy(), int(b["src_lens"].max()), emb), DEV)
    noise = draw_noise(9, (3, 1, 1000, spec.n_mels), 3)

    class Cut:                                   # replayed noise, cut to whatever length the embedding leads to
        def __init__(self):
            self.it = iter(noise)

        def randn(self, *shape, device=None, **_):
            return next(self.it)[:, :, : shape[2]].contiguous().to(device)

        def randn_like(self, x):
            return self.randn(*x.shape, device=x.device)

    random.seed(123)
    out_zs = tool.synthesize(as7(b["spker_embeds"].numpy()), generator=Cut(), ref_audio=wav, speaker_ckpt=ck)
    random.seed(123)                             # the same window of the recording (batcher.py:25 draws it with `random`)
    emb = SE.get_deep_speaker_emb(filepath=wav, batch_size=3, device=DEV, ckpt_path=ck)
    assert emb.shape == (3, 512) and float((emb.norm(dim=1) - 1).abs().max()) < 1e-5
    out_pre = tool.synthesize(as7(emb.cpu().numpy()), generator=Cut())
    torch.cuda.synchronize()
    assert torch.equal(out_zs[11], out_pre[11]) and torch.equal(out_zs[0], out_pre[0])
    # and it is NOT what the batch's own embeddings give
    out_own = tool.synthesize(as7(b["spker_embeds"].numpy()), generator=Cut())
    assert out_own[0].shape != out_zs[0].shape or not torch.equal(out_own[0], out_zs[0])
