"""GPU parity tests: the CUDA path (through the C ABI) against
  (1) the golden fixtures produced by the unmodified reference (tests/golden/), and
  (2) the oracle restatement run live on the host CPU, on fresh seeded inputs.

Tolerances.  north_star: mels within 1e-3 max-abs (fp32) of the reference; length-regulator
indices bit-exact.  The fp32 FFMA path differs from torch-CPU only by summation order, so the
tests hold it to much tighter bounds than the contract (stated per assert); quantiser outputs
(durations, energy bins, pitch bins, mel2ph, mel_lens) must be bit-exact on these seeds.
"""
import glob
import os

import numpy as np
import pytest
import torch

from cmtts_b200 import synthetic
from cmtts_b200.config import HifiGanSpec, ModelSpec
from oracle import cmtts_oracle as O

from conftest import GOLDEN
from gpu_util import (DEV, MEL_TOL as PREC_MEL_TOL, MODEL_OUT_TOL, WAV_TOL, Replay, draw_noise, gpu_model, load_golden,
                      real_hifigan_weights)

pytestmark = pytest.mark.gpu
ACOUSTIC = sorted(glob.glob(os.path.join(GOLDEN, "acoustic_*.pt")))
IDS = [os.path.basename(p) for p in ACOUSTIC]
MEL_TOL = 1e-3  # north_star


@pytest.mark.parametrize("path", ACOUSTIC, ids=IDS)
def test_dpen_vs_reference_golden(path):
    g, m, spec, sd, batch = load_golden(path)
    model = gpu_model(spec, sd, path)
    dp, _ = model.get_segmentation_model()
    out = dp(speakers=batch["speakers"], texts=batch["texts"], src_lens=batch["src_lens"],
             spker_embeds=batch["spker_embeds"])
    torch.cuda.synchronize()
    assert torch.equal(out["d_rounded"].cpu(), g["d_rounded"])          # bit-exact
    assert torch.equal(out["mel_lens"].cpu(), g["mel_lens"])            # bit-exact
    assert out["cond"].shape == g["cond"].shape
    assert (out["enc"].cpu() - g["enc"]).abs().max() <= 2e-5
    assert (out["log_d_predictions"].cpu() - g["log_d"]).abs().max() <= 2e-5
    assert (out["e_predictions"].cpu() - g["e_pred"]).abs().max() <= 5e-5
    assert (out["p_predictions"]["cwt"].cpu() - g["cwt"]).abs().max() <= 5e-5
    assert (out["cond"].cpu() - g["cond"]).abs().max() <= 2e-5
    assert (out["p_predictions"]["f0_denorm"].cpu() - g["f0_denorm"]).abs().max() <= 5e-2   # Hz
    # integer side products against the oracle restatement
    with torch.no_grad():
        ref = O.dpen(O.Weights(sd), spec, **batch)
    assert torch.equal(out["mel2ph"].cpu(), ref["mel2ph"])
    assert torch.equal(out["e_idx"].cpu(), ref["e_idx"])
    assert torch.equal(out["pitch_idx"].cpu(), ref["pitch_idx"])
    assert torch.equal(out["mel_masks"].cpu(), ref["mel_masks"])
    assert torch.equal(out["src_masks"].cpu(), ref["src_masks"])


@pytest.mark.parametrize("precision", ["tc", "fp32"])
@pytest.mark.parametrize("T", [1, 2, 4])
@pytest.mark.parametrize("path", ACOUSTIC, ids=IDS)
def test_sampler_vs_reference_golden(path, T, precision):
    from cmtts_b200.model import KarrasDenoiser
    from cmtts_b200.sampler import karras_sample_tts, sampler_plan
    g, m, spec, sd, batch = load_golden(path)
    model = gpu_model(spec, sd, path, precision)
    diffusion = KarrasDenoiser(distillation=True)
    B, L = g["cond"].shape[:2]
    noise = draw_noise(m["noise_seed"], (B, 1, L, spec.n_mels), g[f"n_noise_T{T}"])
    sampler, steps, ts = sampler_plan(T)
    kw = dict(speakers=batch["speakers"], texts=batch["texts"], src_lens=batch["src_lens"],
              spker_embeds=batch["spker_embeds"])
    trace = {}
    mel = karras_sample_tts(diffusion, model, (B, 1, L, spec.n_mels), steps=steps, model_kwargs=kw, device=DEV,
                            sigma_min=spec.sigma_min, sigma_max=spec.sigma_max, sampler=sampler, ts=ts,
                            generator=Replay(noise), trace=trace)
    torch.cuda.synchronize()
    err = (mel.cpu() - g[f"mel_T{T}"]).abs().max().item()
    assert err <= MEL_TOL, err                                     # north_star contract, all frames incl. padded
    assert err <= PREC_MEL_TOL[precision], f"{precision} path regressed: {err}"
    mo = trace["model_output"][0][:, 0].transpose(1, 2).cpu()      # (B,M,L) like Denoiser.forward
    assert (mo - g[f"model_output0_T{T}"][:, 0]).abs().max() <= MODEL_OUT_TOL[precision]


def test_forward_api_matches_fused_path():
    """CMTotalTTS.forward(x, t, **kw) (tts_net.py:75) re-derives the conditioner like the reference;
    the sampler's fused path must give the same bits."""
    from cmtts_b200.model import KarrasDenoiser
    g, m, spec, sd, batch = load_golden(ACOUSTIC[1])
    model = gpu_model(spec, sd, ACOUSTIC[1], "fp32")
    diffusion = KarrasDenoiser(distillation=True)
    B, L = g["cond"].shape[:2]
    x = draw_noise(3, (B, 1, L, spec.n_mels), 1)[0].to(DEV) * 80.0
    sig = torch.full((B,), 80.0, device=DEV)
    kw = {k: (v.to(DEV) if v is not None else None) for k, v in batch.items()}
    mo, den = diffusion.denoise(model, x, sig, **kw)
    cond = model.dpen(kw["texts"], kw["src_lens"], kw["spker_embeds"], L)
    c_skip, c_out, c_in, _ = diffusion.scalar_plan(80.0)
    t = 1000 * 0.25 * torch.log(torch.full((B,), 80.0) + 1e-44)
    steps = model.prepare_steps(t, cond["speaker_emb"])
    fused = model.denoise_step(x, cond["cond"], steps, c_in, c_out, c_skip)
    torch.cuda.synchronize()
    assert (fused - den).abs().max().item() <= 2e-5


@pytest.mark.parametrize("ds,B,lo,hi,T", [("LJSpeech", 4, 20, 45, 4), ("VCTK", 5, 8, 30, 1), ("LibriTTS", 3, 30, 60, 2)])
def test_pipeline_vs_oracle_fresh_inputs(ds, B, lo, hi, T):
    """Fresh seeds, moderate sizes (oracle takes seconds): mel within 1e-3 on ALL frames,
    integer outputs exact; padding coupling is covered because utterances are ragged."""
    from cmtts_b200.model import KarrasDenoiser
    from cmtts_b200.sampler import karras_sample_tts, sampler_plan
    spec = ModelSpec.preset(ds)
    sd = synthetic.make_acoustic_state_dict(spec, seed=21)
    batch = synthetic.make_batch(spec, B, lo, hi, seed=4321)
    model = gpu_model(spec, sd, ("fresh", ds))
    W = O.Weights(sd)
    with torch.no_grad():
        pre = O.dpen(W, spec, **batch)
    L = pre["cond"].shape[1]
    n_noise = 1 if T == 1 else T + 1
    noise = draw_noise(17, (B, 1, L, spec.n_mels), n_noise)
    it = iter(noise)
    with torch.no_grad():
        ref_mel, _ = O.sample(W, spec, batch, T, lambda s: next(it))
    sampler, steps, ts = sampler_plan(T)
    mel = karras_sample_tts(KarrasDenoiser(distillation=True), model, (B, 1, L, spec.n_mels), steps=steps,
                            model_kwargs=batch, device=DEV, sigma_min=spec.sigma_min, sigma_max=spec.sigma_max,
                            sampler=sampler, ts=ts, generator=Replay(noise))
    out = model.dpen(batch["texts"], batch["src_lens"], batch["spker_embeds"])
    torch.cuda.synchronize()
    assert torch.equal(out["d_rounded"].cpu(), pre["d_rounded"])
    assert torch.equal(out["mel_lens"].cpu(), pre["mel_lens"])
    assert torch.equal(out["mel2ph"].cpu(), pre["mel2ph"])
    flips = int((out["pitch_idx"].cpu() != pre["pitch_idx"]).sum()) + int((out["e_idx"].cpu() != pre["e_idx"]).sum())
    assert flips == 0, f"{flips} quantiser flips"
    assert (mel.cpu() - ref_mel).abs().max().item() <= MEL_TOL


def test_ffma_frontend_option():
    """Encoder + variance-adaptor GEMMs on the fp32 FFMA kernels (tc_frontend=False): the tightest path."""
    from cmtts_b200.model import CMTotalTTS
    g, m, spec, sd, batch = load_golden(ACOUSTIC[1])
    model = CMTotalTTS(spec=spec, precision="tc").load_state_dict(sd).to(DEV)
    model.tc_frontend = False
    out = model.dpen(batch["texts"], batch["src_lens"], batch["spker_embeds"])
    torch.cuda.synchronize()
    assert torch.equal(out["d_rounded"].cpu(), g["d_rounded"])
    assert torch.equal(out["mel_lens"].cpu(), g["mel_lens"])
    assert (out["enc"].cpu() - g["enc"]).abs().max() <= 5e-6
    assert (out["cond"].cpu() - g["cond"]).abs().max() <= 5e-6


def test_padding_coupling_matches_reference_in_each_case():
    """an utterance alone vs inside a longer batch gives DIFFERENT results in the
    reference (unmasked denoiser, CWT stats over padded frames); the CUDA path must equal the
    oracle in each case."""
    spec = ModelSpec.preset("LJSpeech")
    sd = synthetic.make_acoustic_state_dict(spec, seed=21)
    model = gpu_model(spec, sd, ("fresh", "LJSpeech"))
    W = O.Weights(sd)
    full = synthetic.make_batch(spec, 3, 10, 40, seed=9)
    short = int(full["src_lens"].argmin())
    n = int(full["src_lens"][short])
    alone = {"speakers": full["speakers"][short:short + 1], "texts": full["texts"][short:short + 1, :n].contiguous(),
             "src_lens": full["src_lens"][short:short + 1], "spker_embeds": None}
    conds = []
    for batch in (full, alone):
        with torch.no_grad():
            ref = O.dpen(W, spec, **batch)
        out = model.dpen(batch["texts"], batch["src_lens"], None)
        assert (out["cond"].cpu() - ref["cond"]).abs().max() <= 2e-5
        conds.append(ref["cond"])
    ml = int(conds[1].shape[1])
    assert (conds[0][short, :ml] - conds[1][0]).abs().max() > 1e-3   # they differ by design


def _snr_db(ref, got):
    return float(10 * torch.log10(ref.pow(2).mean() / (got - ref).pow(2).mean().clamp_min(1e-30)))


@pytest.mark.parametrize("precision", ["tc", "fp32"])
def test_hifigan_vs_reference_golden_synthetic(precision):
    from cmtts_b200.vocoder import Generator
    g = torch.load(os.path.join(GOLDEN, "hifigan_synthetic.pt"), weights_only=True)
    m = g["meta"]
    ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=m["weight_seed"])
    voc = Generator(hspec=HifiGanSpec(), precision=precision).load_state_dict(ck["generator"]).to(DEV)
    mel = synthetic.make_mels(m["batch"], 80, m["frames"], seed=m["mel_seed"])
    wav = voc(mel.to(DEV))
    torch.cuda.synchronize()
    assert wav.shape == g["wav"].shape
    assert (wav.cpu() - g["wav"]).abs().max().item() <= WAV_TOL[precision]
    assert _snr_db(g["wav"], wav.cpu()) >= (50.0 if precision == "tc" else 100.0)


@pytest.mark.skipif(real_hifigan_weights() is None, reason="real HiFi-GAN weights not shipped to this box")
def test_hifigan_tensor_core_real_weights_snr():
    """fp16-operand tensor-core vocoder on the reference's shipped universal checkpoint."""
    from cmtts_b200.vocoder import Generator
    g = torch.load(os.path.join(GOLDEN, "hifigan_universal.pt"), weights_only=True)
    m = g["meta"]
    sd = torch.load(real_hifigan_weights(), map_location="cpu", weights_only=True)["generator"]
    voc = Generator(hspec=HifiGanSpec(), precision="tc").load_state_dict(sd).to(DEV)
    mel = synthetic.make_mels(m["batch"], 80, m["frames"], seed=m["mel_seed"])
    wav = voc(mel.to(DEV))
    torch.cuda.synchronize()
    assert (wav.cpu() - g["wav"]).abs().max().item() <= WAV_TOL["tc"]
    assert _snr_db(g["wav"], wav.cpu()) >= 48.0


@pytest.mark.skipif(real_hifigan_weights() is None, reason="real HiFi-GAN weights not shipped to this box")
def test_hifigan_vs_reference_golden_real_weights():
    from cmtts_b200.vocoder import Generator, vocoder_infer
    g = torch.load(os.path.join(GOLDEN, "hifigan_universal.pt"), weights_only=True)
    m = g["meta"]
    sd = torch.load(real_hifigan_weights(), map_location="cpu", weights_only=True)["generator"]
    voc = Generator(hspec=HifiGanSpec(), precision="fp32").load_state_dict(sd).to(DEV)
    mel = synthetic.make_mels(m["batch"], 80, m["frames"], seed=m["mel_seed"])
    wav = voc(mel.to(DEV))
    torch.cuda.synchronize()
    assert (wav.cpu() - g["wav"]).abs().max().item() <= 2e-5
    cfgs = ({"vocoder": {"model": "HiFi-GAN"}}, {"preprocessing": {"audio": {"max_wav_value": 32768.0}}})
    i16 = vocoder_infer(mel.to(DEV), voc, cfgs[0], cfgs[1], lengths=[256 * m["frames"] - 100])
    assert i16[0].dtype == np.int16 and i16[0].shape[0] == 256 * m["frames"] - 100
    diff = np.abs(i16[0].astype(np.int32) - g["int16"][0, : i16[0].shape[0]].numpy().astype(np.int32))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.99


@pytest.mark.parametrize("precision", ["tc", "fp32"])
def test_hifigan_vs_oracle_ragged_batch(precision):
    from cmtts_b200.vocoder import Generator
    ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=3)
    voc = Generator(hspec=HifiGanSpec(), precision=precision).load_state_dict(ck["generator"]).to(DEV)
    Wf = O.Weights(synthetic.fold_weight_norm(ck["generator"]))
    for (B, L) in [(1, 1), (3, 37), (2, 130)]:
        mel = synthetic.make_mels(B, 80, L, seed=5 + L)
        with torch.no_grad():
            ref = O.hifigan(Wf, HifiGanSpec(), mel)
        wav = voc(mel.to(DEV))
        torch.cuda.synchronize()
        assert wav.shape == ref.shape
        assert (wav.cpu() - ref).abs().max().item() <= WAV_TOL[precision]


def test_hifigan_vs_oracle_bench_length():
    """A bench-length utterance pair (L = 801 frames -> 205k samples): every CTA of the persistent vocoder kernels runs
    many tiles, which short fixtures never exercise.  Checked against the CPU oracle (max-abs and SNR)."""
    from cmtts_b200.vocoder import Generator
    ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=3)
    voc = Generator(hspec=HifiGanSpec(), precision="tc").load_state_dict(ck["generator"]).to(DEV)
    Wf = O.Weights(synthetic.fold_weight_norm(ck["generator"]))
    mel = synthetic.make_mels(2, 80, 801, seed=11)
    with torch.no_grad():
        ref = O.hifigan(Wf, HifiGanSpec(), mel)
    wav = voc(mel.to(DEV)).cpu()
    assert wav.shape == ref.shape
    err = (wav - ref).abs()
    assert err.max().item() <= WAV_TOL["tc"]
    snr = 10 * torch.log10((ref ** 2).mean() / (err ** 2).mean()).item()
    assert snr >= 48.0, snr


def test_whole_pipeline_int16_vs_oracle():
    from cmtts_b200.synthesize import Pipeline
    spec = ModelSpec.preset("VCTK")
    sd = synthetic.make_acoustic_state_dict(spec, seed=2)
    ck = synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)
    pipe = Pipeline(spec, sd, ck["generator"], DEV)
    batch = synthetic.make_batch(spec, 3, 6, 14, seed=8)
    W, Wf = O.Weights(sd), O.Weights(synthetic.fold_weight_norm(ck["generator"]))
    with torch.no_grad():
        pre = O.dpen(W, spec, **batch)
    L = pre["cond"].shape[1]
    noise = draw_noise(5, (3, 1, L, 80), 3)
    it = iter(noise)
    with torch.no_grad():
        mel, wav, i16, _ = O.synthesize(W, Wf, spec, batch, 2, lambda s: next(it))
    out = pipe(batch["texts"], batch["src_lens"], batch["spker_embeds"], T=2, generator=Replay(noise), want_float_wav=True)
    torch.cuda.synchronize()
    assert (out["mel"].cpu() - mel).abs().max() <= MEL_TOL
    assert (out["wav"].cpu() - wav.squeeze(1)).abs().max() <= WAV_TOL["tc"]
    got = pipe.crop(out["wav_i16"].cpu().numpy(), out["mel_lens"].cpu().tolist())
    for a, b in zip(got, i16):
        assert a.shape == b.shape
        assert np.abs(a.astype(np.int32) - b.astype(np.int32)).max() <= 132   # 4e-3 * 32768 (fp16 vocoder)


def test_edge_cases():
    spec = ModelSpec.preset("LJSpeech")
    sd = synthetic.make_acoustic_state_dict(spec, seed=21)
    model = gpu_model(spec, sd, ("fresh", "LJSpeech"))
    # single phoneme, batch of one
    texts = torch.tensor([[5]]); lens = torch.tensor([1])
    with torch.no_grad():
        ref = O.dpen(O.Weights(sd), spec, torch.zeros(1, dtype=torch.long), texts, lens)
    out = model.dpen(texts, lens, None)
    assert torch.equal(out["mel_lens"].cpu(), ref["mel_lens"])
    if ref["cond"].shape[1] > 1:
        assert (out["cond"].cpu() - ref["cond"]).abs().max() <= 2e-5
    # multi-speaker model without embeddings must raise like cmtts.py:80
    vspec = ModelSpec.preset("VCTK")
    vm = gpu_model(vspec, synthetic.make_acoustic_state_dict(vspec, 0), ("edge", "VCTK"))
    with pytest.raises(AssertionError):
        vm.dpen(texts, lens, None)
    with pytest.raises(ValueError):
        from cmtts_b200.sampler import sampler_plan
        sampler_plan(3)


@pytest.mark.parametrize("ds,B,lo,hi,T,parts", [("VCTK", 6, 10, 40, 4, (slice(0, 2), slice(2, 6))),
                                                  # BASELINE.json configs[1] at full size (C2): B=32, L~800, T=4
                                                  ("LJSpeech", 32, 80, 115, 4, (slice(0, 8), slice(24, 32)))])
def test_acoustic_path_is_batch_invariant_under_global_padding(ds, B, lo, hi, T, parts):
    """Utterances are independent once Tsrc_max and L_max are fixed: a sub-batch padded like the full batch
    must reproduce its rows of the batched run BIT FOR BIT (mels and int16 wavs) — the property the multi-GPU sharding
    relies on, and a size-independent check of the whole path at the bench size."""
    from cmtts_b200.synthesize import Pipeline
    spec = ModelSpec.preset(ds)
    sd = synthetic.make_acoustic_state_dict(spec, seed=4)
    ck = synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)
    pipe = Pipeline(spec, sd, ck["generator"], DEV)
    batch = synthetic.make_batch(spec, B, lo, hi, seed=12)
    pre = pipe.model.dpen(batch["texts"], batch["src_lens"], batch["spker_embeds"], None)
    L = pre["cond"].shape[1]
    noise = draw_noise(9, (B, 1, L, 80), T + 1)
    full = pipe(batch["texts"], batch["src_lens"], batch["spker_embeds"], T=T, generator=Replay(noise))
    full = {k: full[k].clone() for k in ("mel", "mel_lens", "wav_i16")}
    again = pipe(batch["texts"], batch["src_lens"], batch["spker_embeds"], T=T, generator=Replay(noise))
    assert torch.equal(again["mel"], full["mel"]) and torch.equal(again["wav_i16"], full["wav_i16"])      # repeatable
    for rows in parts:
        spk = None if batch["spker_embeds"] is None else batch["spker_embeds"][rows].contiguous()
        sub = pipe(batch["texts"][rows].contiguous(), batch["src_lens"][rows].contiguous(), spk, T=T,
                   generator=Replay([n[rows].contiguous() for n in noise]), l_max_hook=lambda local_max: L)
        torch.cuda.synchronize()
        assert torch.equal(sub["mel_lens"], full["mel_lens"][rows])
        assert torch.equal(sub["mel"], full["mel"][rows])
        assert torch.equal(sub["wav_i16"], full["wav_i16"][rows])


@pytest.mark.parametrize("precision", ["tc", "fp32"])
def test_empty_batch_through_the_pipeline(precision):
    """A rank of a sharded run can hold ZERO utterances (more GPUs than rows): every stage has to accept B = 0 and return
    empty tensors (the 8-rank equality check caught `denoiser_forward_tc` rejecting its null conditioner planes — no
    single-GPU test ran an empty batch)."""
    from cmtts_b200.synthesize import Pipeline
    spec = ModelSpec.preset("VCTK")
    sd = synthetic.make_acoustic_state_dict(spec, seed=0)
    ck = synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)
    pipe = Pipeline(spec, sd, ck["generator"], DEV, precision=precision)
    texts = torch.zeros(0, 9, dtype=torch.int64)
    out = pipe(texts, torch.zeros(0, dtype=torch.int64), torch.zeros(0, spec.ext_speaker_dim), T=4)
    torch.cuda.synchronize()
    assert out["mel"].shape[0] == 0 and out["wav_i16"].shape[0] == 0 and out["mel_lens"].numel() == 0
    # and a real batch afterwards on the same pipeline (grow-only workspaces sized by the empty call)
    batch = synthetic.make_batch(spec, 2, 5, 9, seed=1)
    out = pipe(batch["texts"], batch["src_lens"], batch["spker_embeds"], T=2)
    torch.cuda.synchronize()
    assert out["mel"].shape[0] == 2 and torch.isfinite(out["mel"]).all()
