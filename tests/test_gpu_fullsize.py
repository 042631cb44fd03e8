"""Bench-size parity (BASELINE.json C2 / C3 / C4 at their per-GPU sizes) of the CUDA path against the oracle AND
against the reference's own outputs stored in tests/golden/fullsize_*.pt (oracle/make_fullsize_batches.py).

These are the cases where every persistent CTA of the tcgen05 kernels runs SEVERAL tiles (C2: 199 row tiles of the
flattened denoiser on 148 SMs) — the regime that hid a shared-memory race from the small fixtures in round 1.  The
inputs are cliff-free by construction (every quantiser input clears a margin, see the generator), so the integer stages
must be bit-exact over the whole batch and the mels within north_star's 1e-3 on EVERY frame, padded ones included.
"""
import os

import numpy as np
import pytest
import torch

from cmtts_b200 import synthetic
from cmtts_b200.config import ModelSpec
from oracle import cmtts_oracle as O

from conftest import GOLDEN
from gpu_util import DEV, Replay, gpu_model, real_hifigan_weights

pytestmark = pytest.mark.gpu
MEL_TOL = 1e-3   # north_star: mels within 1e-3 max-abs (fp32) of the reference


def _load(tag):
    g = torch.load(os.path.join(GOLDEN, f"fullsize_{tag}.pt"), map_location="cpu", weights_only=True)
    m = g["meta"]
    spec = ModelSpec.preset(m["dataset"])
    sd = synthetic.make_acoustic_state_dict(spec, m["weight_seed"])
    assert synthetic.state_dict_digest(sd) == m["digest"], "synthetic weight RNG stream drifted"
    batch = {"speakers": torch.zeros(m["batch"], dtype=torch.int64), "texts": g["texts"], "src_lens": g["src_lens"],
             "spker_embeds": g["spker_embeds"]}
    return g, m, spec, sd, batch


def _noise(m, B, L, n_mels):
    gen = torch.Generator().manual_seed(m["noise_seed"])          # the stream oracle/ref_shim.ReplayGenerator draws
    return [torch.randn(B, 1, L, n_mels, generator=gen) for _ in range(m["n_noise"])]


@pytest.mark.parametrize("tag", ["C2", "C3", "C4"])
def test_fullsize_acoustic_path_vs_oracle_and_reference(tag):
    from cmtts_b200.sampler import karras_sample_tts, sampler_plan

    g, m, spec, sd, batch = _load(tag)
    B, L, T = m["batch"], m["L"], m["T"]
    W = O.Weights(sd)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        pre = O.dpen(W, spec, **batch)
    # the oracle on THIS host reproduces the reference's integers stored in the fixture
    assert torch.equal(pre["d_rounded"].to(torch.int16), g["d_rounded"]) and torch.equal(pre["mel_lens"], g["mel_lens"])
    assert torch.equal(pre["e_idx"].to(torch.int16), g["e_idx"]) and torch.equal(pre["pitch_idx"].to(torch.int16), g["pitch_idx"])

    model = gpu_model(spec, sd, "fullsize_" + tag)
    out = model.dpen(batch["texts"], batch["src_lens"], batch["spker_embeds"])
    torch.cuda.synchronize()
    assert out["cond"].shape == (B, L, spec.hidden)
    # integer stages: bit-exact over the whole batch
    assert torch.equal(out["d_rounded"].cpu(), pre["d_rounded"])
    assert torch.equal(out["mel_lens"].cpu(), g["mel_lens"])
    assert torch.equal(out["mel2ph"].cpu(), pre["mel2ph"])
    assert torch.equal(out["e_idx"].cpu(), pre["e_idx"])
    n_pitch = int((out["pitch_idx"].cpu() != pre["pitch_idx"]).sum())
    assert n_pitch == 0, f"{n_pitch} pitch bins differ"
    errs = {"log_d": float((out["log_d_predictions"].cpu() - pre["log_d_predictions"]).abs().max()),
            "e_pred": float((out["e_predictions"].cpu() - pre["e_predictions"]).abs().max()),
            "cwt": float((out["p_predictions"]["cwt"].cpu() - pre["cwt"]).abs().max()),
            "f0_hz": float((out["p_predictions"]["f0_denorm"].cpu() - pre["f0_denorm"]).abs().max()),
            "cond": float((out["cond"].cpu() - pre["cond"]).abs().max())}
    print(f"{tag}: dpen float errors vs oracle {errs}")
    assert errs["log_d"] <= 2e-5 and errs["e_pred"] <= 5e-5 and errs["cwt"] <= 5e-5 and errs["cond"] <= 2e-5

    # sampler with replayed noise: every frame (padded included) within the contract
    noise = _noise(m, B, L, spec.n_mels)
    sampler, steps, ts = sampler_plan(T)
    kw = {"texts": batch["texts"], "src_lens": batch["src_lens"], "spker_embeds": batch["spker_embeds"]}
    mel = karras_sample_tts(model_diffusion(spec), model, (B, 1, L, spec.n_mels), steps=steps, model_kwargs=kw, device=DEV,
                            sigma_min=spec.sigma_min, sigma_max=spec.sigma_max, sampler=sampler, ts=ts,
                            generator=Replay(noise), cond_dict=out)
    torch.cuda.synchronize()
    mel = mel.cpu()
    assert torch.isfinite(mel).all()
    it = iter(noise)
    with torch.no_grad():
        mel_o, _ = O.sample(W, spec, batch, T, lambda s: next(it))
    e_oracle = float((mel - mel_o).abs().max())
    e_ref = float((mel[m["mel_rows"]] - g["mel_rows"]).abs().max())
    e_oracle_ref = float((mel_o[m["mel_rows"]] - g["mel_rows"]).abs().max())
    print(f"{tag}: B={B} L={L} T={T} mel max-abs error vs oracle {e_oracle:.2e}, vs the reference's rows {e_ref:.2e} "
          f"(oracle vs reference {e_oracle_ref:.2e})")
    assert e_oracle_ref <= 2e-5
    assert e_oracle <= MEL_TOL and e_ref <= MEL_TOL


def model_diffusion(spec):
    from cmtts_b200.model import KarrasDenoiser
    return KarrasDenoiser(sigma_data=spec.sigma_data, sigma_max=spec.sigma_max, sigma_min=spec.sigma_min, rho=spec.rho,
                          distillation=True)


def test_fullsize_vocoder_rows_vs_oracle():
    """C2-size vocoder call (B=32, L=810: every CTA of the persistent kernels runs many tiles) against the oracle run
    on four of its rows alone (HiFi-GAN is per-row), real universal weights when staged.  Rows 1 / 30 hold log-mel-like
    inputs (N(-5, 2^2) clipped): the fp16-operand tensor-core path must stay above 48 dB and within 6e-3
    max-abs (measured 4.2e-3 at 49.5 dB with the trained universal weights at this length; the 4e-3 of the short
    fixtures was set on synthetic weights).
    Rows 0 / 31 hold the mels the reference's sampler produced from the SYNTHETIC acoustic weights — not speech-like, they
    drive the trained generator into saturation, where fp16 storage of large activations costs absolute accuracy (3.7e-2
    max-abs measured at 50.8 dB): for them the tensor-core path is held to the SNR only, and the fp32 FFMA path (same
    tiling-independent arithmetic as the reference) proves that there is no structural error at this size."""
    from cmtts_b200.vocoder import Generator

    g, m, spec, sd, batch = _load("C2")
    p = real_hifigan_weights()
    gen_sd = (torch.load(p, map_location="cpu", weights_only=True)["generator"] if p
              else synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)["generator"])
    B, L = m["batch"], m["L"]
    mel = synthetic.make_mels(B, spec.n_mels, L, seed=17).transpose(1, 2).contiguous()      # (B, L, 80)
    hard = m["mel_rows"]                                                                      # [0, B - 1]
    mel[hard] = g["mel_rows"].clamp(-11.5, 2.0)
    easy = [1, B - 2]
    rows = hard + easy
    Wf = O.Weights(synthetic.fold_weight_norm(gen_sd))
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = O.hifigan(Wf, spec.hifigan, mel[rows].transpose(1, 2)).squeeze(1)

    def stats(got, r):
        err = float((got - r).abs().max())
        snr = float(10 * torch.log10(r.pow(2).mean() / (got - r).pow(2).mean().clamp_min(1e-30)))
        return err, snr

    for precision in ("fp32", "tc"):
        voc = Generator(hspec=spec.hifigan, precision=precision).load_state_dict(gen_sd).to(DEV)
        wav, w16 = voc.run(mel.to(DEV), want_float=True, want_int16=True)
        torch.cuda.synchronize()
        assert torch.isfinite(wav).all()
        got = wav.cpu()[rows]
        e_hard, s_hard = stats(got[:2], ref[:2])
        e_easy, s_easy = stats(got[2:], ref[2:])
        i16 = (ref.numpy() * 32768.0).astype("int16").astype(np.int32)
        d16 = np.abs(w16.cpu().numpy()[rows].astype(np.int32) - i16)
        print(f"C2-size vocoder [{precision}, {'universal' if p else 'synthetic'} weights]: log-mel-like rows max-abs {e_easy:.2e} "
              f"SNR {s_easy:.1f} dB int16 max diff {int(d16[2:].max())} LSB | synthetic-acoustic rows max-abs {e_hard:.2e} "
              f"SNR {s_hard:.1f} dB (ref rms {float(ref[:2].pow(2).mean().sqrt()):.3f}, |ref| max {float(ref[:2].abs().max()):.3f})")
        if precision == "fp32":
            assert e_easy <= 2e-5 and e_hard <= 2e-4, (e_easy, e_hard)
        else:
            assert e_easy <= 6e-3 and s_easy >= 48.0, (e_easy, s_easy)      # measured 4.2e-3 at 49.5 dB (real weights, L = 810)
            assert s_hard >= 48.0, s_hard
        del voc
