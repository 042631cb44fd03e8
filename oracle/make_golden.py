"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) on CPU in the build container.

    python oracle/make_golden.py            # rewrites tests/golden/*.pt

The reference ships no tests or known-answer vectors, so these fixtures ARE the pin:
they hold the reference's own outputs at every stage boundary of the hot path for seeded synthetic
inputs.  The GPU box has no /root/reference; there the tests compare the oracle restatement and the
CUDA path against these files.  Inputs are regenerated from seeds (cmtts_b200/synthetic.py); a
sha256 digest of the regenerated weights is stored so a drifting RNG stream fails loudly instead
of silently comparing different models.
"""
from __future__ import annotations

import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cmtts_b200 import synthetic  # noqa: E402
from cmtts_b200.config import HifiGanSpec, ModelSpec  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (dataset, batch, src_lo, src_hi, weight seed, batch seed, noise seed)
ACOUSTIC_CASES = [
    ("LJSpeech", 3, 9, 14, 0, 1234, 1),
    ("VCTK", 4, 6, 17, 0, 1234, 1),
    ("LibriTTS", 2, 10, 12, 0, 1234, 1),
    # ragged: one single-phoneme utterance next to long ones
    ("LJSpeech", 3, 1, 21, 3, 77, 5),
]


def cliff_margins(spec, ref):
    """Distance of every quantiser input from its nearest decision boundary (the quantisers are
    discontinuous: a value sitting on a boundary flips under 1-ulp differences, which says nothing
    about either implementation).  Fixtures are drawn so that all margins are comfortably large."""
    import numpy as np

    bins = torch.linspace(spec.energy_min, spec.energy_max, spec.energy_bins - 1)
    e = (ref["e_predictions"][..., None] - bins).abs().min(-1).values.min().item()
    keep = ~ref["src_masks"]
    dur_in = (torch.exp(ref["log_d_predictions"]) - 1)[keep]
    d = ((dur_in - torch.floor(dur_in)) - 0.5).abs().min().item()
    f0 = ref["p_predictions"]["f0_denorm"]
    mel = 1127 * (1 + f0 / 700).log()
    mn, mx = 1127 * np.log(1 + 50.0 / 700), 1127 * np.log(1 + 1100.0 / 700)
    mel = torch.where(mel > 0, (mel - mn) * 254 / (mx - mn) + 1, mel)
    mel = mel.clamp(1, 255) + 0.5
    pm = mel[f0 > 0]
    p = (pm - torch.round(pm)).abs().min().item() if pm.numel() else 1.0
    uv = ref["p_predictions"]["cwt"][..., -1].abs().min().item() if spec.use_uv else 1.0
    return {"energy": e, "duration": d, "pitch_bins": p, "uv_logit": uv}


def margins_ok(m):
    return m["energy"] > 2e-4 and m["duration"] > 2e-3 and m["pitch_bins"] > 2e-3 and m["uv_logit"] > 2e-4


def acoustic_case(ds, B, lo, hi, wseed, bseed, nseed):
    from model.cm_tool.karras_diffusion import karras_sample_tts

    spec = ModelSpec.preset(ds)
    sd = synthetic.make_acoustic_state_dict(spec, wseed)
    model, diffusion, _ = ref_shim.build_reference_model(ds, spec.energy_min, spec.energy_max)
    model.load_state_dict(sd)
    model.eval()
    dp0, _ = model.get_segmentation_model()
    for bseed in range(bseed, bseed + 50):
        batch = synthetic.make_batch(spec, B, lo, hi, seed=bseed)
        if lo == 1:
            batch["src_lens"][1] = 1
            batch["texts"][1, 1:] = 0
        with torch.no_grad():
            probe = dp0(speakers=batch["speakers"], texts=batch["texts"], src_lens=batch["src_lens"],
                        spker_embeds=batch["spker_embeds"])
        margins = cliff_margins(spec, probe)
        if margins_ok(margins):
            break
        print(f"  batch seed {bseed}: quantiser input on a decision boundary {margins}, drawing another batch")
    else:
        raise RuntimeError("no cliff-free batch found")
    kw = dict(speakers=batch["speakers"], texts=batch["texts"], src_lens=batch["src_lens"],
              spker_embeds=batch["spker_embeds"])
    cap = {}
    h1 = model.duration_pitch_energy_net.text_encoder.register_forward_hook(
        lambda m, i, o: cap.setdefault("enc", o.detach().clone()))
    dp, _ = model.get_segmentation_model()
    with torch.no_grad():
        ref = dp(**kw)
    h1.remove()
    out = {
        "meta": dict(dataset=ds, batch=B, src_lo=lo, src_hi=hi, weight_seed=wseed, batch_seed=bseed,
                     noise_seed=nseed, digest=synthetic.state_dict_digest(sd),
                     torch=str(torch.__version__), single_phoneme_row=(1 if lo == 1 else -1),
                     quantiser_margins=margins),
        "texts": batch["texts"], "src_lens": batch["src_lens"], "spker_embeds": batch["spker_embeds"],
        "enc": cap["enc"], "log_d": ref["log_d_predictions"], "e_pred": ref["e_predictions"],
        "d_rounded": ref["d_rounded"], "mel_lens": ref["mel_lens"], "cond": ref["cond"],
        "cwt": ref["p_predictions"]["cwt"], "f0_denorm": ref["p_predictions"]["f0_denorm"],
        "f0_mean": ref["p_predictions"]["f0_mean"], "f0_std": ref["p_predictions"]["f0_std"],
        "speaker_emb": ref["speaker_emb"],
    }
    Bn, L, _ = ref["cond"].shape
    from oracle.cmtts_oracle import sampler_plan

    for T in (1, 2, 4):
        gen = ref_shim.ReplayGenerator(nseed)
        sampler, steps, ts = sampler_plan(T)
        extra = {} if T == 1 else dict(steps=steps, ts=ts)
        outs = []
        h2 = model.net.register_forward_hook(lambda m, i, o: outs.append(o.detach().clone()))
        with torch.no_grad():
            mel = karras_sample_tts(diffusion=diffusion, model=model, shape=(Bn, 1, L, spec.n_mels),
                                    model_kwargs=kw, device="cpu", sigma_max=spec.sigma_max,
                                    sigma_min=spec.sigma_min, sampler=sampler, generator=gen, **extra)
        h2.remove()
        out[f"mel_T{T}"] = mel
        # Denoiser output of the first evaluation, (B,1,M,L) as the module returns it
        out[f"model_output0_T{T}"] = outs[0]
        out[f"n_noise_T{T}"] = len(gen.drawn)
    return out


def hifigan_case(tag, state_dict, B, L, mel_seed, extra_meta):
    voc = ref_shim.build_reference_vocoder(state_dict)
    mel = synthetic.make_mels(B, 80, L, seed=mel_seed)
    with torch.no_grad():
        wav = voc(mel)
    i16 = (wav.squeeze(1).numpy() * 32768.0).astype("int16")  # utils/model.py:195-198
    meta = dict(tag=tag, batch=B, frames=L, mel_seed=mel_seed, torch=str(torch.__version__))
    meta.update(extra_meta)
    return {"meta": meta, "wav": wav, "int16": torch.from_numpy(i16)}


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref_shim.install()
    for i, case in enumerate(ACOUSTIC_CASES):
        o = acoustic_case(*case)
        f = os.path.join(GOLDEN, f"acoustic_{i}_{case[0]}.pt")
        torch.save(o, f)
        print(f, os.path.getsize(f), "L", o["cond"].shape[1], "mel_lens", o["mel_lens"].tolist())
    ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=7)
    o = hifigan_case("synthetic", ck["generator"], 2, 24, 99,
                     dict(weight_seed=7, digest=synthetic.state_dict_digest(ck["generator"])))
    torch.save(o, os.path.join(GOLDEN, "hifigan_synthetic.pt"))
    print("hifigan_synthetic", float(o["wav"].abs().max()))
    for spk in ("universal",):
        p = os.path.join(ref_shim.REFERENCE_ROOT, "hifigan", f"generator_{spk}.pth.tar")
        sd = torch.load(p, map_location="cpu", weights_only=True)["generator"]
        o = hifigan_case(spk, sd, 1, 32, 99, dict(weights_sha256=sha256_file(p)))
        torch.save(o, os.path.join(GOLDEN, f"hifigan_{spk}.pt"))
        print("hifigan", spk, float(o["wav"].abs().max()))


if __name__ == "__main__":
    main()
