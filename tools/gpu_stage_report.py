"""Stage-by-stage error report of the CUDA path against the oracle (diagnostics for gpurun)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmtts_b200 import synthetic  # noqa: E402
from cmtts_b200.config import HifiGanSpec, ModelSpec  # noqa: E402
from cmtts_b200.model import CMTotalTTS, KarrasDenoiser  # noqa: E402
from cmtts_b200.sampler import karras_sample_tts, sampler_plan  # noqa: E402
from cmtts_b200.vocoder import Generator  # noqa: E402
from oracle import cmtts_oracle as O  # noqa: E402

DEV = "cuda:0"


def err(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "tc"
    print("precision", prec)
    for ds, B, lo, hi in [("LJSpeech", 3, 20, 45), ("VCTK", 4, 10, 40), ("LibriTTS", 8, 60, 115)]:
        spec = ModelSpec.preset(ds)
        sd = synthetic.make_acoustic_state_dict(spec, 0)
        batch = synthetic.make_batch(spec, B, lo, hi, seed=1234)
        W = O.Weights(sd)
        with torch.no_grad():
            ref = O.dpen(W, spec, **batch)
        m = CMTotalTTS(spec=spec, precision=prec).load_state_dict(sd).to(DEV)
        m.tc_frontend = os.environ.get("CMTTS_TC_FRONTEND", "0") == "1"
        out = m.dpen(batch["texts"], batch["src_lens"], batch["spker_embeds"])
        torch.cuda.synchronize()
        print(f"== {ds} B={B} L={ref['cond'].shape[1]}")
        for k_mine, k_ref in [("enc", "enc"), ("log_d_predictions", "log_d_predictions"), ("e_predictions", "e_predictions"),
                              ("d_rounded", "d_rounded"), ("mel_lens", "mel_lens")]:
            print(f"  {k_mine:20s} {err(out[k_mine], ref[k_ref]):.3e}")
        if out["cond"].shape == ref["cond"].shape:
            print(f"  cwt                  {err(out['p_predictions']['cwt'], ref['cwt']):.3e}")
            print(f"  f0_denorm            {err(out['p_predictions']['f0_denorm'], ref['f0_denorm']):.3e}")
            print(f"  e_idx flips          {int((out['e_idx'].cpu() != ref['e_idx']).sum())}")
            print(f"  pitch flips          {int((out['pitch_idx'].cpu() != ref['pitch_idx']).sum())}")
            print(f"  mel2ph eq            {bool(torch.equal(out['mel2ph'].cpu(), ref['mel2ph']))}")
            print(f"  cond                 {err(out['cond'], ref['cond']):.3e}")
        else:
            print("  cond shape mismatch", out["cond"].shape, ref["cond"].shape)
            continue
        L = ref["cond"].shape[1]
        for T in (1, 4):
            g = torch.Generator().manual_seed(1)
            noise = [torch.randn(B, 1, L, 80, generator=g) for _ in range(1 if T == 1 else T + 1)]
            it = iter(noise)
            tr = {}
            with torch.no_grad():
                rm, _ = O.sample(W, spec, batch, T, lambda s: next(it), trace=tr)
            it2 = iter(noise)

            class G:
                def randn(self, *s, device=None, **k): return next(it2).to(device)
                def randn_like(self, x): return next(it2).to(x.device)
            sampler, steps, ts = sampler_plan(T)
            tr2 = {}
            mel = karras_sample_tts(KarrasDenoiser(distillation=True), m, (B, 1, L, 80), steps=steps, model_kwargs=batch,
                                    device=DEV, sampler=sampler, ts=ts, generator=G(), cond_dict=out, trace=tr2)
            torch.cuda.synchronize()
            print(f"  T={T} model_out0 {err(tr2['model_output'][0], tr['model_output'][0]):.3e}  mel {err(mel, rm):.3e}  |mel|max {float(rm.abs().max()):.2f}")
    ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=7)
    Wf = O.Weights(synthetic.fold_weight_norm(ck["generator"]))
    voc = Generator(hspec=HifiGanSpec(), precision=prec).load_state_dict(ck["generator"]).to(DEV)
    for (B, L) in [(2, 24), (2, 150)]:
        mel = synthetic.make_mels(B, 80, L, seed=99)
        with torch.no_grad():
            ref = O.hifigan(Wf, HifiGanSpec(), mel)
        wav = voc(mel.to(DEV))
        torch.cuda.synchronize()
        d = (wav.cpu() - ref)
        snr = 10 * torch.log10(ref.pow(2).mean() / d.pow(2).mean())
        print(f"== hifigan synthetic (B={B}, L={L}): wav max err {err(wav, ref):.3e} rms err {float(d.pow(2).mean().sqrt()):.3e} "
              f"|wav|max {float(ref.abs().max()):.3f} SNR {float(snr):.1f} dB")
    pass
    real = os.path.join(ROOT, "oracle", "_ref", "hifigan", "generator_universal.pth.tar")
    if os.path.isfile(real):
        sd = torch.load(real, map_location="cpu", weights_only=True)["generator"]
        Wr = O.Weights(synthetic.fold_weight_norm(sd))
        voc = Generator(hspec=HifiGanSpec(), precision=prec).load_state_dict(sd).to(DEV)
        mel = synthetic.make_mels(1, 80, 100, seed=3)
        with torch.no_grad():
            ref = O.hifigan(Wr, HifiGanSpec(), mel)
        wav = voc(mel.to(DEV))
        torch.cuda.synchronize()
        d = (wav.cpu() - ref)
        snr = 10 * torch.log10(ref.pow(2).mean() / d.pow(2).mean())
        i16 = (wav.cpu() * 32768).to(torch.int32) - (ref * 32768).to(torch.int32)
        print(f"== hifigan REAL universal weights: wav max err {err(wav, ref):.3e} rms {float(d.pow(2).mean().sqrt()):.3e} "
              f"|wav|max {float(ref.abs().max()):.3f} SNR {float(snr):.1f} dB int16 maxdiff {int(i16.abs().max())}")


if __name__ == "__main__":
    main()
