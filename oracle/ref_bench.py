"""CPU timing of the UNMODIFIED reference's hot path (test / measurement infrastructure, part of the ORACLE side:
only bench.py's `--impl reference` arm and its `cpu_baseline` leg call it; the product package never does).

What runs is the reference's own code, imported through oracle/ref_shim.py from /root/reference (build container) or
from the copy `__graft_entry__.build()` stages under oracle/_ref/ (GPU box): `create_model_and_diffusion_tts`
(script_util.py:56-75), `DurationPitchSpeakerNet.forward` pre-pass, `karras_sample_tts` (karras_diffusion.py:480-577,
which re-runs encoder + variance adaptor inside every solver step through `CMTotalTTS.forward`), `get_vocoder` +
`vocoder_infer` (utils/model.py:155-205) — the sequence of p_rtf_cm.py:174-226, on all host threads, fp32, no_grad.
If neither tree is present the oracle port (oracle/cmtts_oracle.py, literal schedule) is timed instead and the result
says `kind: "port"`.
"""
from __future__ import annotations

import contextlib
import os
import time
from typing import Dict, Optional

import torch

from . import ref_shim


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


class ReferenceRunner:
    """Builds the reference model + vocoder once; `step(batch, T)` runs one whole pass and returns timings."""

    def __init__(self, dataset: str, spec, acoustic_sd: Dict[str, torch.Tensor], hifigan_sd: Optional[Dict] = None,
                 force_port: bool = False):
        self.kind = "reference" if (ref_shim.reference_available() and not force_port) else "port"
        self.spec = spec
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        if self.kind == "reference":
            ref_shim.install()
            self.model, self.diffusion, cfgs = ref_shim.build_reference_model(dataset, spec.energy_min, spec.energy_max)
            self.preprocess_config, self.model_config, self.train_config = cfgs
            self.model.load_state_dict(acoustic_sd)
            self.model.eval()
            if hifigan_sd is None:
                import utils.model as um            # the reference's own loader, cwd-relative paths (utils/model.py:171-178)

                # BASELINE.json C2 names the universal HiFi-GAN; LJSpeech's model.yaml asks for generator_LJSpeech,
                # which the reference ships only as a zip
                self.model_config["vocoder"]["speaker"] = "universal"
                with _cwd(ref_shim.REFERENCE_ROOT):
                    self.vocoder = um.get_vocoder(self.model_config, torch.device("cpu"))
            else:
                self.vocoder = ref_shim.build_reference_vocoder(hifigan_sd)
        else:
            from cmtts_b200 import synthetic
            from . import cmtts_oracle as O

            if hifigan_sd is None:
                raise RuntimeError("the oracle port needs an explicit HiFi-GAN state_dict")
            self.O = O
            self.W = O.Weights(acoustic_sd)
            self.Wf = O.Weights(synthetic.fold_weight_norm(hifigan_sd))

    def step(self, batch: Dict[str, Optional[torch.Tensor]], T: int, seed: int = 1) -> Dict[str, float]:
        g = torch.Generator().manual_seed(seed)
        spec = self.spec
        if self.kind == "port":
            O = self.O
            t0 = time.perf_counter()
            with torch.no_grad():
                mel, wav, i16, pre = O.synthesize(self.W, self.Wf, spec, batch, T, lambda s: torch.randn(*s, generator=g),
                                                  literal=True)
            dt = time.perf_counter() - t0
            return {"seconds": dt, "seconds_after_prepass": dt, "valid_frames": int(pre["mel_lens"].sum()),
                    "first_utt_seconds_audio": int(pre["mel_lens"][0]) * spec.hop_length / spec.sampling_rate,
                    "padded_frames": int(mel.shape[0] * mel.shape[1])}
        from model.cm_tool.karras_diffusion import karras_sample_tts
        import utils.model as um

        class Gen:                                   # `generator` seam, karras_diffusion.py:498
            def randn(self, *shape, device=None, **_):
                return torch.randn(*shape, generator=g)

            def randn_like(self, x):
                return torch.randn(*x.shape, generator=g)

        kw = dict(speakers=batch["speakers"], texts=batch["texts"], src_lens=batch["src_lens"],
                  spker_embeds=batch.get("spker_embeds"))
        sampler, steps, ts = {1: ("onestep", 2, None), 2: ("multistep", 2, (0, 0, 1)), 4: ("multistep", 2, (0, 0, 0, 0, 1))}[T]
        extra = {} if T == 1 else dict(steps=steps, ts=ts)          # synthesize.py:106-146
        dp, _ = self.model.get_segmentation_model()
        t0 = time.perf_counter()
        with torch.no_grad():
            out_dict = dp(**kw)                                       # pre-pass, synthesize.py:102
            t1 = time.perf_counter()                                  # p_rtf_cm.py:192 starts its Timer here
            B, L, _ = out_dict["cond"].shape
            mel = karras_sample_tts(diffusion=self.diffusion, model=self.model, shape=(B, 1, L, spec.n_mels),
                                    model_kwargs=kw, device="cpu", sigma_max=spec.sigma_max, sigma_min=spec.sigma_min,
                                    sampler=sampler, generator=Gen(), **extra)
            lengths = out_dict["mel_lens"] * spec.hop_length
            wavs = um.vocoder_infer(mel.transpose(1, 2), self.vocoder, self.model_config, self.preprocess_config,
                                    lengths=lengths)
        t2 = time.perf_counter()
        assert len(wavs) == B and wavs[0].dtype.name == "int16"
        return {"seconds": t2 - t0, "seconds_after_prepass": t2 - t1, "valid_frames": int(out_dict["mel_lens"].sum()),
                "first_utt_seconds_audio": int(out_dict["mel_lens"][0]) * spec.hop_length / spec.sampling_rate,
                "padded_frames": int(B * L)}

    def vocoder_step(self, mel_bml: torch.Tensor) -> Dict[str, float]:
        """C5: HiFi-GAN only, (B, 80, L) -> int16 (utils/model.py:187-205)."""
        t0 = time.perf_counter()
        with torch.no_grad():
            if self.kind == "reference":
                import utils.model as um
                wavs = um.vocoder_infer(mel_bml, self.vocoder, self.model_config, self.preprocess_config)
            else:
                wavs = self.O.wav_to_int16(self.O.hifigan(self.Wf, self.spec.hifigan, mel_bml), None)
        dt = time.perf_counter() - t0
        return {"seconds": dt, "valid_frames": int(mel_bml.shape[0] * mel_bml.shape[2]), "n": len(wavs)}
