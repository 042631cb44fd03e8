"""Consistency-model sampler with the reference's call protocol:

    karras_sample_tts(diffusion, model, shape, steps=2, ..., sampler="onestep"|"multistep",
                      generator=None, ts=None) -> (B, L, 80)       karras_diffusion.py:480-577

`generator` is the noise-injection seam (any object with randn / randn_like), kept bit-for-bit.
When `model` is a cmtts_b200 CMTotalTTS the solver runs the fused device path: encoder + variance
adaptor once per call, the sigma-only step embedding once per call, and c_in / c_out / c_skip
folded into the denoiser's first and last GEMM epilogues (results identical to the reference's
schedule, which recomputes all of that T+1 times — SURVEY.md §0.4, App. C).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops  # noqa: F401  (registers torch.ops.cmtts_b200.*)
from .model import CMTotalTTS, KarrasDenoiser


class DummyGenerator:
    """model/cm_tool/random_util.py:17-25 — global torch RNG on the target device."""

    def randn(self, *args, **kwargs):
        return torch.randn(*args, **kwargs)

    def randint(self, *args, **kwargs):
        return torch.randint(*args, **kwargs)

    def randn_like(self, *args, **kwargs):
        return torch.randn_like(*args, **kwargs)


def get_generator(generator, num_samples=0, seed=0):
    if generator == "dummy":
        return DummyGenerator()
    raise NotImplementedError("only the 'dummy' generator is on the inference path (random_util.py:7-15)")


def append_zero(x):
    return torch.cat([x, x.new_zeros([1])])


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0, device="cpu"):
    """karras_diffusion.py:580-586 (fp32 tensor ops on the host, like the reference)."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return append_zero(sigmas).to(device)


def sampler_plan(T: int):
    """synthesize.py:106-146: T -> (sampler, steps, ts)."""
    if T == 1:
        return "onestep", 2, None
    if T == 2:
        return "multistep", 2, (0, 0, 1)
    if T == 4:
        return "multistep", 2, (0, 0, 0, 0, 1)
    raise ValueError(f"T must be 1, 2 or 4 (synthesize.py:106-146), got {T}")


def evaluation_sigma(sampler: str, steps: int, sigma_min: float, sigma_max: float, rho: float) -> float:
    """The sigma every distiller evaluation of one karras_sample_tts call runs at: sigmas[0] of the fp32
    Karras grid for `onestep` (karras_diffusion.py:800-811), the Python-float t of ts = 0 for `multistep` (:829-854)."""
    if sampler == "onestep":
        return float(get_sigmas_karras(steps, sigma_min, sigma_max, rho, device="cpu")[0])
    t_max_rho, t_min_rho = sigma_max ** (1 / rho), sigma_min ** (1 / rho)
    return (t_max_rho + 0 / (steps - 1) * (t_min_rho - t_max_rho)) ** rho


def rescaled_timesteps(batch: int, sigma_value: float) -> torch.Tensor:
    """`1000 * 0.25 * log(sigma + 1e-44)` as the reference computes it: an fp32 (B,) HOST tensor through torch.log
    (karras_diffusion.py:404)."""
    sig = torch.full((batch,), sigma_value, dtype=torch.float64).to(torch.float32)
    return 1000 * 0.25 * torch.log(sig + 1e-44)


class _FusedDistiller:
    """denoiser(x_t, sigma) of karras_sample_tts (karras_diffusion.py:560-566) on the device path."""

    def __init__(self, diffusion: KarrasDenoiser, model: CMTotalTTS, model_kwargs: dict, cond: Optional[dict],
                 prepared_steps=None):
        self.diffusion, self.model = diffusion, model
        self.kw = model_kwargs
        self.cond = cond
        self._steps_key = None
        self._steps = None
        self._prepared = prepared_steps      # (ds_all, dsp_all) made ahead by the caller for this call's sigma
        self._cond_proj = None       # (cond tensor, its per-layer projections): made once per sampler call
        self.model_outputs = None  # set to a list to capture F per evaluation (parity tests)

    def conditioner(self, L: int) -> dict:
        if self.cond is None or self.cond["cond"].shape[1] != L:
            kw = self.kw
            self.cond = self.model.dpen(kw["texts"], kw["src_lens"], self.model.speaker_input(kw.get("speakers"), kw.get("spker_embeds")), L,
                                        kw.get("p_control", 1.0), kw.get("e_control", 1.0), kw.get("d_control", 1.0))
        return self.cond

    def __call__(self, x_t: torch.Tensor, sigma_value: float) -> torch.Tensor:
        B, _, L, _ = x_t.shape
        cond = self.conditioner(L)
        c_skip, c_out, c_in, _ = self.diffusion.scalar_plan(sigma_value)
        if self._prepared is not None:
            self._steps = self._prepared
        elif self._steps_key != sigma_value:
            self._steps = self.model.prepare_steps(rescaled_timesteps(B, sigma_value), cond["speaker_emb"])
            self._steps_key = sigma_value
        cp = None
        if self.model.precision == "tc":
            if self._cond_proj is None or self._cond_proj[0] is not cond["cond"]:
                self._cond_proj = (cond["cond"], self.model.project_cond(cond["cond"]))
            cp = self._cond_proj[1]
        if self.model_outputs is not None:
            out, mo = self.model.denoise_step(x_t, cond["cond"], self._steps, c_in, c_out, c_skip, want_model_out=True,
                                              cond_proj=cp)
            self.model_outputs.append(mo)
            return out
        return self.model.denoise_step(x_t, cond["cond"], self._steps, c_in, c_out, c_skip, cond_proj=cp)


def karras_sample_tts(diffusion, model, shape, steps=2, clip_denoised=False, progress=False, callback=None,
                      model_kwargs=None, device=None, sigma_min=0.002, sigma_max=80, rho=7.0,
                      sampler="onestep", s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0,
                      generator=None, ts=None, T=None, cond_dict: Optional[dict] = None, trace: Optional[dict] = None,
                      prepared_steps=None):
    """Drop-in for karras_diffusion.py:480-577 (samplers 'onestep' and 'multistep', the two
    synthesize.py uses).  Extra keyword `cond_dict`: the pre-pass output of
    duration_pitch_energy_net, reused instead of recomputing the conditioner; `prepared_steps`: the output of
    `model.prepare_steps` for this call's evaluation sigma (`evaluation_sigma`), made ahead by a caller that keeps host
    to device copies out of the solver loop (the CUDA-graph path of cmtts_b200.synthesize.Pipeline)."""
    if generator is None:
        generator = get_generator("dummy")
    if not isinstance(model, CMTotalTTS):
        raise TypeError("karras_sample_tts: model must be a cmtts_b200 CMTotalTTS")
    if sampler not in ("onestep", "multistep"):
        raise NotImplementedError(f"sampler {sampler!r}: only 'onestep' and 'multistep' are on the hot path")
    device = model.device if device is None else torch.device(device)
    model_kwargs = model_kwargs or {}
    sigmas = get_sigmas_karras(steps, sigma_min, sigma_max, rho, device="cpu")
    x_T = generator.randn(*shape, device=device) * sigma_max
    x_T = x_T.to(device=device, dtype=torch.float32)
    distiller = _FusedDistiller(diffusion, model, model_kwargs, cond_dict, prepared_steps)
    if trace is not None:
        distiller.model_outputs = trace.setdefault("model_output", [])
    if sampler == "onestep":
        # sample_onestep, karras_diffusion.py:800-811
        x_0 = distiller(x_T, float(sigmas[0]))
    else:
        # stochastic_iterative_sampler, karras_diffusion.py:829-854 (Python-float sigma arithmetic)
        t_min, t_max, rho_ = sigma_min, sigma_max, diffusion.rho
        t_max_rho = t_max ** (1 / rho_)
        t_min_rho = t_min ** (1 / rho_)
        x = x_T
        for i in range(len(ts) - 1):
            t = (t_max_rho + ts[i] / (steps - 1) * (t_min_rho - t_max_rho)) ** rho_
            x0 = distiller(x, t)
            next_t = (t_max_rho + ts[i + 1] / (steps - 1) * (t_min_rho - t_max_rho)) ** rho_
            next_t = np.clip(next_t, t_min, t_max)
            noise = generator.randn_like(x).to(device=device, dtype=torch.float32).contiguous()
            # x = x0 + noise * np.sqrt(next_t**2 - t_min**2) * 0.85: two fp32 multiplies then an add
            s1 = float(np.float32(np.sqrt(next_t ** 2 - t_min ** 2)))
            s2 = float(np.float32(0.85))
            x = torch.ops.cmtts_b200.renoise(x0, noise, s1, s2)
        x_0 = x
    return x_0[:, 0]
