"""Host-side mirror of the reference's acoustic-model seam.

`CMTotalTTS` keeps the call protocol of model/cm_tool/tts_net.py:40-183:

    model, diffusion = create_model_and_diffusion_tts(use_fp16, weight_schedule, tts_model_config, ...)
    model.load_state_dict(torch.load(".../CMDenoiserTTS/model000000.pt")); model.to(dev); model.eval()
    dpen, denoise_fun = model.get_segmentation_model()
    out_dict = dpen(speakers=, texts=, src_lens=, spker_embeds=)          # cmtts.py:44-122
    y = model(x[B,1,L,80], timesteps[B], speakers=, texts=, src_lens=, spker_embeds=)   # tts_net.py:75

but every tensor op runs in libcmtts_b200.so (hand-written sm_100a CUDA) through the C ABI.
PyTorch is used for device memory, streams and the module-like surface only.  There is no CPU
path: tensors must live on a CUDA device and the library must be built.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from . import ops  # noqa: F401  (registers torch.ops.cmtts_b200.*)
from .config import ModelSpec
from .weights import PackedAcoustic


class _Workspace:
    """Grow-only scratch buffers in HBM, one per stage, reused across calls on the same stream."""

    def __init__(self, device):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}

    def get(self, name: str, nbytes: int) -> torch.Tensor:
        b = self.bufs.get(name)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self.bufs[name] = b
        return b


class DurationPitchSpeakerNet:
    """`duration_pitch_energy_net` callable: FastSpeech2 FFT encoder + variance adaptor
    (model/cmtts.py:44-122, inference branch)."""

    def __init__(self, owner: "CMTotalTTS"):
        self.owner = owner

    def __call__(self, speakers=None, texts=None, src_lens=None, mels=None, mel_lens=None,
                 p_targets=None, e_targets=None, d_targets=None, mel2phs=None, spker_embeds=None,
                 p_control=1.0, e_control=1.0, d_control=1.0) -> dict:
        if p_targets is not None or e_targets is not None or d_targets is not None:
            raise NotImplementedError("teacher-forced targets are a training path (out of scope)")
        max_mel_len = None if mels is None else int(mels.shape[2])  # cmtts.py:60-63
        spker_embeds = self.owner.speaker_input(speakers, spker_embeds)
        return self.owner.dpen(texts, src_lens, spker_embeds, max_mel_len, float(p_control),
                               float(e_control), float(d_control))


class CMTotalTTS:
    """B200-native stand-in for model/cm_tool/tts_net.py:CMTotalTTS (inference only)."""

    #: "tc"   — residual stack on tcgen05 tensor cores with fp16 hi/lo operand pairs (fp32-class);
    #: "fp32" — everything on the fp32 FFMA kernels (the device-side yardstick).
    PRECISIONS = ("tc", "fp32")

    def __init__(self, use_fp16=False, args=None, preprocess_config=None, model_config=None,
                 train_config=None, spec: Optional[ModelSpec] = None, precision: str = "tc", **_ignored):
        if use_fp16:
            raise NotImplementedError("use_fp16 converts the torso for training; inference here is fp32-class")
        if spec is None:
            spec = ModelSpec.from_reference_configs(preprocess_config, model_config, train_config)
        if precision not in self.PRECISIONS:
            raise ValueError(f"precision must be one of {self.PRECISIONS}")
        self.precision = precision
        # Encoder + variance-adaptor GEMMs on the hi/lo tensor-core kernel when precision == "tc".  These
        # stages feed the duration / energy / pitch quantisers, so accuracy matters: tcgen05 accumulates
        # in fp32 with truncation (error grows with the number of accumulation steps), which is why the
        # kernel keeps the hi/lo cross terms in their own accumulator.  Measured against the reference:
        # log_d 1.4-1.9e-6 (FFMA 1.1-1.2e-6), energy 1.0e-5 (6e-6), cwt 6.9e-6 (5.2e-6).  Set False to
        # run them on the fp32 FFMA kernels (4 ms slower per C2 batch).
        self.tc_frontend = True
        self.spec = spec
        self.device = torch.device("cpu")
        self._sd: Optional[Dict[str, torch.Tensor]] = None
        self.packed: Optional[PackedAcoustic] = None
        self.training = False
        self.duration_pitch_energy_net = DurationPitchSpeakerNet(self)
        self._ws: Optional[_Workspace] = None
        self.handle: Optional[int] = None        # names this model in the torch.ops.cmtts_b200 custom ops (cmtts_b200/ops.py)
        self.lib = _lib.load()

    # ---- nn.Module-like surface used by synthesize.py:79-86 ---------------------------------
    def load_state_dict(self, state_dict, strict: bool = True):
        self._sd = {k: v.detach().to("cpu") for k, v in state_dict.items()}
        if self.device.type == "cuda":
            self._repack()
        return self

    def state_dict(self):
        return dict(self._sd) if self._sd is not None else {}

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.CmttsError("cmtts_b200 runs on CUDA devices only (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        if self._sd is not None:
            self._repack()
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("training is out of scope for the B200 inference path")
        return self

    def parameters(self):
        return iter(())

    def _repack(self):
        from . import ops
        self.packed = PackedAcoustic(self.spec, self._sd, self.device)
        self._dims = self.packed.dims()
        self._ws = _Workspace(self.device)
        if self.handle is None:
            self.handle = ops.register(self)

    def _ready(self):
        if self.packed is None:
            raise _lib.CmttsError("CMTotalTTS: call load_state_dict(...) and .to('cuda') first")

    # ---- encoder + variance adaptor --------------------------------------------------------
    def dpen_head(self, texts: torch.Tensor, src_lens: torch.Tensor, spker_embeds: Optional[torch.Tensor],
                  e_control=1.0, d_control=1.0) -> dict:
        """Token-rate half of dpen: encoder + the token-rate variance adaptor, everything BEFORE the host learns the
        output length.  Inputs must already be contiguous CUDA tensors (int64 / fp32).  No host sync, no host-side
        branching on device data: this piece is what a CUDA graph can hold (cmtts_b200/synthesize.py)."""
        s = self.spec
        ops = torch.ops.cmtts_b200
        B, T = texts.shape
        bad_tok = ((texts < 0) | (texts >= s.vocab)).any() if texts.numel() else None   # nn.Embedding would raise IndexError
        enc = ops.encoder_forward(self.handle, texts, src_lens)
        out1, log_d, d_rounded, e_pred, e_idx, cumsum, mel_lens, spk, f0_stats = ops.variance_token(
            self.handle, enc, src_lens, spker_embeds if s.multi_speaker else None, float(e_control), float(d_control))
        m = mel_lens.max() if B > 0 else torch.zeros((), dtype=torch.int64, device=self.device)
        return {"texts": texts, "src_lens": src_lens, "enc": enc, "out1": out1, "log_d": log_d, "d_rounded": d_rounded,
                "e_pred": e_pred, "e_idx": e_idx, "cumsum": cumsum, "mel_lens": mel_lens,
                "spk": spk if s.multi_speaker else None, "f0_stats": f0_stats, "max_len": m, "bad_tok": bad_tok}

    def dpen_tail(self, head: dict, L: int, local_max: int, p_control=1.0, extra=()) -> dict:
        """Frame-rate half of dpen for a known padded length L (length regulator, CWT pitch path, conditioner)."""
        ops = torch.ops.cmtts_b200
        dev = self.device
        src_lens, mel_lens, f0_stats = head["src_lens"], head["mel_lens"], head["f0_stats"]
        T = head["texts"].shape[1]
        cond, mel2ph, cwt, f0_denorm, pitch_idx = ops.variance_frame(self.handle, head["out1"], head["cumsum"], mel_lens,
                                                                      f0_stats, float(p_control), int(L))
        ar_t = torch.arange(T, device=dev)
        ar_l = torch.arange(local_max, device=dev)
        return {
            "cond": cond,
            "p_targets": None,
            "p_predictions": {"pitch_pred": None, "f0_denorm": f0_denorm, "cwt": cwt,
                              "f0_mean": f0_stats[:, 0], "f0_std": f0_stats[:, 1]},
            "e_predictions": head["e_pred"],
            "log_d_predictions": head["log_d"],
            "d_rounded": head["d_rounded"],
            "mel_lens": mel_lens,
            "mel_masks": ar_l[None, :] >= mel_lens[:, None],       # get_mask_from_lengths(mel_len)
            "src_masks": ar_t[None, :] >= src_lens[:, None],
            "speaker_emb": head["spk"],
            "src_lens": src_lens,
            # extras (not in the reference dict)
            "enc": head["enc"], "mel2ph": mel2ph, "e_idx": head["e_idx"], "pitch_idx": pitch_idx, "l_max_extra": list(extra),
        }

    def read_lengths(self, head: dict, l_max_hook=None):
        """THE host round trip of the path: the padded output length (and the bad-token flag) in one read.  The
        multi-GPU hook reduces the device scalar BEFORE it is read (cmtts_b200/dist.py).  -> (local_max, extra)."""
        m = head["max_len"]
        if l_max_hook is not None:
            m = l_max_hook(m)
        bad_tok = head["bad_tok"]
        if isinstance(m, torch.Tensor):
            if bad_tok is not None:
                m = torch.cat([m.reshape(-1).to(torch.int64), bad_tok.reshape(1).to(torch.int64)])
            vals = m.reshape(-1).tolist()
            if bad_tok is not None and vals.pop():
                raise IndexError(f"token id outside [0, {self.spec.vocab}) (wrong symbol table?)")
            return int(vals[0]), [int(v) for v in vals[1:]]
        if bad_tok is not None and bool(bad_tok.item()):
            raise IndexError(f"token id outside [0, {self.spec.vocab}) (wrong symbol table?)")
        return int(m), []

    def speaker_input(self, speakers, spker_embeds):
        """What `dpen` takes as its speaker argument: the external embeddings (cmtts.py:79-81), or — `speaker_embedder: none`,
        cmtts.py:76-78 `self.speaker_emb(speakers)` — one-hot rows of the speaker ids, which the packed table turns into
        the embedding rows exactly.  An id outside the table raises IndexError as nn.Embedding does."""
        n = self.spec.n_speakers
        if not n:
            return spker_embeds
        if speakers is None:
            raise AssertionError("speaker ids are needed (speaker_embedder: none)")
        ids = torch.as_tensor(speakers).to(torch.int64).reshape(-1)
        if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= n):
            raise IndexError(f"speaker id outside [0, {n})")
        onehot = torch.zeros(ids.numel(), self.spec.ext_speaker_dim, dtype=torch.float32, device=ids.device)
        onehot[torch.arange(ids.numel(), device=ids.device), ids] = 1.0
        return onehot

    def prepare_inputs(self, texts, src_lens, spker_embeds):
        """Host-side checks the reference makes, then contiguous device tensors."""
        s, dev = self.spec, self.device
        if s.multi_speaker and spker_embeds is None:
            raise AssertionError("Speaker embedding should not be None")  # cmtts.py:80
        if not texts.is_cuda and texts.numel() and (int(texts.min()) < 0 or int(texts.max()) >= s.vocab):
            # nn.Embedding raises IndexError here (modules.py:145); the gather kernel would read out of bounds
            raise IndexError(f"token id outside [0, {s.vocab}) (wrong symbol table?)")
        texts = texts.to(dev, torch.int64).contiguous()
        src_lens = src_lens.to(dev, torch.int64).contiguous()
        spk_in = None if spker_embeds is None or not s.multi_speaker else spker_embeds.to(dev, torch.float32).contiguous()
        return texts, src_lens, spk_in

    def dpen(self, texts: torch.Tensor, src_lens: torch.Tensor, spker_embeds: Optional[torch.Tensor],
             max_mel_len: Optional[int] = None, p_control=1.0, e_control=1.0, d_control=1.0,
             l_max_hook=None) -> dict:
        """Returns the reference's out_dict (cmtts.py:108-121) plus the integer side products
        (`mel2ph`, `e_idx`, `pitch_idx`).  One host sync: `mel_lens` must reach the host to size
        the frame-rate tensors (the reference syncs B*T times in LengthRegulator.expand).
        `l_max_hook(local_max) -> global_max` lets the multi-GPU driver all-reduce L_max."""
        self._ready()
        texts, src_lens, spk_in = self.prepare_inputs(texts, src_lens, spker_embeds)
        head = self.dpen_head(texts, src_lens, spk_in, e_control, d_control)
        local_max, extra = self.read_lengths(head, l_max_hook)
        if max_mel_len and int(max_mel_len) < local_max:
            # the reference's pad(output, max_len) fails here too (utils/tools.py:724-742: negative F.pad of a
            # longer row); truncating silently would leave mel_lens / mel_masks inconsistent with cond
            raise ValueError(f"dpen: max_mel_len={int(max_mel_len)} is shorter than the predicted length {local_max}")
        L = int(max_mel_len) if max_mel_len else local_max
        return self.dpen_tail(head, L, local_max, p_control, extra)

    def packed_check_rows(self, n: int):
        if n + 2 > self.packed.pe_rows:
            self.packed.ensure_pe_rows(n)
            self._dims = self.packed.dims()

    # ---- denoiser ----------------------------------------------------------------------------
    def prepare_steps(self, timesteps: torch.Tensor, speaker_emb: Optional[torch.Tensor]):
        """Step-embedding MLP and the per-layer diffusion/speaker projections (blocks.py:633-640,
        :669-674): sigma-only work, hoisted out of the solver loop."""
        self._ready()
        t = timesteps.to(self.device, torch.float32).contiguous()
        spk = None if (speaker_emb is None or not self.spec.multi_speaker) else speaker_emb.to(self.device, torch.float32).contiguous()
        return torch.ops.cmtts_b200.denoiser_prepare(self.handle, t, spk)

    def denoise_step(self, x_t: torch.Tensor, cond: torch.Tensor, steps: Tuple[torch.Tensor, torch.Tensor],
                     c_in: float = 1.0, c_out: float = 1.0, c_skip: float = 0.0, want_model_out: bool = False,
                     cond_proj: Optional[torch.Tensor] = None):
        """out = c_out * F(c_in * x_t) + c_skip * x_t on (B,1,L,M) / (B,L,M) channels-last mels.
        `cond_proj`: the conditioner's per-layer projections from `project_cond(cond)` when the caller evaluates several
        solver steps on one conditioner (the sampler does); made here per call otherwise."""
        self._ready()
        lib, s, dev = self.lib, self.spec, self.device
        shape = tuple(x_t.shape)
        lead = 1
        for n in shape[:-2]:
            lead *= int(n)
        x = x_t.to(dev, torch.float32).contiguous().view(lead, shape[-2], shape[-1])   # explicit: B may be 0
        B, L, M = x.shape
        if M != s.n_mels or tuple(cond.shape) != (B, L, s.hidden):
            raise ValueError(f"denoise_step: x {tuple(shape)} / cond {tuple(cond.shape)} mismatch")
        cond = cond.to(dev, torch.float32).contiguous()
        out, mo = torch.ops.cmtts_b200.denoiser_forward(self.handle, x, cond, cond_proj, steps[0], steps[1], float(c_in),
                                                        float(c_out), float(c_skip), bool(want_model_out))
        out = out.view(shape)
        return (out, mo.view(shape)) if want_model_out else out

    def project_cond(self, cond: torch.Tensor) -> torch.Tensor:
        """Conditioner projections of all residual layers (blocks.py:675) for a conditioner (B, L, hidden): one GEMM per
        batch.  The conditioner is the same for every solver step, so the sampler makes this once
        per call and hands it to `denoise_step`; nothing is cached behind the caller's back (a pointer-keyed cache goes
        stale when a buffer is refilled through raw pointers)."""
        return torch.ops.cmtts_b200.denoiser_cond(self.handle, cond.to(self.device, torch.float32).contiguous())

    def get_segmentation_model(self):
        """tts_net.py:66-73 -> (dpen callable, denoise_fun(mel[B,1,L,80]... ) in the reference's
        argument order: denoise_fun(mel, diffusion_step, conditioner[B,L,256], speaker_emb)."""

        def denoise_fun(mel, diffusion_step, conditioner, speaker_emb, mask=None):
            steps = self.prepare_steps(diffusion_step, speaker_emb)
            return self.denoise_step(mel, conditioner, steps)

        return self.duration_pitch_energy_net, denoise_fun

    def forward(self, x, timesteps, speakers=None, texts=None, src_lens=None, spker_embeds=None,
                p_control=1.0, e_control=1.0, d_control=1.0, pitch=None, **kwargs):
        """tts_net.py:75-183: conditioner from (texts, src_lens, spker_embeds), padded to x's length,
        then the denoiser on x (already scaled by c_in by the caller).  -> (B,1,L,80)."""
        if pitch is not None:
            raise NotImplementedError("training targets are out of scope")
        out = self.dpen(texts, src_lens, self.speaker_input(speakers, spker_embeds), int(x.shape[2]), p_control, e_control, d_control)
        steps = self.prepare_steps(timesteps, out["speaker_emb"])
        return self.denoise_step(x, out["cond"], steps)

    __call__ = forward

    def get_tts_loss(self):
        return None


class CMDenoiserTTS:
    """tts_net.py:13-37 — denoiser-only wrapper (unused by the reference's scripts, kept for API
    completeness): forward(x[B,1,L,80], timesteps, conditioner[B,L,256], speaker_emb)."""

    def __init__(self, total: CMTotalTTS):
        self.total = total

    def forward(self, x, timesteps, conditioner=None, speaker_emb=None, mask=None):
        steps = self.total.prepare_steps(timesteps, speaker_emb)
        return self.total.denoise_step(x, conditioner, steps)

    __call__ = forward


class KarrasDenoiser:
    """model/cm_tool/karras_diffusion.py:KarrasDenoiser — the inference members only."""

    def __init__(self, sigma_data: float = 0.5, sigma_max=80.0, sigma_min=0.002, rho=7.0,
                 weight_schedule="karras", distillation=False, loss_norm="mel_loss"):
        self.sigma_data, self.sigma_max, self.sigma_min, self.rho = sigma_data, sigma_max, sigma_min, rho
        self.weight_schedule, self.distillation, self.loss_norm = weight_schedule, distillation, loss_norm
        self.num_timesteps = 40

    def get_scalings(self, sigma):
        c_skip = self.sigma_data ** 2 / (sigma ** 2 + self.sigma_data ** 2)
        c_out = sigma * self.sigma_data / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        c_in = 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        return c_skip, c_out, c_in

    def get_scalings_for_boundary_condition(self, sigma):
        """karras_diffusion.py:87-102."""
        c_skip = self.sigma_data ** 2 / ((sigma - self.sigma_min) ** 2 + self.sigma_data ** 2)
        c_out = (sigma - self.sigma_min) * self.sigma_data / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        c_in = 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        return c_skip, c_out, c_in

    def scalar_plan(self, sigma_value: float):
        """The fp32 scalars the reference's tensor ops produce for a uniform sigma
        (`t * s_in` is an fp32 tensor; all following ops are fp32): (c_skip, c_out, c_in, 250 ln sigma)."""
        sig = torch.tensor([sigma_value], dtype=torch.float64).to(torch.float32)
        fn = self.get_scalings_for_boundary_condition if self.distillation else self.get_scalings
        c_skip, c_out, c_in = fn(sig)
        rescaled_t = 1000 * 0.25 * torch.log(sig + 1e-44)
        return float(c_skip), float(c_out), float(c_in), float(rescaled_t)

    def denoise(self, model, x_t, sigmas, **model_kwargs):
        """karras_diffusion.py:392-407 (generic form: any model callable, per-sample sigmas)."""
        fn = self.get_scalings_for_boundary_condition if self.distillation else self.get_scalings
        c_skip, c_out, c_in = [v[(...,) + (None,) * (x_t.ndim - v.ndim)] for v in fn(sigmas)]
        rescaled_t = 1000 * 0.25 * torch.log(sigmas + 1e-44)
        model_output = model(c_in * x_t, rescaled_t, **model_kwargs)
        denoised = c_out * model_output + c_skip * x_t
        return model_output, denoised


def create_model_and_diffusion_tts(use_fp16, weight_schedule, tts_model_config, sigma_min=0.002,
                                   sigma_max=80.0, distillation=False, loss_norm="mel_loss", **kwargs):
    """model/cm_tool/script_util.py:56-75."""
    model = CMTotalTTS(use_fp16=use_fp16, **tts_model_config)
    diffusion = KarrasDenoiser(sigma_data=0.5, sigma_max=sigma_max, sigma_min=sigma_min,
                               distillation=distillation, weight_schedule=weight_schedule, loss_norm=loss_norm)
    return model, diffusion
