"""Utterance sharding across the GPUs of one box.

The path shards naturally by utterance: every row of the batch is independent once the two
padded lengths are fixed.  Results do depend on those paddings (unmasked energy/cwt predictors,
unmasked denoiser and vocoder, inverse-CWT statistics over padded frames), so to stay
bit-comparable with the single-GPU batched run every rank pads to the GLOBAL maxima:
  * Tsrc_max is known on the host when the batch is split;
  * L_max needs one 8-byte MAX all-reduce after the duration predictor (`l_max_hook`).
The only other collective is the final gather of int16 wavs + mel_lens (+ mels) to every rank.
One process per GPU; `dist_mod` is torch.distributed (NCCL on GPUs, gloo in the CPU tests of the
host logic) or None for a single process.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


def shard_rows(n_rows: int, world: int, rank: int) -> slice:
    """Contiguous, balanced split of batch rows: rank r gets rows [lo, hi)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def split_batch(batch: Dict[str, Optional[torch.Tensor]], world: int, rank: int) -> Dict[str, Optional[torch.Tensor]]:
    """Rank-local slice of a host batch, keeping the GLOBAL token padding (texts keep their width)."""
    sl = shard_rows(batch["texts"].shape[0], world, rank)
    out = {}
    for k, v in batch.items():
        out[k] = None if v is None else v[sl].contiguous()
    return out


class GlobalMax:
    """`l_max_hook` for CMTotalTTS.dpen: MAX all-reduce of the local padded length."""

    def __init__(self, dist_mod, device):
        self.dist, self.device = dist_mod, device
        self.calls = 0

    def __call__(self, local_max: int) -> int:
        self.calls += 1
        if self.dist is None or self.dist.get_world_size() == 1:
            return local_max
        t = torch.tensor([local_max], dtype=torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return int(t.item())


def gather_rows(dist_mod, t: torch.Tensor) -> torch.Tensor:
    """all_gather along dim 0 (equal shapes on every rank thanks to the global paddings)."""
    if dist_mod is None or dist_mod.get_world_size() == 1:
        return t
    t = t.contiguous()
    # NCCL has no 16-bit integer type: int16 wavs travel as raw bytes
    wire = t.view(torch.uint8) if t.dtype == torch.int16 else t
    parts = [torch.empty_like(wire) for _ in range(dist_mod.get_world_size())]
    dist_mod.all_gather(parts, wire)
    out = torch.cat(parts, dim=0)
    return out.view(torch.int16) if t.dtype == torch.int16 else out


class ShardedSynthesizer:
    """Per-rank driver of the hot path (used by bench.py and the multi-GPU tests)."""

    def __init__(self, pipe, dist_mod=None):
        self.pipe = pipe
        self.dist = dist_mod
        self.hook = GlobalMax(dist_mod, pipe.device)

    def run(self, texts, src_lens, spker_embeds, T: int, generator=None, gather: bool = False):
        out = self.pipe(texts, src_lens, spker_embeds, T=T, generator=generator, l_max_hook=self.hook)
        if gather and self.dist is not None:
            out["wav_i16_all"] = gather_rows(self.dist, out["wav_i16"])
            out["mel_lens_all"] = gather_rows(self.dist, out["mel_lens"])
        return out

    def stage_times(self, texts, src_lens, spker_embeds, T: int, reps: int = 3) -> Dict[str, float]:
        """Instrumented passes of the same step: CUDA events on the launching stream around each
        stage.  Returns the MEDIAN milliseconds per stage over `reps` passes (one pass may absorb an allocator
        or clock hiccup that a mean would smear over the stage numbers)."""
        from .sampler import karras_sample_tts, sampler_plan

        pipe, dev = self.pipe, self.pipe.device
        names = ["dpen", "sampler", "vocoder"]
        acc = {n: [] for n in names}
        for _ in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            out = pipe.model.dpen(texts, src_lens, spker_embeds, None, l_max_hook=self.hook)
            ev[1].record()
            B, L, _ = out["cond"].shape
            sampler, steps, ts = sampler_plan(T)
            kw = {"texts": texts, "src_lens": src_lens, "spker_embeds": spker_embeds}
            mel = karras_sample_tts(pipe.diffusion, pipe.model, (B, 1, L, pipe.spec.n_mels), steps=steps,
                                    model_kwargs=kw, device=dev, sigma_min=pipe.spec.sigma_min,
                                    sigma_max=pipe.spec.sigma_max, sampler=sampler, ts=ts, cond_dict=out)
            ev[2].record()
            pipe.vocoder.run(mel, want_float=False, want_int16=True, max_wav_value=pipe.spec.max_wav_value)
            ev[3].record()
            torch.cuda.synchronize(dev)
            for i, n in enumerate(names):
                acc[n].append(ev[i].elapsed_time(ev[i + 1]))
        return {n: sorted(v)[len(v) // 2] for n, v in acc.items()}

    def dtype_label(self) -> str:
        if self.pipe.precision == "tc":
            return "f16 operands / f32 accumulate (tcgen05); vocoder plain f16, denoiser + encoder + variance adaptor f16 hi+lo pairs (fp32-class products)"
        return "f32"

    def dominant_kernel(self) -> str:
        if self.pipe.precision == "tc":
            return "umma_conv_kernel (tcgen05 implicit-GEMM conv, TMA + TMEM)"
        return "conv1d_simt_kernel (fp32 FFMA implicit GEMM)"
