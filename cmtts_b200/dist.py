"""Utterance sharding across the GPUs of one box.

The path shards naturally by utterance: every row of the batch is independent once the two
padded lengths are fixed.  Results do depend on those paddings (unmasked energy/cwt predictors,
unmasked denoiser and vocoder, inverse-CWT statistics over padded frames), which gives two modes:

  * padding="global" — every rank pads to the GLOBAL maxima, so the G-GPU output is bit-identical
    to the single-GPU batched run: Tsrc_max is known on the host when the batch is split, L_max
    needs one 8-byte MAX all-reduce after the duration predictor (`GlobalMax`, reduced on the
    device: one host sync per step, the same one the single-GPU path has);
  * padding="local" — every rank pads to its own maxima (what the reference computes when handed
    that rank's rows as a batch; its own scripts run DataLoader batches of 8, synthesize.py:366-371).
    No collective inside the model.  Combined with `balanced_partition` (length-sorted shards of
    equal padded work) the padded/valid ratio FALLS as GPUs are added instead of growing.

The only other collective is the final collation of int16 wavs + mel_lens: a gather to ONE rank
(`dst`), issued asynchronously so that it overlaps the next batch's compute; nobody needs the
wavs on all ranks.  Ranks may hold different numbers of rows (B % world != 0, balanced shards,
even zero rows): buffers are padded to the largest shard and trimmed with the row counts, which
every rank derives from the same host-side partition (no extra exchange).
One process per GPU; `dist_mod` is torch.distributed (NCCL on GPUs, gloo in the CPU tests of the
host logic) or None for a single process.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch


def shard_rows(n_rows: int, world: int, rank: int) -> slice:
    """Contiguous, balanced split of batch rows: rank r gets rows [lo, hi)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def shard_counts(n_rows: int, world: int) -> List[int]:
    return [len(range(n_rows)[shard_rows(n_rows, world, r)]) for r in range(world)]


def split_batch(batch: Dict[str, Optional[torch.Tensor]], world: int, rank: int, rows=None
                ) -> Dict[str, Optional[torch.Tensor]]:
    """Rank-local slice of a host batch, keeping the GLOBAL token padding (texts keep their width).
    `rows`: explicit row indices (e.g. from balanced_partition) instead of the contiguous split."""
    sl = shard_rows(batch["texts"].shape[0], world, rank) if rows is None else torch.as_tensor(rows, dtype=torch.int64)
    out = {}
    for k, v in batch.items():
        out[k] = None if v is None else v[sl].contiguous()
    return out


def balanced_partition(src_lens: Sequence[int], world: int) -> List[List[int]]:
    """Length-bucketed shards: rows sorted by phoneme count (longest first) and cut into `world` contiguous groups
    minimising the largest padded shard, max_r n_r * max_len_r (frames are ~proportional to phonemes, and a shard
    costs its PADDED size because the denoiser and vocoder run unmasked over padding).  Exact dynamic programme,
    O(world * n^2); deterministic, so every rank computes the same partition from the host batch.  Returns the
    original row indices per rank (possibly empty when n < world)."""
    lens = [int(x) for x in src_lens]
    order = sorted(range(len(lens)), key=lambda i: (-lens[i], i))
    n = len(order)
    if n == 0:
        return [[] for _ in range(world)]
    s = [lens[i] for i in order]
    INF = float("inf")
    # best[g][j]: minimal max-cost of cutting the first j rows into g groups; group (i, j] costs (j - i) * s[i]
    best = [[INF] * (n + 1) for _ in range(world + 1)]
    cut = [[0] * (n + 1) for _ in range(world + 1)]
    best[0][0] = 0.0
    for g in range(1, world + 1):
        for j in range(0, n + 1):
            b, c = (best[g - 1][j], j)            # an empty group is allowed (n < world)
            for i in range(j):
                if best[g - 1][i] == INF:
                    continue
                v = max(best[g - 1][i], (j - i) * s[i])
                if v < b:
                    b, c = v, i
            best[g][j], cut[g][j] = b, c
    parts: List[List[int]] = []
    j = n
    for g in range(world, 0, -1):
        i = cut[g][j]
        parts.append(sorted(order[i:j]))
        j = i
    parts.reverse()
    return parts


class GlobalMax:
    """`l_max_hook` for CMTotalTTS.dpen: MAX all-reduce of the local padded length.  Called with the 0-d device
    tensor `mel_lens.max()` it reduces on the device and returns a tensor (the caller's single `.item()` is then the
    only host sync of the step); called with an int (host-side callers, the gloo tests) it returns an int."""

    def __init__(self, dist_mod, device):
        self.dist, self.device = dist_mod, device
        self.calls = 0

    def __call__(self, local_max):
        self.calls += 1
        if self.dist is None or self.dist.get_world_size() == 1:
            return local_max
        if isinstance(local_max, torch.Tensor):
            t = local_max.reshape(1).to(torch.int64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return t[0]
        t = torch.tensor([int(local_max)], dtype=torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return int(t.item())


class LocalMaxWithWire(GlobalMax):
    """`l_max_hook` of padding="local": the shard keeps its OWN padded length, but the global maximum rides along in
    the same device tensor (-> the same single `.item()` of dpen) because it sizes the wire buffers of the final
    collation.  Returns the tensor [local_max, global_max]."""

    def __call__(self, local_max):
        self.calls += 1
        if not isinstance(local_max, torch.Tensor):
            local_max = torch.tensor(int(local_max), dtype=torch.int64, device=self.device)
        loc = local_max.reshape(1).to(torch.int64)
        if self.dist is None or self.dist.get_world_size() == 1:
            return torch.cat([loc, loc])
        glob = loc.clone()
        self.dist.all_reduce(glob, op=self.dist.ReduceOp.MAX)
        return torch.cat([loc, glob])


def _wire(t: torch.Tensor) -> torch.Tensor:
    # NCCL has no 16-bit integer type: int16 wavs travel as raw bytes
    return t.view(torch.uint8) if t.dtype == torch.int16 else t


def _pad_rows(t: torch.Tensor, rows: int) -> torch.Tensor:
    if t.shape[0] == rows:
        return t
    out = t.new_zeros((rows,) + tuple(t.shape[1:]))
    out[: t.shape[0]] = t
    return out


def gather_rows(dist_mod, t: torch.Tensor, counts: Optional[Sequence[int]] = None, dst: Optional[int] = None,
                async_op: bool = False):
    """Concatenate every rank's rows along dim 0.  `counts[r]` = rows held by rank r (None: equal shards); shards are
    padded to max(counts) on the wire and trimmed afterwards.  dst=None: all_gather, every rank gets the result;
    dst=r: gather, only rank r does (others get None).  With async_op=True returns a `PendingGather` whose
    `.result()` waits for the collective."""
    if dist_mod is None or dist_mod.get_world_size() == 1:
        return PendingGather(None, [t], [t.shape[0]], t.dtype, True) if async_op else t
    world, rank = dist_mod.get_world_size(), dist_mod.get_rank()
    counts = [t.shape[0]] * world if counts is None else [int(c) for c in counts]
    if counts[rank] != t.shape[0]:
        raise ValueError(f"gather_rows: rank {rank} holds {t.shape[0]} rows, counts say {counts[rank]}")
    width = max(counts)
    wire = _wire(_pad_rows(t.contiguous(), width))
    mine = dst is None or rank == dst
    parts = [torch.empty_like(wire) for _ in range(world)] if mine else None
    if dst is None:
        work = dist_mod.all_gather(parts, wire, async_op=async_op)
    else:
        work = dist_mod.gather(wire, parts, dst=dst, async_op=async_op)
    pend = PendingGather(work if async_op else None, parts, counts, t.dtype, mine, keep=wire)
    return pend if async_op else pend.result()


class PendingGather:
    def __init__(self, work, parts, counts, dtype, mine, keep=None):
        self.work, self.parts, self.counts, self.dtype, self.mine, self.keep = work, parts, counts, dtype, mine, keep

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None

    def result(self) -> Optional[torch.Tensor]:
        self.wait()
        if not self.mine:
            return None
        rows = [p[:c] for p, c in zip(self.parts, self.counts)]
        out = rows[0] if len(rows) == 1 else torch.cat(rows, dim=0)
        return out.view(torch.int16) if self.dtype == torch.int16 and out.dtype == torch.uint8 else out


class ShardedSynthesizer:
    """Per-rank driver of the hot path (used by bench.py and the multi-GPU tests).

    padding: "global" (bit-identical to the single-GPU batched run; one MAX all-reduce) or "local" (per-shard
    padding, no collective in the model).  `counts`: rows per rank when shards are uneven.  `dst`: rank that
    collates the wavs (None = every rank, the round-1 behaviour)."""

    def __init__(self, pipe, dist_mod=None, padding: str = "global", counts: Optional[Sequence[int]] = None,
                 dst: Optional[int] = 0):
        if padding not in ("global", "local"):
            raise ValueError("padding must be 'global' or 'local'")
        self.pipe = pipe
        self.dist = dist_mod
        self.padding = padding
        self.counts = None if counts is None else [int(c) for c in counts]
        self.dst = dst
        self.hook = GlobalMax(dist_mod, pipe.device) if padding == "global" else None
        self.wire_hook = LocalMaxWithWire(dist_mod, pipe.device) if padding == "local" else None
        self._pending: List[PendingGather] = []

    def run(self, texts, src_lens, spker_embeds, T: int, generator=None, gather: bool = False):
        multi = self.dist is not None and self.dist.get_world_size() > 1
        hook = self.hook if self.padding == "global" else (self.wire_hook if (gather and multi) else None)
        out = self.pipe(texts, src_lens, spker_embeds, T=T, generator=generator, l_max_hook=hook)
        if gather and multi:
            self.flush()                                   # at most one collation in flight
            wav = out["wav_i16"]
            if self.padding == "local":
                # shards differ in L: pad the sample axis to the widest shard, whose length came back with dpen's one
                # host round trip (LocalMaxWithWire) — no extra sync, the width only sizes the wire buffer
                w = int(out["dpen"]["l_max_extra"][0]) * self.pipe.spec.hop_length
                if w != wav.shape[1]:
                    wide = wav.new_zeros((wav.shape[0], w))
                    wide[:, : wav.shape[1]] = wav
                    wav = wide
            self._pending = [gather_rows(self.dist, wav, self.counts, self.dst, async_op=True),
                             gather_rows(self.dist, out["mel_lens"], self.counts, self.dst, async_op=True)]
            out["collation"] = self._pending
        return out

    def flush(self):
        """Wait for the collation of the previous step (call before reading `collated()` or stopping a timer)."""
        for p in self._pending:
            p.wait()

    @staticmethod
    def collated(out):
        """(wav_i16_all, mel_lens_all) on the collating rank, (None, None) elsewhere."""
        if "collation" not in out:
            return out["wav_i16"], out["mel_lens"]
        w, l = out["collation"]
        return w.result(), l.result()

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
