#!/usr/bin/env python
"""bench.py — headline benchmark of the CM-TTS inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config C1|C2|C3|C4|C5] [--T t] [--batch b] [--scaling weak|strong] [--shard balanced|contiguous]

One "step" = one pass of the hot path over one synthetic batch: encoder + variance adaptor -> T consistency
evaluations -> HiFi-GAN -> int16.  Workloads are BASELINE.json's configs:
  C1  LJSpeech, single text (1 utterance of 20..40 phonemes), T=1          — latency of single_synthesize_lj.sh
  C2  LJSpeech, 32 utterances per GPU of 80..115 phonemes (L ~ 800), T=4   — THE DEFAULT, the line the driver records
  C3  VCTK multi-speaker, global batch 64 (8 per GPU on 8 GPUs) of 20..60 phonemes, T=1
  C4  LibriTTS zero-shot shapes, global batch 128 (16 per GPU on 8 GPUs) of 60..150 phonemes, T=4
  C5  HiFi-GAN generator only: (batch, 80, 1024) mels -> int16, --batch 1..256
--scaling weak keeps the per-GPU batch fixed as N grows, strong keeps the global batch fixed.
Metric: valid mel-frames per second (sum of mel_lens / time), whole job over all N GPUs.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "mel_frames_per_sec"
UNIT = "mel-frames/s"

# per_gpu: utterances per GPU under weak scaling; global_batch: the fixed batch under strong scaling
CONFIGS = {
    "C1": dict(dataset="LJSpeech", per_gpu=1, global_batch=1, T=1, lo=20, hi=40,
               what="LJSpeech single text (single_synthesize_lj.sh:2-7)"),
    "C2": dict(dataset="LJSpeech", per_gpu=32, global_batch=32, T=4, lo=80, hi=115,
               what="LJSpeech batch=32, ~800 frames (BASELINE.json configs[1])"),
    "C3": dict(dataset="VCTK", per_gpu=8, global_batch=64, T=1, lo=20, hi=60,
               what="VCTK multi-speaker batch=64 over 8 GPUs (configs[2])"),
    "C4": dict(dataset="LibriTTS", per_gpu=16, global_batch=128, T=4, lo=60, hi=150,
               what="LibriTTS zero-shot batch=128 over 8 GPUs (configs[3])"),
    "C5": dict(dataset="LJSpeech", per_gpu=32, global_batch=32, T=0, lo=0, hi=0,
               what="HiFi-GAN generator only, 80 x 1024 mels (configs[4])"),
}
C5_FRAMES = 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--T", type=int, default=None, help="solver steps (default: the config's)")
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU (weak) / global batch (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--shard", default="balanced", choices=["balanced", "contiguous"],
                    help="N>1: balanced = length-bucketed shards, per-shard padding (no collective in the model); "
                         "contiguous = equal row slices padded to the global L_max (bit-identical to the single-GPU batch)")
    ap.add_argument("--dataset", default=None)
    ap.add_argument("--src-lo", type=int, default=None)
    ap.add_argument("--src-hi", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=None, help="utterances in the CPU sample (default: 4 beside the GPU "
                    "arm, 8 = the reference's own DataLoader batch size in the reference arm)")
    ap.add_argument("--zero-shot", default="auto", choices=["auto", "on", "off"],
                    help="include the zero-shot path's speaker encoder in the step: the DeepSpeaker embedding of one reference "
                         "recording per batch (synthesize_zeroshot_*.py) computed on the GPU and used for every utterance "
                         "(auto: on for C4, BASELINE.json's zero-shot config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel launch profile / roofline pass")
    ap.add_argument("--graphs", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from CUDA graphs (cmtts_b200.synthesize.Pipeline(graphs=True)); auto = on for the "
                         "launch-latency-bound configs (C1, C3), off for the GPU-bound ones")
    ap.add_argument("--synthetic-vocoder", action="store_true",
                    help="synthetic HiFi-GAN weights even when the reference's generator_universal.pth.tar is staged")
    ap.add_argument("--ffma-frontend", action="store_true",
                    help="run the encoder / variance-adaptor GEMMs on the fp32 FFMA kernels instead of the hi/lo "
                         "tensor-core kernel")
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"],
                    help="tc: tcgen05 tensor cores (fp16 operands, fp32 accumulate; hi/lo pairs in the denoiser); fp32: FFMA yardstick")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    a.dataset = a.dataset or c["dataset"]
    a.T = c["T"] if a.T is None else a.T
    a.src_lo = c["lo"] if a.src_lo is None else a.src_lo
    a.src_hi = c["hi"] if a.src_hi is None else a.src_hi
    return a


def global_batch_size(args, world: int) -> int:
    c = CONFIGS[args.config]
    if args.scaling == "weak":
        return (args.batch if args.batch is not None else c["per_gpu"]) * world
    return args.batch if args.batch is not None else c["global_batch"]


# ------------------------------------------------------------------------------------------------
# algorithmic work (2 * MAC on padded shapes — padded frames are part of the result)
# ------------------------------------------------------------------------------------------------
def hifigan_flops_per_frame(hs) -> float:
    f = 2.0 * 80 * hs.upsample_initial_channel * 7
    ch, rate = hs.upsample_initial_channel, 1
    for u, k in zip(hs.upsample_rates, hs.upsample_kernel_sizes):
        cin, ch = ch, ch // 2
        rate *= u
        f += 2.0 * cin * ch * (k / u) * rate                       # ConvTranspose: k/u taps per output sample
        for rk, dils in zip(hs.resblock_kernel_sizes, hs.resblock_dilation_sizes):
            f += 2.0 * ch * ch * rk * 2 * len(dils) * rate
    f += 2.0 * ch * 7 * rate
    return f


def acoustic_flops(spec, B, Tsrc, L, T) -> dict:
    H, C, M = spec.hidden, spec.res_channels, spec.n_mels
    n_tok, n_frm = B * Tsrc, B * L
    enc = n_tok * spec.enc_layers * (2.0 * H * 3 * H + 2.0 * H * H + 4.0 * Tsrc * H
                                     + 2.0 * H * 4 * H * spec.ffn_kernel + 2.0 * 4 * H * H)
    va = n_tok * (2.0 * H * spec.filter_size * spec.dur_kernel * 2 + 2.0 * H * spec.filter_size * spec.pred_kernel * 2) \
        + n_frm * (2.0 * H * spec.cwt_hidden + 2.0 * spec.cwt_hidden * spec.filter_size * spec.pred_kernel
                   + 2.0 * spec.filter_size ** 2 * spec.pred_kernel)
    per_frame_step = 2.0 * M * C + spec.res_layers * (2.0 * H * C + 2.0 * C * 2 * C * 3 + 2.0 * C * 2 * C) + 2.0 * C * C + 2.0 * C * M
    dn = T * n_frm * per_frame_step
    return {"encoder": enc, "variance": va, "denoiser": dn, "denoiser_per_frame_step": per_frame_step}


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(gpu_index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1, t_load0=None):
        """Samples inside the timed region [t0, t1]; if it was too short for nvidia-smi to report at least two, the
        samples taken under the same load since `t_load0` (the warm-up steps) are used and the line says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        good = [(t, r) for (t, r) in self.rows if len(r) >= 9]
        rows = [r for (t, r) in good if t0 <= t <= t1 + 0.05]
        window = "timed region"
        if len(rows) < 2 and t_load0 is not None:
            rows = [r for (t, r) in good if t_load0 <= t <= t1 + 0.05]
            window = "warm-up + timed region (same load)"
        if not rows:
            rows = [r for (_, r) in good]
            window = "whole run"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "window": window, "power_w_max": max(float(r[3]) for r in rows)}


# ------------------------------------------------------------------------------------------------
# per-kernel launch profile of ONE step (cmtts_prof_begin / cmtts_prof_end of the C ABI: a CUDA event after every
# launch of the library on the launching stream) -> each kernel's share of the step, algorithmic rate, roofline
# ------------------------------------------------------------------------------------------------
def kernel_profile(lib, step_fn, dev):
    from cmtts_b200 import _lib
    torch.cuda.synchronize(dev)
    _lib.check(lib.cmtts_prof_begin(_lib.stream_ptr(dev)), "prof_begin")
    step_fn()
    cap = 1 << 18
    buf = ctypes.create_string_buffer(cap)
    n = lib.cmtts_prof_end(buf, cap)
    if n < 0:
        _lib.check(int(n), "prof_end")
    rows = []
    for line in buf.value.decode().splitlines():
        label, cnt, us, fl, by = line.split("\t")
        rows.append({"kernel": label, "launches": int(cnt), "us": float(us), "flops": float(fl), "bytes": float(by)})
    tot = sum(r["us"] for r in rows) or 1.0
    for r in rows:
        r["share"] = r["us"] / tot
    rows.sort(key=lambda r: -r["us"])
    return rows, tot


def traffic_table():
    """DRAM bytes per launch from the committed `ncu --set full` capture of THIS round's build (profiles/), keyed by the
    profiler label; {} if none has been committed for the current sources."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic_r2.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f)
    return {}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tflops_burst": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_sustained": 1400.0, "tflops_burst": 1650.0, "source": "fallback (B200_PROFILING.md)"}


def roofline_blocks(rows, total_us, peaks, step_ms, total_flops, stage_ms, stage_flops):
    """`roofline` = the kernel with the LARGEST share of the profiled step, rated against the SUSTAINED measured peak
    (it is timed inside a long step); achieved = its algorithmic FLOPs (or bytes) / its in-step time."""
    ridge = peaks["tflops_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)     # FLOP per byte
    tr = traffic_table()

    def block(r):
        sec = r["us"] * 1e-6
        tensor = r["flops"] > 0 and (r["bytes"] <= 0 or r["flops"] / r["bytes"] >= ridge or "hi/lo" in r["kernel"]
                                     or ",1,e" in r["kernel"])
        if tensor:
            ach = r["flops"] / sec / 1e12
            b = {"kernel": r["kernel"], "bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"],
                 "unit": "TFLOP/s", "frac": ach / peaks["tflops_sustained"]}
            if "e4m3" in r["kernel"]:
                b["mma_frac"] = 2.0 * b["frac"]
                b["note"] = ("fp16 hi/lo operand pairs with e4m3 cross terms: per algorithmic MAC one f16 MMA + two f8f6f4 MMAs at "
                             "twice the rate = 2 f16-equivalents (mma_frac = executed tensor work / f16 peak)")
            elif "hi/lo" in r["kernel"] or ",1,e" in r["kernel"]:
                b["mma_frac"] = 3.0 * b["frac"]
                b["note"] = "fp16 hi/lo operand pairs: 3 tcgen05.mma per algorithmic MAC (mma_frac = executed MMA rate / peak)"
        else:
            ach = r["bytes"] / sec / 1e9 if r["bytes"] > 0 else 0.0
            b = {"kernel": r["kernel"], "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": ach / peaks["hbm_gbs"]}
        keys = [k for k in tr if not k.startswith("_") and r["kernel"].startswith(k)]
        t = tr[max(keys, key=len)] if keys else {}
        b.update({"traffic": t.get("dram_bytes_per_launch"), "traffic_tensor_pipe_pct_ncu": t.get("tensor_pipe_pct"), "launches_per_step": r["launches"],
                  "us_per_launch": r["us"] / r["launches"], "share_of_step": r["share"],
                  "algorithmic_flops_per_launch": r["flops"] / r["launches"],
                  "algorithmic_bytes_per_launch": r["bytes"] / r["launches"]})
        return b

    top = block(rows[0])
    from cmtts_b200.build import kernel_stamp
    top["traffic_same_build"] = bool(tr.get("_kernel_stamp")) and kernel_stamp() == tr.get("_kernel_stamp")   # same kernel sources as the capture
    top["peak_source"] = peaks["source"] + ", sustained (kernel timed inside the step)"
    top["step_frac"] = total_flops / (step_ms * 1e-3) / 1e12 / peaks["tflops_sustained"]
    top["stage_fracs"] = {k: (stage_flops[k] / (v * 1e-3) / 1e12 / peaks["tflops_sustained"]) for k, v in stage_ms.items()
                          if v > 0 and k in stage_flops}
    top["how"] = ("per-kernel time = distance between CUDA events recorded after every launch of one profiled step "
                  f"(sum {total_us / 1e3:.2f} ms vs {step_ms:.2f} ms unprofiled); traffic = dram bytes of the committed ncu "
                  "--set full capture (profiles/roofline_traffic_r2.json) or null")
    others = [block(r) for r in rows[1:8]]
    return top, others


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own CPU path (oracle/ref_bench.py)
# ------------------------------------------------------------------------------------------------
def real_hifigan_path():
    for p in (os.path.join(ROOT, "oracle", "_ref", "hifigan", "generator_universal.pth.tar"),
              "/root/reference/hifigan/generator_universal.pth.tar"):
        if os.path.isfile(p):
            return p
    return None


def load_hifigan(spec, synthetic_only: bool):
    """BASELINE.json names the universal HiFi-GAN checkpoint the reference ships; it is staged (git-ignored) by
    __graft_entry__.build().  Falls back to synthetic weights in the same checkpoint layout."""
    from cmtts_b200 import synthetic
    p = None if synthetic_only else real_hifigan_path()
    if p is not None:
        return torch.load(p, map_location="cpu", weights_only=True)["generator"], "generator_universal.pth.tar (reference's shipped weights)"
    return synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)["generator"], "synthetic (reference checkpoint layout)"


def cpu_reference_run(args, spec, sd, hifigan_sd, batch, n_utt: int, steps: int, warmup: int):
    """Times the reference's own CPU path on the host cores (oracle/ref_bench.py: the unmodified reference when its
    tree is available — /root/reference or the staged oracle/_ref — else the oracle port) on a bounded sample: the
    first `n_utt` utterances of the bench batch, same T."""
    import contextlib
    import traceback
    from oracle.ref_bench import ReferenceRunner

    def timed(force_port):
        h = hifigan_sd
        if force_port and h is None:
            h = load_hifigan(spec, False)[0]
        runner = ReferenceRunner(args.dataset, spec, sd, h, force_port=force_port)
        if args.config == "C5":
            from cmtts_b200 import synthetic
            mel = synthetic.make_mels(n_utt, spec.n_mels, C5_FRAMES, seed=99)
            runs = [runner.vocoder_step(mel) for _ in range(warmup + steps)][warmup:]
            sample = f"{n_utt} of the batch's (80 x {C5_FRAMES}) mels"
        else:
            sub = {k: (None if v is None else v[:n_utt].contiguous()) for k, v in batch.items()}
            runs = [runner.step(sub, args.T) for _ in range(warmup + steps)][warmup:]
            sample = (f"first {n_utt} utterances of the batch, T={args.T}, the reference's schedule ((T+1) encoder passes, "
                      f"Python-loop length regulator)")
        return runner, runs, sample

    with contextlib.redirect_stdout(sys.stderr):          # the reference prints ("Removing weight norm..."); stdout is the JSON line's
        fallback_note = ""
        try:
            runner, runs, sample = timed(False)
        except Exception as e:                            # reference tree unusable on this host: time the oracle port, and say so
            traceback.print_exc(file=sys.stderr)
            runner, runs, sample = timed(True)
            fallback_note = f" [the reference tree failed here ({type(e).__name__}: {e}); oracle port timed instead]"
    t = sum(r["seconds"] for r in runs) / len(runs)
    frames = runs[0]["valid_frames"]
    out = {"value": frames / t, "unit": UNIT, "cores": runner.cores, "kind": runner.kind,
           "sample": f"{sample}; {frames} valid frames, {len(runs)} timed run(s), {t:.2f} s each{fallback_note}",
           "seconds_per_run": t}
    if "first_utt_seconds_audio" in runs[0]:
        t_rtf = sum(r["seconds_after_prepass"] for r in runs) / len(runs)
        out["rtf_ref_p_rtf_cm"] = t_rtf / runs[0]["first_utt_seconds_audio"]
    return out


def workload_text(args, hifigan_src, B_per_gpu_text):
    c = CONFIGS[args.config]
    if args.config == "C5":
        return f"C5: {c['what']}; batch {B_per_gpu_text}; HiFi-GAN V1 weights: {hifigan_src}"
    return (f"{args.config}: {c['what']}; {args.dataset} {B_per_gpu_text} T={args.T} phonemes {args.src_lo}..{args.src_hi}, "
            f"80 mels, HiFi-GAN V1 weights: {hifigan_src}; acoustic weights synthetic in the reference checkpoint layout")


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from cmtts_b200 import synthetic
    from cmtts_b200.config import ModelSpec

    spec = ModelSpec.preset(args.dataset)
    sd = synthetic.make_acoustic_state_dict(spec, seed=0)
    hifigan_sd, hifigan_src = load_hifigan(spec, args.synthetic_vocoder)
    GB = global_batch_size(args, world if args.impl == "ours" else max(args.gpus, 1))
    gb = None if args.config == "C5" else synthetic.make_batch(spec, GB, args.src_lo, args.src_hi, seed=1234)
    btxt = (f"{GB // max(world, 1)} utterances/GPU (weak)" if args.scaling == "weak" else f"global batch {GB} (strong)")

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
        n_utt = min(args.cpu_sample or 8, GB)
        cb = cpu_reference_run(args, spec, sd, None if real_hifigan_path() and not args.synthetic_vocoder else hifigan_sd,
                               gb, n_utt, steps, warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": cb["seconds_per_run"] * 1e3,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_text(args, hifigan_src, btxt) + f" [CPU sample: {cb['sample']}]"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        if "rtf_ref_p_rtf_cm" in cb:
            line["rtf"] = {"rtf_ref_p_rtf_cm": cb["rtf_ref_p_rtf_cm"], "definition": "p_rtf_cm.py:190-230 on the CPU sample"}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from cmtts_b200 import _lib
    from cmtts_b200.dist import ShardedSynthesizer, balanced_partition, shard_counts, shard_rows
    from cmtts_b200.synthesize import Pipeline

    lib = _lib.load()
    use_graphs = args.graphs == "on" or (args.graphs == "auto" and args.config in ("C1", "C3"))
    zero_shot = args.zero_shot == "on" or (args.zero_shot == "auto" and args.config == "C4")
    pipe = Pipeline(spec, sd, hifigan_sd, dev, precision=args.precision, tc_frontend=not args.ffma_frontend, graphs=use_graphs)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    pinned = {}

    def to_pinned(name, t):
        """device -> pinned host staging buffer (grow-only), as cmtts_b200.output does for the WAV writer"""
        buf = pinned.get(name)
        if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
            buf = pinned[name] = torch.empty(max(t.numel(), 1), dtype=t.dtype, pin_memory=True)
        view = buf[: t.numel()].view(t.shape)
        view.copy_(t, non_blocking=True)
        return view

    # ---- the per-rank shard --------------------------------------------------------------------
    if args.config == "C5":
        rows = list(range(GB))[shard_rows(GB, world, rank)]
        mels = synthetic.make_mels(GB, spec.n_mels, C5_FRAMES, seed=99)[rows].transpose(1, 2).contiguous()   # (b, L, 80)
        h_mel = mels.pin_memory()
        d_mel = h_mel.to(dev)
        synth = ShardedSynthesizer(pipe, dist if world > 1 else None)
        padding = "n/a"

        def step_resident():
            _, w16 = pipe.vocoder.run(d_mel, want_float=False, want_int16=True, max_wav_value=spec.max_wav_value)
            return {"wav_i16": w16, "mel_lens": None}

        def step_e2e():
            m = h_mel.to(dev, non_blocking=True)
            _, w16 = pipe.vocoder.run(m, want_float=False, want_int16=True, max_wav_value=spec.max_wav_value)
            w = to_pinned("wav", w16)
            torch.cuda.current_stream(dev).synchronize()
            return w, None

        h2d = h_mel.numel() * 4
    else:
        if world > 1 and args.shard == "balanced":
            parts = balanced_partition(gb["src_lens"].tolist(), world)
            rows, counts, padding = parts[rank], [len(p) for p in parts], "local"
        else:
            rows, counts, padding = list(range(GB))[shard_rows(GB, world, rank)], shard_counts(GB, world), "global"
        idx = torch.as_tensor(rows, dtype=torch.int64)
        # every shard keeps the GLOBAL token padding in "global" mode; per-shard token padding in "local" mode
        h_lens = gb["src_lens"][idx].contiguous()
        tmax = int(gb["src_lens"].max()) if padding == "global" or not len(rows) else int(h_lens.max())
        h_texts = gb["texts"][idx][:, :tmax].contiguous().pin_memory()
        h_lens = h_lens.pin_memory()
        h_spk = None if gb["spker_embeds"] is None else gb["spker_embeds"][idx].contiguous().pin_memory()
        d_texts, d_lens = h_texts.to(dev), h_lens.to(dev)
        d_spk = None if h_spk is None else h_spk.to(dev)
        synth = ShardedSynthesizer(pipe, dist if world > 1 else None, padding=padding, counts=counts, dst=0)

        # zero-shot (BASELINE.json configs[3]; synthesize_zeroshot_lj.py:93-102): every utterance of the batch is conditioned on
        # the DeepSpeaker embedding of ONE reference recording, computed inside the step (cmtts_rescnn_forward).  The host side
        # of it (WAV decode, filter-bank features: numpy, as in the reference) runs once, outside the timed region; its
        # 160 x 64 window is resident (value) or uploaded from pinned memory every step (e2e).
        zs_enc, zs_src = None, None
        if zero_shot and h_spk is not None and len(rows):
            from cmtts_b200 import speaker_encoder as SE
            zs_enc = SE.DeepSpeakerModel(str(dev))
            ck = os.path.join(ROOT, "oracle", "_ref", "deepspeaker", "pretrained_models", "ResCNN_triplet_training_checkpoint_265.h5")
            if os.path.isfile(ck):
                zs_enc.load_weights(ck)
                zs_src = "reference checkpoint ResCNN_triplet_training_checkpoint_265.h5"
            else:
                zs_enc.set_keras_weights(synthetic.make_deepspeaker_weights(seed=0))
                zs_src = "synthetic weights (checkpoint not staged)"
            mfcc = SE.read_mfcc(synthetic.make_voice_like(2.5, SE.SAMPLE_RATE, seed=3), SE.SAMPLE_RATE, SE.WIN_LENGTH)
            h_fb = torch.from_numpy(SE.sample_from_mfcc(mfcc, SE.NUM_FRAMES, offset=0)[None, ..., 0].copy()).pin_memory()
            d_fb = h_fb.to(dev)
            n_rows = int(h_spk.shape[0])

        def zs_embed(fb):
            return zs_enc.predict_tensor(fb).expand(n_rows, -1).contiguous()

        def step_resident():
            spk = zs_embed(d_fb) if zs_enc is not None else d_spk
            return synth.run(d_texts, d_lens, spk, args.T, gather=(world > 1))

        def step_e2e():
            t = h_texts.to(dev, non_blocking=True)
            l = h_lens.to(dev, non_blocking=True)
            if zs_enc is not None:
                s = zs_embed(h_fb.to(dev, non_blocking=True))
            else:
                s = None if h_spk is None else h_spk.to(dev, non_blocking=True)
            out = synth.run(t, l, s, args.T, gather=(world > 1))
            w = to_pinned("wav", out["wav_i16"])
            ml = to_pinned("mel_lens", out["mel_lens"])
            torch.cuda.current_stream(dev).synchronize()          # the step's result is on the host
            return w, ml

        h2d = h_texts.numel() * 8 + h_lens.numel() * 8 + (0 if h_spk is None else (h_fb.numel() if zs_enc is not None else h_spk.numel()) * 4)

    # ---- warm-up (the clock sampler is already running: nvidia-smi needs a moment before its first report) ----
    clocks = ClockSampler(local_rank)
    t_load0 = time.time()
    for _ in range(max(args.warmup, 3)):
        out = step_resident()
    synth.flush()
    barrier()
    if args.config == "C5":
        B, L, Tsrc = d_mel.shape[0], C5_FRAMES, 0
        valid_local = B * L
    else:
        B, L = out["mel"].shape[0], out["mel"].shape[1]
        valid_local = int(out["mel_lens"].sum().item()) if B else 0
        Tsrc = h_texts.shape[1]

    # ---- timed: device-resident inputs ----
    launches0 = lib.cmtts_launch_count() + pipe.graph_kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        out = step_resident()
    synth.flush()                                   # the last step's collation is part of the job
    ev1.record()
    barrier()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = lib.cmtts_launch_count() + pipe.graph_kernel_launches - launches0
    clk = clocks.stop(wall0, wall1, t_load0)

    # ---- timed: end to end with host buffers ----
    for _ in range(2):
        step_e2e()
    synth.flush()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        w_host, ml_host = step_e2e()
    synth.flush()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    d2h = w_host.numel() * 2 + (0 if ml_host is None else ml_host.numel() * 8)

    # ---- instrumented passes: per-stage CUDA events, per-kernel launch profile (same workload, same stream) ----
    hf = hifigan_flops_per_frame(spec.hifigan)
    if args.config == "C5":
        stage_ms = {"vocoder": ms / args.steps}
        stage_flops = {"vocoder": hf * B * L}
        total_flops = stage_flops["vocoder"]
        rtf = None
        prof_step = step_resident
    else:
        stage_ms = synth.stage_times(d_texts, d_lens, d_spk, args.T, reps=max(2, min(args.steps, 5))) if B else {}
        af = acoustic_flops(spec, B, Tsrc, L, args.T)
        stage_flops = {"dpen": af["encoder"] + af["variance"], "sampler": af["denoiser"], "vocoder": hf * B * L}
        if zs_enc is not None and B:
            z0, z1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            zs_embed(d_fb)
            z0.record()
            for _ in range(5):
                zs_embed(d_fb)
            z1.record()
            torch.cuda.synchronize(dev)
            stage_ms["speaker_encoder"] = z0.elapsed_time(z1) / 5
            # 28 convs of the ResCNN on a 160 x 64 window (deepspeaker/conv_models.py:110-133) + Dense
            zf, hh, ww, ci = 0.0, 160, 64, 1
            for co in (64, 128, 256, 512):
                hh, ww = (hh + 1) // 2, (ww + 1) // 2
                zf += 2.0 * hh * ww * co * (25 * ci + 6 * 9 * co)
                ci = co
            stage_flops["speaker_encoder"] = zf + 2.0 * 2048 * 512
        total_flops = sum(stage_flops.values())
        # RTF as p_rtf_cm.py defines it (informational; every rank computes its own, rank 0 reports)
        from cmtts_b200.synthesize import rtf_like_reference
        rtf = rtf_like_reference(pipe, d_texts, d_lens, d_spk, args.T) if B else None
        def prof_step():                 # the profiler times launches: one EAGER step (graph replays launch nothing through the library)
            was, pipe.graphs = pipe.graphs, False
            try:
                pipe(d_texts, d_lens, zs_embed(d_fb) if zs_enc is not None else d_spk, T=args.T)
            finally:
                pipe.graphs = was
    prof = None
    if not args.no_profile and B:
        prof = kernel_profile(lib, prof_step, dev)

    # max over ranks / totals
    stats = torch.tensor([ms, ms_e2e, float(valid_local), float(B * L), float(launches)], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, ms_e2e = float(mx[0]), float(mx[1])
        valid_total, padded_total, launches_total = float(sm[2]), float(sm[3]), float(sm[4])
    else:
        valid_total, padded_total, launches_total = float(valid_local), float(B * L), float(launches)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    sec = ms / 1e3 / args.steps
    sec_e2e = ms_e2e / 1e3 / args.steps
    value = valid_total / sec
    peaks = measured_peaks()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": synth.dtype_label(), "data": "synthetic",
        "config": {"workload": workload_text(args, hifigan_src, btxt), "name": args.config, "T": args.T,
                   "global_batch": GB, "rank0_utterances": B, "rank0_padded_frames": B * L, "rank0_L_max": L,
                   "rank0_Tsrc_max": Tsrc, "valid_frames_total": valid_total, "padded_frames_total": padded_total,
                   "padded_over_valid": padded_total / max(valid_total, 1.0),
                   "parallelism": f"utterance-sharded x{world}" + (f", {args.shard} shards" if world > 1 else ""),
                   "padding_mode": {"global": "global L_max (one 8-byte MAX all-reduce; bit-identical to the single-GPU batch)",
                                    "local": "per-shard L_max (length-bucketed shards; each shard = the reference run on its rows)",
                                    "n/a": "fixed-length mels"}[padding] if world > 1 else "single batch",
                   "cuda_graphs": bool(use_graphs), "graph_replays": int(pipe.graph_replays),
                   "zero_shot_speaker_encoder": (zs_src if args.config != "C5" and zs_enc is not None else None),
                   "collation": "async gather of int16 wavs + mel_lens to rank 0, inside the timed region" if world > 1 else "none",
                   "l2_policy": "working set (GBs of activations per step) is far larger than the 126 MB L2; no flush needed"
                   if B * L >= 8000 else "small batch: activations of consecutive launches stay L2-resident by design "
                                         "(one step's working set is the workload; inputs are re-uploaded / re-generated every step)"},
        "clocks": clk,
        "e2e": {"value": valid_total / sec_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": sec_e2e * 1e3},
        "gpu_launches": int(launches_total),
        "stages_ms": stage_ms,
        "algorithmic_tflop_per_step": total_flops / 1e12,
        "whole_step_tflops": total_flops * (1 if world == 1 else padded_total / max(B * L, 1)) / sec / 1e12,
        "padded_frames_per_sec": padded_total / sec,
    }
    if rtf is not None:
        line["rtf"] = {"rtf_ref_p_rtf_cm": rtf[0], "rtf_total": rtf[1], "elapsed_s": rtf[2],
                       "definition": "p_rtf_cm.py:190-230 (timer after the pre-pass; / duration of utterance 0)"}
    if prof is not None:
        rows_, tot_us = prof
        top, others = roofline_blocks(rows_, tot_us, peaks, sec * 1e3, total_flops, stage_ms, stage_flops)
        line["roofline"] = top
        line["roofline_other"] = others
        line["kernels"] = [{"kernel": r["kernel"], "launches": r["launches"], "ms": r["us"] / 1e3, "share": r["share"]}
                           for r in rows_[:40]]
    else:
        line["roofline"] = {"kernel": synth.dominant_kernel(), "bound": "tensor", "achieved": total_flops / sec / 1e12,
                            "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                            "frac": total_flops / sec / 1e12 / peaks["tflops_sustained"], "traffic": None,
                            "note": "whole-step aggregate (per-kernel profile skipped)"}
    if world == 1 and not args.no_cpu_baseline:
        n_utt = min(args.cpu_sample or 4, GB)
        try:
            cb = cpu_reference_run(args, spec, sd, None if real_hifigan_path() and not args.synthetic_vocoder else hifigan_sd,
                                   gb, n_utt, 1, 0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:          # the CPU leg is a reported baseline: its failure must not cost the measured GPU line
            import traceback
            traceback.print_exc(file=sys.stderr)
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable",
                                    "sample": f"CPU baseline failed on this host: {type(e).__name__}: {e}"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
