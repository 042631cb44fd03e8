#!/usr/bin/env python
"""bench.py — headline benchmark of the CM-TTS inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--T 4] [--batch 32]

One "step" = one pass of the hot path over one synthetic LJSpeech-shape batch
(BASELINE.json configs[1]: B=32 utterances, 80..115 phonemes -> L ~ 800 mel frames, T=4 solver
steps, HiFi-GAN V1): encoder + variance adaptor -> T consistency evaluations -> vocoder -> int16.
Metric: valid mel-frames per second (sum of mel_lens / time), whole job over all N GPUs.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--T", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU")
    ap.add_argument("--dataset", default="LJSpeech")
    ap.add_argument("--src-lo", type=int, default=80)
    ap.add_argument("--src-hi", type=int, default=115)
    ap.add_argument("--cpu-sample", type=int, default=6, help="utterances in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ffma-frontend", action="store_true",
                    help="run the encoder / variance-adaptor GEMMs on the fp32 FFMA kernels instead of the hi/lo "
                         "tensor-core kernel")
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"],
                    help="tc: tcgen05 tensor cores (fp16 operands, fp32 accumulate; hi/lo pairs in the denoiser); fp32: FFMA yardstick")
    return ap.parse_args()


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
# live single-kernel probes for the roofline block: the dominant conv shapes at bench size, launched
# through the C ABI (cmtts_umma_conv1d) and timed with CUDA events on the launching stream
# ------------------------------------------------------------------------------------------------
def _probe(lib, desc, ptrs, reps=10):
    import ctypes as C
    from cmtts_b200 import _lib
    p = _lib.ptr
    args = [p(t) for t in ptrs]

    def launch():
        _lib.check(lib.cmtts_umma_conv1d(C.byref(desc), *args, _lib.stream_ptr()), "umma_conv1d")
    for _ in range(3):
        launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3          # seconds per launch


def kernel_probes(lib, dev, B, L):
    """(a) HiFi-GAN level-1 ResBlock conv, C=128 k=11 dilation 5 with residual (largest FLOP consumer of the vocoder):
    plain fp16 operands, algorithmic FLOPs == executed FLOPs.  (b) the denoiser's k=3 gate conv (K=768, N=512) on
    fp16 hi/lo pairs: 3 MMAs per algorithmic MAC.  Inputs are far larger than L2 (0.4 GB / 26 MB x 2 operands re-read
    from L2 by design), launches back to back."""
    from cmtts_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(3)
    out = {}
    # (a)
    Cc, k, dil, rows = 128, 11, 5, L * 64
    a = torch.randn(B, rows, Cc, generator=g).half().to(dev)
    w = (torch.randn(k * Cc, Cc, generator=g) / (Cc * k) ** 0.5).half().to(dev)
    bias = torch.randn(Cc, generator=g).to(dev)
    res = a.clone()
    o = torch.empty_like(a)
    d = _lib.UmmaDesc(B=B, M=rows, Lin=rows, N=Cc, Cin=Cc, taps=k, split=0, epi=0, a_ld=Cc, res_ld=Cc, out_ld=Cc, x_ld=0,
                      a_bstride=rows * Cc, res_bstride=rows * Cc, out_bstride=rows * Cc, x_bstride=0, addvec_bstride=0,
                      alpha=1.0, res_inv_slope=10.0, out_slope=0.1, out_scale=1.0, skip_accumulate=0)
    for i in range(k):
        d.shift[i] = (i - (k - 1) // 2) * dil
    t = _probe(lib, d, [a, None, w, None, bias, res, None, o, None, None, None, None])
    out["vocoder_c128_k11"] = {"seconds": t, "flops": 2.0 * B * rows * Cc * Cc * k,
                               "algorithmic_bytes": float(B * rows * Cc * 2 * 3)}
    del a, res, o
    # (b)
    Cd, R = 256, B * (L + 1)
    y = torch.randn(1, R, Cd, generator=g)
    yh = y.half(); yl = (y - yh.float()).half()
    wk = torch.randn(3 * 2 * Cd, Cd, generator=g) / (3 * Cd) ** 0.5 * 1024.0
    wh = wk.half(); wl = (wk - wh.float()).half()
    bias2 = (torch.randn(2 * Cd, generator=g) * 0.1).to(dev)
    gh = torch.empty(1, R, Cd, dtype=torch.float16, device=dev); gl = torch.empty_like(gh)
    d2 = _lib.UmmaDesc(B=1, M=R, Lin=R, N=2 * Cd, Cin=Cd, taps=3, split=1, epi=2, a_ld=Cd, res_ld=Cd, out_ld=Cd, x_ld=0,
                       a_bstride=R * Cd, res_bstride=0, out_bstride=R * Cd, x_bstride=0, addvec_bstride=0,
                       alpha=1.0 / 1024.0, res_inv_slope=1.0, out_slope=1.0, out_scale=1.0, skip_accumulate=0)
    d2.shift[0], d2.shift[1], d2.shift[2] = -1, 0, 1
    t2 = _probe(lib, d2, [yh.to(dev), yl.to(dev), wh.to(dev), wl.to(dev), bias2, None, None, gh, gl, None, None, None])
    out["denoiser_gate_k3"] = {"seconds": t2, "flops": 2.0 * R * 3 * Cd * 2 * Cd, "mma_flops": 3 * 2.0 * R * 3 * Cd * 2 * Cd}
    return out


def profile_traffic():
    """DRAM bytes per launch of the probed kernels from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic_r1.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f)
    return {}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference's CPU path, literal schedule
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(args, spec, sd, hifigan_sd, n_utt: int, steps: int, warmup: int):
    """Times the reference's own schedule on the host cores: (T+1) encoder/variance-adaptor
    passes with the Python-loop length regulator, T denoiser passes, HiFi-GAN, int16 (p_rtf_cm.py:
    174-226), through oracle/cmtts_oracle.py (`kind: port`; the Python reference cannot travel to
    the GPU box).  Bounded sample: the first `n_utt` utterances of the bench batch."""
    from cmtts_b200 import synthetic
    from oracle import cmtts_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synthetic.make_batch(spec, args.batch, args.src_lo, args.src_hi, seed=1234)
    sub = {"speakers": batch["speakers"][:n_utt], "texts": batch["texts"][:n_utt].contiguous(),
           "src_lens": batch["src_lens"][:n_utt],
           "spker_embeds": None if batch["spker_embeds"] is None else batch["spker_embeds"][:n_utt]}
    W = O.Weights(sd)
    Wf = O.Weights(synthetic.fold_weight_norm(hifigan_sd))
    g = torch.Generator().manual_seed(1)
    times, frames = [], 0
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            mel, wav, i16, pre = O.synthesize(W, Wf, spec, sub, args.T, lambda s: torch.randn(*s, generator=g), literal=True)
            dt = time.perf_counter() - t0
            frames = int(pre["mel_lens"].sum())
            if i >= warmup:
                times.append(dt)
    t = sum(times) / len(times)
    return {"value": frames / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n_utt} of {args.batch} utterances ({frames} valid frames), T={args.T}, reference schedule "
                      f"((T+1) encoder passes, Python-loop length regulator), {len(times)} timed run(s), {t:.2f} s each",
            "seconds_per_run": t}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from cmtts_b200 import synthetic
    from cmtts_b200.config import ModelSpec

    spec = ModelSpec.preset(args.dataset)
    sd = synthetic.make_acoustic_state_dict(spec, seed=0)
    hifigan_sd = synthetic.make_hifigan_checkpoint(spec.hifigan, seed=7)["generator"]
    workload = (f"{args.dataset} B={args.batch}/GPU T={args.T} phonemes {args.src_lo}..{args.src_hi} (L~800), "
                f"80 mels, HiFi-GAN V1; synthetic weights in the reference checkpoint layout")

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        cb = cpu_reference_run(args, spec, sd, hifigan_sd, args.cpu_sample, steps, min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": cb["seconds_per_run"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload + f" [CPU sample: {cb['sample']}]"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
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
    from cmtts_b200.dist import ShardedSynthesizer
    from cmtts_b200.synthesize import Pipeline

    lib = _lib.load()
    pipe = Pipeline(spec, sd, hifigan_sd, dev, precision=args.precision, tc_frontend=not args.ffma_frontend)
    synth = ShardedSynthesizer(pipe, dist if world > 1 else None)
    # per-rank shard of the global synthetic batch (weak scaling: args.batch utterances per GPU)
    gb = synthetic.make_batch(spec, args.batch * world, args.src_lo, args.src_hi, seed=1234)
    sl = slice(rank * args.batch, (rank + 1) * args.batch)
    h_texts = gb["texts"][sl].contiguous().pin_memory()
    h_lens = gb["src_lens"][sl].contiguous().pin_memory()
    h_spk = None if gb["spker_embeds"] is None else gb["spker_embeds"][sl].contiguous().pin_memory()
    d_texts, d_lens = h_texts.to(dev), h_lens.to(dev)
    d_spk = None if h_spk is None else h_spk.to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        return synth.run(d_texts, d_lens, d_spk, args.T, gather=(world > 1))

    pinned = {}

    def to_pinned(name, t):
        """device -> pinned host staging buffer (grow-only), as cmtts_b200.output does for the WAV writer"""
        buf = pinned.get(name)
        if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
            buf = pinned[name] = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
        view = buf[: t.numel()].view(t.shape)
        view.copy_(t, non_blocking=True)
        return view

    def step_e2e():
        t = h_texts.to(dev, non_blocking=True)
        l = h_lens.to(dev, non_blocking=True)
        s = None if h_spk is None else h_spk.to(dev, non_blocking=True)
        out = synth.run(t, l, s, args.T, gather=(world > 1))
        w = to_pinned("wav", out["wav_i16"])
        ml = to_pinned("mel_lens", out["mel_lens"])
        torch.cuda.current_stream(dev).synchronize()          # the step's result is on the host
        return out, w, ml

    # ---- warm-up (the clock sampler is already running: nvidia-smi needs a moment before its first report) ----
    clocks = ClockSampler(local_rank)
    t_load0 = time.time()
    for _ in range(max(args.warmup, 3)):
        out = step_resident()
    barrier()
    mel_lens = out["mel_lens"].cpu()
    B, L = out["mel"].shape[0], out["mel"].shape[1]
    valid_local = int(mel_lens.sum())
    Tsrc = h_texts.shape[1]

    # ---- timed: device-resident inputs ----
    launches0 = lib.cmtts_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        out = step_resident()
    ev1.record()
    barrier()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = lib.cmtts_launch_count() - launches0
    clk = clocks.stop(wall0, wall1, t_load0)

    # ---- timed: end to end with host buffers ----
    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        _, w_host, ml_host = step_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    # ---- instrumented pass: per-stage CUDA events (same workload, same stream) ----
    stage_ms = synth.stage_times(d_texts, d_lens, d_spk, args.T, reps=max(2, min(args.steps, 5)))

    # ---- RTF as p_rtf_cm.py defines it (rank 0 only, informational) ----
    from cmtts_b200.synthesize import rtf_like_reference
    rtf_ref, rtf_total, rtf_elapsed = rtf_like_reference(pipe, d_texts, d_lens, d_spk, args.T)

    # max over ranks / totals
    stats = torch.tensor([ms, ms_e2e, float(valid_local), float(B * L)], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, ms_e2e = float(mx[0]), float(mx[1])
        valid_total, padded_total = float(sm[2]), float(sm[3])
    else:
        valid_total, padded_total = float(valid_local), float(B * L)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    sec = ms / 1e3 / args.steps
    sec_e2e = ms_e2e / 1e3 / args.steps
    value = valid_total / sec
    hf = hifigan_flops_per_frame(spec.hifigan)
    af = acoustic_flops(spec, B, Tsrc, L, args.T)
    voc_flops = hf * B * L
    peaks = measured_peaks()
    voc_s = stage_ms["vocoder"] / 1e3
    achieved = voc_flops / voc_s / 1e12
    total_flops = (voc_flops + af["encoder"] + af["variance"] + af["denoiser"])
    h2d = h_texts.numel() * 8 + h_lens.numel() * 8 + (0 if h_spk is None else h_spk.numel() * 4)
    d2h = w_host.numel() * 2 + ml_host.numel() * 8

    # ---- roofline: the dominant conv shape timed live, kernel by kernel ----
    if args.precision == "tc":
        pr = kernel_probes(lib, dev, B, L)
        tr = profile_traffic()
        a = pr["vocoder_c128_k11"]; dn = pr["denoiser_gate_k3"]
        ach = a["flops"] / a["seconds"] / 1e12
        roofline = {"kernel": "umma_halo_kernel<128,64,2,11> (HiFi-GAN level-1 ResBlock conv, C=128 k=11 d=5 + residual; "
                              "6 launches per step, the largest FLOP consumer)",
                    "bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"],
                    "traffic": tr.get("umma_halo_kernel<128,64,2,11>", {}).get("dram_bytes_per_launch"),
                    "peak_source": peaks["source"], "algorithmic_flops_per_launch": a["flops"],
                    "algorithmic_bytes_per_launch": a["algorithmic_bytes"], "us_per_launch": a["seconds"] * 1e6,
                    "note": "timed live with CUDA events (10 back-to-back launches through the C ABI at bench size); "
                            "traffic = dram__bytes_read+write of the committed ncu --set full capture (profiles/)"}
        roofline_other = [
            {"kernel": "umma_gate_kernel<3,4> (denoiser k=3 gate conv, K=768 N=512, fp16 hi/lo: 3 MMAs per MAC; "
                       "80 launches per step)", "bound": "tensor", "unit": "TFLOP/s", "peak": peaks["tflops"],
             "achieved": dn["flops"] / dn["seconds"] / 1e12, "frac": dn["flops"] / dn["seconds"] / 1e12 / peaks["tflops"],
             "achieved_mma": dn["mma_flops"] / dn["seconds"] / 1e12, "frac_mma": dn["mma_flops"] / dn["seconds"] / 1e12 / peaks["tflops"],
             "us_per_launch": dn["seconds"] * 1e6,
             "traffic": tr.get("umma_gate_kernel<3,4>", {}).get("dram_bytes_per_launch")},
            {"kernel": "HiFi-GAN stage (all vocoder launches of one step)", "bound": "tensor", "unit": "TFLOP/s",
             "peak": peaks["tflops"], "achieved": achieved, "frac": achieved / peaks["tflops"], "stage_ms": stage_ms["vocoder"],
             "algorithmic_flops": voc_flops,
             "note": "614.1 MFLOP/mel-frame x padded frames / CUDA-event time of the vocoder stage (86% of step FLOPs)"}]
    else:
        roofline = {"kernel": synth.dominant_kernel(), "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"],
                    "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": None, "peak_source": peaks["source"],
                    "note": "fp32 FFMA yardstick path: vocoder stage aggregate"}
        roofline_other = []

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": synth.dtype_label(), "data": "synthetic",
        "config": {"workload": workload, "global_batch": B * world, "padded_frames_per_gpu": B * L, "L_max": L,
                   "Tsrc_max": Tsrc, "valid_frames_total": valid_total, "parallelism": f"utterance-sharded x{world}",
                   "l2_policy": "working set (>3 GB of activations per step) is far larger than the 126 MB L2; no flush needed",
                   "padding_mode": "global L_max (all-reduce MAX)" if world > 1 else "single batch"},
        "clocks": clk,
        "e2e": {"value": valid_total / sec_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": sec_e2e * 1e3},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_other": roofline_other,
        "stages_ms": stage_ms,
        "algorithmic_tflop_per_step": total_flops / 1e12,
        "whole_step_tflops": total_flops / sec / 1e12,
        "padded_frames_per_sec": padded_total / sec,
        "rtf": {"rtf_ref_p_rtf_cm": rtf_ref, "rtf_total": rtf_total, "elapsed_s": rtf_elapsed,
                "definition": "p_rtf_cm.py:190-230 (timer after the pre-pass; / duration of utterance 0)"},
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_run(args, spec, sd, hifigan_sd, args.cpu_sample, 1, 0)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
