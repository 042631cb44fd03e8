"""BASELINE.json config 5: HiFi-GAN generator only, 80 x 1024 mel -> 22.05 kHz wav, batch sweep 1..256 on one B200.

    python tools/vocoder_sweep.py [--frames 1024] [--batches 1,2,4,...,256] [--reps 5] [--weights universal|synthetic]

Synthetic mels N(-5, 2^2) clipped to [-11.5, 2], int16 output on the device, CUDA events on the
launching stream, 2 warm-ups per size.  Prints one JSON line per batch size: mel-frames/s, ms per pass, the achieved
fraction of the measured tensor peak on the vocoder's 614.1 MFLOP per mel frame, and the workspace size (B = 256 at
L = 1024 needs 17 GB: fits one GPU's 180 GB without time tiling).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmtts_b200 import synthetic  # noqa: E402
from cmtts_b200.config import HifiGanSpec  # noqa: E402
from cmtts_b200.vocoder import Generator  # noqa: E402

FLOP_PER_FRAME = 614_105_088.0          # probed with forward hooks on the reference generator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--weights", default="universal", choices=["universal", "synthetic"])
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    hs = HifiGanSpec()
    real = os.path.join(ROOT, "oracle", "_ref", "hifigan", "generator_universal.pth.tar")
    if args.weights == "universal" and os.path.isfile(real):
        sd = torch.load(real, map_location="cpu", weights_only=True)["generator"]
        wname = "generator_universal.pth.tar"
    else:
        sd = synthetic.make_hifigan_checkpoint(hs, seed=7)["generator"]
        wname = "synthetic (seed 7)"
    voc = Generator(hspec=hs, precision=args.precision).load_state_dict(sd).to(dev)
    peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        peak = float(mp.get("bf16_tflops_sustained") or mp.get("bf16_tflops") or 0) or None
    except Exception:
        pass
    g = torch.Generator().manual_seed(5)
    for B in [int(b) for b in args.batches.split(",")]:
        mel = (torch.randn(B, args.frames, hs.n_mels, generator=g) * 2.0 - 5.0).clamp_(-11.5, 2.0).to(dev)
        for _ in range(2):
            voc.run(mel, want_float=False, want_int16=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.reps):
            voc.run(mel, want_float=False, want_int16=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        frames = B * args.frames
        tflops = frames * FLOP_PER_FRAME / (ms * 1e-3) / 1e12
        line = {"workload": f"HiFi-GAN V1 only, B={B}, {hs.n_mels}x{args.frames} mel -> int16 wav", "weights": wname,
                "precision": args.precision, "ms": ms, "mel_frames_per_sec": frames / (ms * 1e-3),
                "audio_seconds_per_sec": frames * 256 / 22050 / (ms * 1e-3), "tflops": tflops,
                "frac_of_measured_tensor_peak": (tflops / peak) if peak else None,
                "workspace_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        print(json.dumps(line))
        sys.stdout.flush()
        del mel
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
