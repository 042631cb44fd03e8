"""Run ONE stage of the hot path a few times at a bench config's size — the short command `ncu --set full` wants.

    python tools/stage_only.py --stage vocoder|sampler|dpen|all [--config C2] [--T 4] [--reps 2]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cmtts_b200 import synthetic  # noqa: E402
from cmtts_b200.config import ModelSpec  # noqa: E402
from cmtts_b200.sampler import karras_sample_tts, sampler_plan  # noqa: E402
from cmtts_b200.synthesize import Pipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="vocoder", choices=["vocoder", "sampler", "dpen", "all"])
    ap.add_argument("--config", default="C2", choices=sorted(bench.CONFIGS))
    ap.add_argument("--T", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    c = bench.CONFIGS[a.config]
    T = a.T or c["T"] or 1
    dev = torch.device("cuda", 0)
    spec = ModelSpec.preset(c["dataset"])
    sd = synthetic.make_acoustic_state_dict(spec, seed=0)
    hsd, _ = bench.load_hifigan(spec, False)
    pipe = Pipeline(spec, sd, hsd, dev)
    B = a.batch or c["per_gpu"]
    if a.config == "C5":
        mel = synthetic.make_mels(B, spec.n_mels, bench.C5_FRAMES, seed=99).transpose(1, 2).contiguous().to(dev)
        for _ in range(a.reps):
            pipe.vocoder.run(mel, want_float=False, want_int16=True)
        torch.cuda.synchronize()
        return
    if a.stage == "vocoder":
        # timing does not depend on the values: mels of the config's shape (L ~ 7 frames per phoneme), no acoustic pass
        L = int(c["hi"] * 6.9)
        mel = synthetic.make_mels(B, spec.n_mels, L, seed=99).transpose(1, 2).contiguous().to(dev)
        for _ in range(a.reps):
            pipe.vocoder.run(mel, want_float=False, want_int16=True)
        torch.cuda.synchronize()
        return
    b = synthetic.make_batch(spec, B, c["lo"], c["hi"], seed=1234)
    t, l = b["texts"].to(dev), b["src_lens"].to(dev)
    s = None if b["spker_embeds"] is None else b["spker_embeds"].to(dev)
    if a.stage == "all":
        for _ in range(a.reps):
            pipe(t, l, s, T=T)
        torch.cuda.synchronize()
        return
    out = pipe.model.dpen(t, l, s, None)
    if a.stage == "dpen":
        for _ in range(a.reps - 1):
            pipe.model.dpen(t, l, s, None)
        torch.cuda.synchronize()
        return
    Bn, L, _ = out["cond"].shape
    sampler, steps, ts = sampler_plan(T)
    kw = {"texts": t, "src_lens": l, "spker_embeds": s}
    mel = None
    for _ in range(a.reps if a.stage == "sampler" else 1):
        mel = karras_sample_tts(pipe.diffusion, pipe.model, (Bn, 1, L, spec.n_mels), steps=steps, model_kwargs=kw, device=dev,
                                sigma_min=spec.sigma_min, sigma_max=spec.sigma_max, sampler=sampler, ts=ts, cond_dict=out)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
