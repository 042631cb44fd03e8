"""In-process A/B timing of kernel variants at C2 size (run on the GPU box).

The SM clock of these boxes moves between 1.5 and 1.95 GHz under the software power cap, run to run and within a
run, so two `bench.py` processes cannot be compared at the 3 % level.  This tool flips the library's experiment
switches (cmtts_debug_set: CMTTS_UMMA_DBG bits / PDL) between back-to-back repetitions of the SAME stage in ONE
process and prints the median CUDA-event time per variant.

    python tools/ab_switch.py [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmtts_b200 import _lib, synthetic  # noqa: E402
from cmtts_b200.config import HifiGanSpec, ModelSpec  # noqa: E402
from cmtts_b200.model import CMTotalTTS  # noqa: E402
from cmtts_b200.vocoder import Generator  # noqa: E402

DEV = "cuda:0"
VARIANTS = [("gate+pdl", 0, 1), ("gate", 0, 0), ("general+pdl", 128, 1), ("general", 128, 0)]


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    lib = _lib.load()
    spec = ModelSpec.preset("LJSpeech")
    m = CMTotalTTS(spec=spec, precision="tc").load_state_dict(synthetic.make_acoustic_state_dict(spec, 0)).to(DEV)
    b = synthetic.make_batch(spec, 32, 80, 115, seed=1234)
    out = m.dpen(b["texts"], b["src_lens"], None)
    B, L, _ = out["cond"].shape
    x = torch.randn(B, 1, L, 80, device=DEV) * 80
    steps = m.prepare_steps(torch.full((B,), 1095.5), None)
    ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=7)
    voc = Generator(hspec=HifiGanSpec(), precision="tc").load_state_dict(ck["generator"]).to(DEV)
    mel = synthetic.make_mels(B, 80, L, seed=1).transpose(1, 2).contiguous().to(DEV)

    def dn():
        m.denoise_step(x, out["cond"], steps, 0.0125, 0.5, 0.0)

    def vc():
        voc.run(mel, want_float=False, want_int16=True)

    for name, fn, n in [("denoiser step (ms)", dn, 4), ("vocoder (ms)", vc, 2)]:
        res = {v[0]: [] for v in VARIANTS}
        for _ in range(2):
            fn()
        for _ in range(reps):
            for tag, dbg, pdl in VARIANTS:
                lib.cmtts_debug_set(dbg, pdl)
                fn()                                   # settle (first launch after a switch)
                res[tag].append(timed(fn, n))
        print(f"== {name}, L={L}: median / min over {reps} interleaved repetitions")
        for tag, _, _ in VARIANTS:
            v = sorted(res[tag])
            print(f"  {tag:14s} {v[len(v) // 2]:8.3f} {v[0]:8.3f}")
        sys.stdout.flush()
    lib.cmtts_debug_set(-1, -1)


if __name__ == "__main__":
    main()
