"""Launch the probed conv shapes of bench.py's roofline block at bench size (for `ncu --set full`):

    ncu --set full --clock-control none --import-source on -k regex:umma -o gpurun_out/prof_final python tools/profile_convs.py
    python tools/ncu_summary.py gpurun_out/prof_final.ncu-rep profiles/ncu_r1_final_summary.csv
"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]]
import bench  # noqa: E402
from cmtts_b200 import _lib  # noqa: E402
from tools.umma_check import case_time  # noqa: E402

lib = _lib.load()
B, F = 32, 793
orig = bench._probe
bench._probe = lambda lib_, d, ptrs, reps=10: orig(lib_, d, ptrs, reps=1)     # 3 warm-ups + 1 launch per shape
bench.kernel_probes(lib, torch.device("cuda", 0), B, F)   # level-1 C=128 k=11 d=5 + residual; denoiser k=3 gate conv (hi/lo)
case_time(B, F * 8, 256, 7, 1, True, reps=2)              # level 0 (general kernel, BN=256)
case_time(B, F * 64, 128, 3, 1, True, reps=2)             # level 1 k=3 (resident weights)
case_time(B, F * 128, 64, 11, 5, True, reps=2)            # level 2 k=11 (unfused)
torch.cuda.synchronize()
print("ok")
