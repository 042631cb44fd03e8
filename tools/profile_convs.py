"""Launch the dominant HiFi-GAN conv shapes at bench size (for `ncu --set full`)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]]
from tools.umma_check import case_time  # noqa: E402
B, F = 32, 793
case_time(B, F * 64, 128, 11, 5, True, reps=2)    # level 1, k=11 d=5, with residual
case_time(B, F * 128, 64, 7, 3, True, reps=2)     # level 2
case_time(B, F * 256, 32, 3, 1, True, reps=2)     # level 3
case_time(B, F * 8, 256, 7, 1, True, reps=2)      # level 0 (general kernel)
