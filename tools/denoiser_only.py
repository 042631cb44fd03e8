"""One denoiser evaluation at C2 size (for ncu launch lists)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmtts_b200 import synthetic
from cmtts_b200.config import ModelSpec
from cmtts_b200.model import CMTotalTTS
spec = ModelSpec.preset("LJSpeech")
m = CMTotalTTS(spec=spec, precision=sys.argv[1] if len(sys.argv) > 1 else "tc").load_state_dict(synthetic.make_acoustic_state_dict(spec, 0)).to("cuda:0")
b = synthetic.make_batch(spec, 32, 80, 115, seed=1234)
out = m.dpen(b["texts"], b["src_lens"], None)
B, L, _ = out["cond"].shape
x = torch.randn(B, 1, L, 80, device="cuda:0") * 80
steps = m.prepare_steps(torch.full((B,), 1095.5), None)
for _ in range(2):
    y = m.denoise_step(x, out["cond"], steps, 0.0125, 0.5, 0.0)
torch.cuda.synchronize()
print("ok", L)
