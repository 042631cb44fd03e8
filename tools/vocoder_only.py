"""One HiFi-GAN pass at C2 size (for ncu launch lists)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmtts_b200 import synthetic
from cmtts_b200.config import HifiGanSpec
from cmtts_b200.vocoder import Generator
ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=7)
voc = Generator(hspec=HifiGanSpec(), precision=sys.argv[1] if len(sys.argv) > 1 else "tc").load_state_dict(ck["generator"]).to("cuda:0")
mel = synthetic.make_mels(32, 80, 793, seed=1).transpose(1, 2).contiguous().to("cuda:0")
for _ in range(2):
    voc.run(mel, want_float=False, want_int16=True)
torch.cuda.synchronize()
print("ok")
