"""Is the tensor-core vocoder deterministic and batch-invariant?  (diagnostics for gpurun)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmtts_b200 import synthetic
from cmtts_b200.config import HifiGanSpec
from cmtts_b200.vocoder import Generator

ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=7)
voc = Generator(hspec=HifiGanSpec(), precision="tc").load_state_dict(ck["generator"]).to("cuda:0")
for B, L in [(6, 236), (4, 276), (8, 801)]:
    mel = synthetic.make_mels(B, 80, L, seed=1).transpose(1, 2).contiguous().to("cuda:0")
    a = voc.run(mel, want_float=True, want_int16=True)
    a = (a[0].clone(), a[1].clone())
    b = voc.run(mel, want_float=True, want_int16=True)
    b = (b[0].clone(), b[1].clone())
    h = B // 2
    c = voc.run(mel[:h].contiguous(), want_float=True, want_int16=True)
    c = (c[0].clone(), c[1].clone())
    d = voc.run(mel[h:].contiguous(), want_float=True, want_int16=True)
    torch.cuda.synchronize()
    print(f"B={B} L={L}: repeat equal {torch.equal(a[0], b[0])}; first half vs batched: float max diff "
          f"{(a[0][:h] - c[0]).abs().max().item():.3e}, n diff {(a[0][:h] != c[0]).sum().item()} ; second half: "
          f"{(a[0][h:] - d[0]).abs().max().item():.3e}, n diff {(a[0][h:] != d[0]).sum().item()} of {c[0].numel()}", flush=True)
    if not torch.equal(a[0][:h], c[0]):
        nz = (a[0][:h] != c[0]).nonzero()
        print("   first diffs (row, sample):", nz[:5].tolist(), " last:", nz[-3:].tolist(), " rows:", sorted(set(nz[:, 0].tolist())))
# level by level: env CMTTS_UMMA_DBG=4 (unfused), =2 (no halo)
