"""Where and how is the halo kernel wrong when an MRF partial sum is added? (diagnostics for gpurun)"""
import os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]]
from tools.umma_check import umma, DEV  # noqa: E402

B, L, Cc, k, dil = 6, 15104, 128, 3, 1
g = torch.Generator().manual_seed(1)
a = torch.randn(B, L, Cc, generator=g).half().to(DEV)
w = (torch.randn(k * Cc, Cc, generator=g) / (Cc * k) ** 0.5).half().to(DEV)
bias = torch.randn(Cc, generator=g).to(DEV)
r = torch.randn(B, L, Cc, generator=g).half().to(DEV)
sm = torch.randn(B, L, Cc, generator=g).half().to(DEV)
shifts = [(i - (k - 1) // 2) * dil for i in range(k)]
wk = w.view(k, Cc, Cc).float()                                  # [tap][N][Cin]
conv = F.conv1d(a.float().transpose(1, 2), wk.permute(1, 2, 0).contiguous(), bias, padding=(k - 1) // 2 * dil, dilation=dil).transpose(1, 2)
rr = r.float(); rr = torch.where(rr > 0, rr, rr * 10.0)
for use_res, use_sum in [(True, False), (False, True), (True, True)]:
    ref = conv + (rr if use_res else 0) + (sm.float() if use_sum else 0)
    ref = torch.where(ref > 0, ref, ref * 0.1)
    ob = torch.full((B, L, Cc), 7.0, dtype=torch.float16, device=DEV)
    umma(a, w, bias, shifts, Cc, res=r if use_res else None, res_inv=10.0, out_buf=ob, sum_h=sm if use_sum else None, out_slope=0.1)
    err = (ob.float() - ref).abs()
    bad = err > 0.05
    print(f"res={use_res} sum={use_sum}: max err {err.max().item():.3f}, bad elements {int(bad.sum())} of {bad.numel()}", flush=True)
    if bad.any():
        idx = bad.nonzero()
        tiles = (idx[:, 0] * ((L + 127) // 128) + idx[:, 1] // 128)
        rounds = sorted(set((tiles // 148).tolist()))
        print("   rounds with errors:", rounds[:10], " rows-in-tile min/max:", int((idx[:, 1] % 128).min()), int((idx[:, 1] % 128).max()),
              " cols min/max:", int(idx[:, 2].min()), int(idx[:, 2].max()))
        # classify: does the wrong value equal the result WITHOUT the sum, or with a sum row from elsewhere?
        ref_nosum = conv + (rr if use_res else 0); ref_nosum = torch.where(ref_nosum > 0, ref_nosum, ref_nosum * 0.1)
        e2 = (ob.float() - ref_nosum).abs()
        print("   of the bad elements, matching 'sum missing':", int((e2[bad] < 0.05).sum()))
        b0, r0, c0 = idx[0].tolist()
        print("   first bad (b,row,col):", (b0, r0, c0), "got", ob[b0, r0, c0:c0 + 4].tolist(), "want", ref[b0, r0, c0:c0 + 4].tolist(),
              "nosum", ref_nosum[b0, r0, c0:c0 + 4].tolist())
        # per (row%32==?, col block) histogram
        q = ((idx[:, 1] % 128) // 32); hcol = idx[:, 2] // 64
        print("   errors by lane-quarter:", [int((q == i).sum()) for i in range(4)], "by column half:", [int((hcol == i).sum()) for i in range(2)])
