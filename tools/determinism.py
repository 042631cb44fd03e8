"""Run each tcgen05 conv shape several times on the same input: bitwise repeatability (diagnostics for gpurun)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]]
from tools.umma_check import umma, DEV  # noqa: E402


def case(B, L, Cc, k, dil, res, reps=4, sum_=False):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(B, L, Cc, generator=g).half().to(DEV)
    w = (torch.randn(k * Cc, Cc, generator=g) / (Cc * k) ** 0.5).half().to(DEV)
    bias = torch.randn(Cc, generator=g).to(DEV)
    r = torch.randn(B, L, Cc, generator=g).half().to(DEV) if res else None
    sm = torch.randn(B, L, Cc, generator=g).half().to(DEV) if sum_ else None
    shifts = [(i - (k - 1) // 2) * dil for i in range(k)]
    outs = []
    for _ in range(reps):
        ob = torch.full((B, L, Cc), 7.0, dtype=torch.float16, device=DEV)
        umma(a, w, bias, shifts, Cc, res=r, res_inv=10.0, out_buf=ob, sum_h=sm, out_slope=0.1)
        outs.append(ob)
    nd = [int((outs[0] != o).sum()) for o in outs[1:]]
    md = [float((outs[0].float() - o.float()).abs().max()) for o in outs[1:]]
    bad = (outs[0] != outs[1]).nonzero()
    where = f" first diff (b,row,ch) {bad[0].tolist()} rows {sorted(set((bad[:, 1] // 128).tolist()))[:8]}" if len(bad) else ""
    print(f"B={B} L={L} C={Cc} k={k} dil={dil} res={res} sum={sum_}: differing elements {nd} max diff {md}{where}", flush=True)


if __name__ == "__main__":
    print("CMTTS_UMMA_DBG =", os.environ.get("CMTTS_UMMA_DBG"))
    for (Cc, k, dil) in [(128, 3, 1), (128, 7, 3), (128, 11, 5), (64, 11, 5), (64, 3, 1), (32, 3, 1), (32, 11, 5), (256, 3, 1)]:
        rate = {256: 8, 128: 64, 64: 128, 32: 256}[Cc]
        case(6, 236 * rate, Cc, k, dil, False)
        case(6, 236 * rate, Cc, k, dil, True)
    case(6, 236 * 64, 128, 3, 1, True, sum_=True)
