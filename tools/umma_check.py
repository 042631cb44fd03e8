"""Device-side check of the tcgen05 conv kernel against torch-CPU fp32 (diagnostics for gpurun)."""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmtts_b200 import _lib  # noqa: E402

DEV = "cuda:0"
lib = _lib.load()


def umma(a_hi, w_hi, bias, shifts, N, *, a_lo=None, w_lo=None, epi=0, alpha=1.0, res=None, res_inv=1.0, sum_h=None,
         out_slope=1.0, out_lo=False, addvec=None, x_f32=None, skip=None, skip_acc=0, out_scale=1.0, out_ch=None,
         sync=True, out_buf=None):
    B, L, Cin = a_hi.shape
    out_ch = out_ch or N
    out_h = out_buf if out_buf is not None else (torch.empty(B, L, out_ch, dtype=torch.float16, device=DEV) if epi != 3 else None)
    o_lo = torch.empty(B, L, out_ch, dtype=torch.float16, device=DEV) if out_lo else None
    d = _lib.UmmaDesc(B=B, M=L, Lin=L, N=N, Cin=Cin, taps=len(shifts), split=int(a_lo is not None), epi=epi,
                      a_ld=Cin, res_ld=out_ch, out_ld=out_ch, x_ld=(x_f32.shape[-1] if x_f32 is not None else 0),
                      a_bstride=L * Cin, res_bstride=L * out_ch, out_bstride=L * out_ch,
                      x_bstride=(L * x_f32.shape[-1] if x_f32 is not None else 0),
                      addvec_bstride=(addvec.shape[-1] if addvec is not None else 0),
                      alpha=alpha, res_inv_slope=res_inv, out_slope=out_slope, out_scale=out_scale, skip_accumulate=skip_acc)
    for i, s in enumerate(shifts):
        d.shift[i] = s
    p = _lib.ptr
    _lib.check(lib.cmtts_umma_conv1d(C.byref(d), p(a_hi), p(a_lo), p(w_hi), p(w_lo), p(bias), p(res), p(sum_h), p(out_h),
                                     p(o_lo), p(addvec), p(x_f32), p(skip), _lib.stream_ptr()), "umma_conv1d")
    if sync:
        torch.cuda.synchronize()
    return out_h, o_lo


def ref_conv(a, w_kNC, bias, shifts):
    """a (B,L,Cin) fp32 cpu; w [k][N][Cin]; out (B,L,N)."""
    B, L, Cin = a.shape
    out = torch.zeros(B, L, w_kNC.shape[1])
    for i, s in enumerate(shifts):
        sh = torch.zeros_like(a)
        if s < 0:
            sh[:, -s:] = a[:, :L + s] if L + s > 0 else 0
        elif s > 0:
            sh[:, :L - s] = a[:, s:] if L - s > 0 else 0
        else:
            sh = a
        out += sh @ w_kNC[i].t()
    return out + bias


def case_plain(B, L, Cin, N, k, dil, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, L, Cin, generator=g).half()
    w = (torch.randn(k, N, Cin, generator=g) / (Cin * k) ** 0.5).half()
    bias = torch.randn(N, generator=g)
    shifts = [(i - (k - 1) // 2) * dil for i in range(k)]
    ref = ref_conv(a.float(), w.float(), bias, shifts)
    out, _ = umma(a.to(DEV), w.reshape(k * N, Cin).contiguous().to(DEV), bias.to(DEV), shifts, N)
    e = (out.float().cpu() - ref).abs().max().item()
    print(f"plain B={B} L={L} Cin={Cin} N={N} k={k} dil={dil}: max err {e:.3e} (|ref| max {ref.abs().max():.2f})", flush=True)
    return e


def case_voc_epilogue():
    g = torch.Generator().manual_seed(5)
    B, L, Cin, N, k = 2, 333, 64, 64, 3
    a = torch.randn(B, L, Cin, generator=g).half()
    w = (torch.randn(k, N, Cin, generator=g) / (Cin * k) ** 0.5).half()
    bias = torch.randn(N, generator=g)
    y = torch.randn(B, L, N, generator=g)
    res_act = F.leaky_relu(y, 0.1).half()
    part = torch.randn(B, L, N, generator=g).half()
    shifts = [-1, 0, 1]
    conv = ref_conv(a.float(), w.float(), bias * 0, shifts)
    yrec = torch.where(res_act.float() > 0, res_act.float(), res_act.float() * 10)
    ref = F.leaky_relu(conv * 0.5 + bias + yrec + part.float(), 0.01)
    out, _ = umma(a.to(DEV), w.reshape(k * N, Cin).contiguous().to(DEV), bias.to(DEV), shifts, N, alpha=0.5,
                  res=res_act.to(DEV), res_inv=10.0, sum_h=part.to(DEV), out_slope=0.01)
    print(f"voc epilogue: max err {(out.float().cpu() - ref).abs().max().item():.3e}", flush=True)


def split(x):
    hi = x.half()
    return hi, (x - hi.float()).half()


def wsplit(w):
    from cmtts_b200.weights import split_f16
    return split_f16(w)


W_INV = 1.0 / 1024.0


def case_split2():
    g = torch.Generator().manual_seed(8)
    B, L, Cc = 2, 200, 256
    y = torch.randn(B, L, Cc, generator=g)
    x = torch.randn(B, L, Cc, generator=g)
    vec = torch.randn(B, Cc, generator=g)
    from cmtts_b200.weights import gate_permutation
    # DN_COND: y = Wc cond + b + vec[b] + x
    cond = torch.randn(B, L, Cc, generator=g)
    w = torch.randn(1, Cc, Cc, generator=g) / 16
    b = torch.randn(Cc, generator=g)
    ref = ref_conv(cond, w, b, [0]) + vec[:, None] + x
    ch, cl = split(cond); wh, wl = wsplit(w.reshape(Cc, Cc))
    yh_, yl_ = umma(ch.to(DEV), wh.to(DEV), b.to(DEV), [0], Cc, a_lo=cl.to(DEV), w_lo=wl.to(DEV), epi=1, out_lo=True,
                    addvec=vec.to(DEV), x_f32=x.to(DEV), alpha=W_INV)
    got = yh_.float().cpu() + yl_.float().cpu()
    print(f"split DN_COND: max err {(got - ref).abs().max().item():.3e}", flush=True)
    # DN_GATE: k3 conv 256 -> 512, gate/filter pairs
    w = torch.randn(3, 2 * Cc, Cc, generator=g) / (3 * Cc) ** 0.5
    b = torch.randn(2 * Cc, generator=g) * 0.1
    conv = ref_conv(y, w, b, [-1, 0, 1])
    ref = torch.sigmoid(conv[..., :Cc]) * torch.tanh(conv[..., Cc:])
    perm = gate_permutation(Cc)
    wp, bp = w[:, perm], b[perm]
    yh, yl = split(y); wh, wl = wsplit(wp.reshape(3 * 2 * Cc, Cc))
    gh, gl = umma(yh.to(DEV), wh.to(DEV), bp.to(DEV), [-1, 0, 1], 2 * Cc, a_lo=yl.to(DEV), w_lo=wl.to(DEV), epi=2,
                  out_lo=True, out_ch=Cc, alpha=W_INV)
    got = gh.float().cpu() + gl.float().cpu()
    print(f"split DN_GATE: max err {(got - ref).abs().max().item():.3e}", flush=True)
    # DN_OUT
    w = torch.randn(1, 2 * Cc, Cc, generator=g) / 16
    b = torch.randn(2 * Cc, generator=g) * 0.1
    gact = ref
    conv = ref_conv(gact, w, b, [0])
    ref_x = (conv[..., :Cc] + vec[:, None] + x) * 0.70710678
    skip0 = torch.randn(B, L, Cc, generator=g)
    ref_s = skip0 + conv[..., Cc:]
    ah, al = split(gact); wh, wl = wsplit(w.reshape(2 * Cc, Cc))
    xd = x.to(DEV).clone(); sd = skip0.to(DEV).clone()
    umma(ah.to(DEV), wh.to(DEV), b.to(DEV), [0], 2 * Cc, a_lo=al.to(DEV), w_lo=wl.to(DEV), epi=3, addvec=vec.to(DEV),
         x_f32=xd, skip=sd, skip_acc=1, out_scale=0.70710678, out_ch=Cc, alpha=W_INV)
    print(f"split DN_OUT: x err {(xd.cpu() - ref_x).abs().max().item():.3e} skip err {(sd.cpu() - ref_s).abs().max().item():.3e}", flush=True)


def case_time(B, L, Cc, k, dil, res=False, reps=20):
    g = torch.Generator().manual_seed(1)
    a = (torch.randn(B, L, Cc, generator=g)).half().to(DEV)
    w = (torch.randn(k * Cc, Cc, generator=g) / (Cc * k) ** 0.5).half().to(DEV)
    bias = torch.randn(Cc, generator=g).to(DEV)
    r = a.clone() if res else None
    shifts = [(i - (k - 1) // 2) * dil for i in range(k)]
    ob = torch.empty(B, L, Cc, dtype=torch.float16, device=DEV)
    umma(a, w, bias, shifts, Cc, res=r, res_inv=10.0, out_buf=ob)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        umma(a, w, bias, shifts, Cc, res=r, res_inv=10.0, sync=False, out_buf=ob)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * B * L * Cc * Cc * k
    byts = B * L * Cc * 2 * (3 if res else 2)
    print(f"time B={B} L={L} C={Cc} k={k} dil={dil} res={res}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s  {byts/ms/1e6:7.1f} GB/s (algorithmic)", flush=True)


def case_accum():
    """How accurate is the hi/lo tensor-core GEMM over a long K (tcgen05 fp32 accumulation) vs fp32 FFMA?"""
    import ctypes as C
    g = torch.Generator().manual_seed(11)
    for taps in (1, 3, 9):
        B, L, Cc, N = 2, 256, 256, 256
        a = torch.randn(B, L, Cc, generator=g)
        w = torch.randn(taps, N, Cc, generator=g) / (Cc * taps) ** 0.5
        bias = torch.zeros(N)
        shifts = [i - (taps - 1) // 2 for i in range(taps)]
        ref = ref_conv(a.double(), w.double(), bias.double(), shifts)
        ah, al = split(a); wh, wl = wsplit(w.reshape(taps * N, Cc))
        # generic fp32 epilogue (epi 4) is not exposed through umma(); use DN_COND with zero x / addvec
        zx = torch.zeros(B, L, N, device=DEV); zv = torch.zeros(B, N, device=DEV)
        yh, yl = umma(ah.to(DEV), wh.to(DEV), bias.to(DEV), shifts, N, a_lo=al.to(DEV), w_lo=wl.to(DEV), epi=1, out_lo=True,
                      addvec=zv, x_f32=zx, alpha=W_INV)
        got = yh.double().cpu() + yl.double().cpu()
        # exact product of the rounded operands (isolates accumulation error from operand rounding)
        a_r = ah.double() + al.double(); w_r = (wh.double() + wl.double()).reshape(taps, N, Cc) / 1024.0
        ref_r = ref_conv(a_r, w_r, bias.double(), shifts)
        ref32 = ref_conv(a, w, bias, shifts)
        print(f"accum taps={taps} K={taps*Cc}: tc vs fp64 {(got-ref).abs().max():.3e} | tc vs exact-of-rounded-operands {(got-ref_r).abs().max():.3e}"
              f" | operand rounding alone {(ref_r-ref).abs().max():.3e} | torch fp32 vs fp64 {(ref32.double()-ref).abs().max():.3e} | hi/lo output quantisation ~{float(ref.abs().max())*2**-22:.1e}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "plain"):
        case_plain(1, 128, 64, 64, 1, 1)
        case_plain(2, 300, 64, 64, 3, 1)
        case_plain(2, 517, 128, 128, 7, 3)
        case_plain(1, 1000, 256, 256, 11, 5)
        case_plain(3, 260, 32, 32, 3, 1)
        case_plain(2, 700, 32, 32, 11, 5)
        case_plain(2, 90, 512, 2048, 3, 1)
        case_plain(1, 40000, 64, 64, 7, 1)
        case_plain(2, 333, 128, 128, 3, 1)
        case_plain(2, 1000, 128, 128, 11, 5)
    if which in ("all", "epi"):
        case_voc_epilogue()
    if which in ("time",):
        for res in (False, True):
            case_time(32, 204800, 32, 3, 1, res)
            case_time(32, 204800, 32, 11, 5, res)
            case_time(32, 102400, 64, 3, 1, res)
            case_time(32, 102400, 64, 7, 3, res)
            case_time(32, 51200, 128, 3, 1, res)
            case_time(32, 51200, 128, 7, 1, res)
            case_time(32, 51200, 128, 11, 5, res)
            case_time(32, 6400, 256, 7, 1, res)
    if which in ("all", "split"):
        case_split2()
    if which in ("accum",):
        case_accum()
