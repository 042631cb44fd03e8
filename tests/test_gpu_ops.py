"""Single-op checks of the CUDA kernels against plain torch fp32 (on the device or CPU)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from cmtts_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def conv1d_cl(x, w_pk, bias=None, shifts=(0,), pre_slope=None, act=0, alpha=1.0, beta=1.0, res=None, res_scale=1.0,
              addvec=None, out_scale=1.0, lens=None, accumulate_into=None, act_slope=0.0):
    lib = _lib.load()
    B, M, Cin = x.shape
    taps, _, N = w_pk.shape
    nout = N // 2 if act == 5 else N
    out = accumulate_into if accumulate_into is not None else torch.empty(B, M, nout, device=x.device)
    d = _lib.ConvDesc(B=B, M=M, Lin=M, Cin=Cin, N=N, taps=taps, x_ld=Cin, out_ld=nout, res_ld=nout,
                      x_bstride=M * Cin, out_bstride=M * nout, res_bstride=M * nout,
                      addvec_bstride=N, pre_lrelu=int(pre_slope is not None), pre_slope=pre_slope or 0.0,
                      alpha=alpha, beta=beta, act=act, act_slope=act_slope, res_scale=res_scale, out_scale=out_scale,
                      accumulate=int(accumulate_into is not None))
    for i, s in enumerate(shifts):
        d.shift[i] = s
    _lib.check(lib.cmtts_conv1d(C.byref(d), _lib.ptr(x), _lib.ptr(w_pk), _lib.ptr(bias), _lib.ptr(addvec),
                                _lib.ptr(res), _lib.ptr(lens), _lib.ptr(out), _lib.stream_ptr()), "conv1d")
    return out


@pytest.mark.parametrize("B,L,Cin,Cout,k,dil", [(2, 37, 32, 64, 3, 1), (1, 300, 80, 512, 7, 1), (3, 129, 64, 64, 11, 5),
                                               (2, 50, 256, 80, 1, 1), (1, 200, 128, 32, 7, 3), (2, 1, 16, 4, 1, 1)])
def test_conv1d_matches_torch(B, L, Cin, Cout, k, dil):
    g = torch.Generator().manual_seed(k * 100 + Cin)
    x = torch.randn(B, L, Cin, generator=g).to(DEV)
    w = (torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5).to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    # reference on the CPU: cuDNN on the device may silently use TF32
    ref = F.conv1d(F.leaky_relu(x.cpu(), 0.1).transpose(1, 2), w.cpu(), b.cpu(), dilation=dil,
                   padding=(k - 1) // 2 * dil).transpose(1, 2).to(DEV)
    shifts = [(i - (k - 1) // 2) * dil for i in range(k)]
    out = conv1d_cl(x, w.permute(2, 1, 0).contiguous(), b, shifts, pre_slope=0.1)
    torch.cuda.synchronize()
    assert (out - ref).abs().max().item() <= 2e-5


def test_conv1d_epilogues():
    g = torch.Generator().manual_seed(1)
    B, L, Cin, N = 2, 70, 64, 128
    x = torch.randn(B, L, Cin, generator=g).to(DEV)
    w = (torch.randn(1, Cin, N, generator=g) / 8).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(B, L, N, generator=g).to(DEV)
    vec = torch.randn(B, N, generator=g).to(DEV)
    lens = torch.tensor([70, 33], device=DEV)
    base = (x.cpu() @ w[0].cpu()).to(DEV)   # CPU fp32 reference
    # gelu((acc + b) * beta)
    out = conv1d_cl(x, w, b, act=2, beta=1 / 3)
    assert (out - F.gelu((base + b) / 3)).abs().max() <= 1e-5
    # residual + addvec + out_scale + row mask
    out = conv1d_cl(x, w, b, res=res, addvec=vec, out_scale=0.5, lens=lens, alpha=2.0)
    ref = (base * 2 + b + vec[:, None] + res) * 0.5
    ref[1, 33:] = 0
    assert (out - ref).abs().max() <= 1e-5
    # accumulate
    acc = res.clone()
    conv1d_cl(x, w, b, accumulate_into=acc)
    assert (acc - (res + base + b)).abs().max() <= 1e-5
    # gated: tile = 64 gates | 64 filters
    out = conv1d_cl(x, w, b, act=5)
    y = base + b
    ref = torch.sigmoid(y[..., :64]) * torch.tanh(y[..., 64:])
    assert out.shape == (B, L, 64) and (out - ref).abs().max() <= 1e-5
    torch.cuda.synchronize()


def test_layernorm_and_attention():
    lib = _lib.load()
    g = torch.Generator().manual_seed(2)
    B, T, Cc, H = 3, 45, 256, 2
    x = torch.randn(B, T, Cc, generator=g).to(DEV)
    w = torch.randn(Cc, generator=g).to(DEV); b = torch.randn(Cc, generator=g).to(DEV)
    lens = torch.tensor([45, 7, 1], device=DEV)
    out = torch.empty_like(x)
    _lib.check(lib.cmtts_layernorm(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), 1e-12, _lib.ptr(out), B, T, Cc, None, _lib.stream_ptr()), "ln")
    assert (out - F.layer_norm(x, (Cc,), w, b, 1e-12)).abs().max() <= 1e-5
    qkv = torch.randn(B, T, 3 * Cc, generator=g).to(DEV)
    att = torch.empty(B, T, Cc, device=DEV)
    _lib.check(lib.cmtts_attention(_lib.ptr(qkv), _lib.ptr(lens), _lib.ptr(att), B, T, Cc, H, _lib.stream_ptr()), "attn")
    q, k, v = qkv.chunk(3, -1)
    d = Cc // H
    def heads(t): return t.view(B, T, H, d).transpose(1, 2)
    q, k, v = q.cpu(), k.cpu(), v.cpu()
    sc = (heads(q) * d ** -0.5) @ heads(k).transpose(-1, -2)
    mask = torch.arange(T)[None, :] >= lens.cpu()[:, None]
    sc = sc.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(sc, -1) @ heads(v)).transpose(1, 2).reshape(B, T, Cc).to(DEV)
    torch.cuda.synchronize()
    assert (att - ref).abs().max() <= 1e-5


def test_length_regulator_bit_exact_including_zero_durations():
    from oracle import cmtts_oracle as O
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    B, T, Cc = 4, 33, 256
    x = torch.randn(B, T, Cc, generator=g)
    src_lens = torch.tensor([33, 20, 1, 5])
    log_d = torch.randn(B, T, generator=g) * 0.8 + 1.2
    log_d[0, 3] = -5.0                      # rounds to zero frames
    log_d[3, :5] = -5.0                     # whole utterance empty -> mel_len 0
    log_d = log_d * (torch.arange(T)[None] < src_lens[:, None])
    ref_d = O.round_durations(log_d)
    ref_x, ref_len = O.length_regulate(x, ref_d, None)
    ref_m2p = O.dur_to_mel2ph(ref_d, torch.arange(T)[None] >= src_lens[:, None])
    L = int(ref_len.max())
    dl = log_d.to(DEV); xd = x.to(DEV); sl = src_lens.to(DEV)
    d_r = torch.empty(B, T, device=DEV); cs = torch.empty(B, 2, T, dtype=torch.int64, device=DEV)
    ml = torch.empty(B, dtype=torch.int64, device=DEV)
    _lib.check(lib.cmtts_round_durations(_lib.ptr(dl), 1.0, _lib.ptr(sl), _lib.ptr(d_r), _lib.ptr(cs), _lib.ptr(ml), B, T, _lib.stream_ptr()), "round")
    out = torch.empty(B, L, Cc, device=DEV); m2p = torch.empty(B, L, dtype=torch.int64, device=DEV)
    _lib.check(lib.cmtts_length_regulate(_lib.ptr(xd), _lib.ptr(cs), _lib.ptr(ml), _lib.ptr(out), _lib.ptr(m2p), B, T, L, Cc, _lib.stream_ptr()), "lr")
    torch.cuda.synchronize()
    assert torch.equal(d_r.cpu(), ref_d)
    assert torch.equal(ml.cpu(), ref_len) and int(ml[3]) == 0
    assert torch.equal(out.cpu(), ref_x)                     # gathered rows bitwise equal
    assert torch.equal(m2p.cpu()[:, : ref_m2p.shape[1]], ref_m2p)


# ---- size-independent properties at sizes where every CTA processes several tiles (persistent-loop hazards only show
# ---- up from the second round on: a shared-memory residual ring once corrupted exactly those tiles)
@pytest.mark.gpu
def test_halo_conv_residual_plus_partial_sum_large():
    import torch.nn.functional as F
    from tools.umma_check import umma
    B, L, Cc, k, dil = 6, 15104, 128, 3, 1
    g = torch.Generator().manual_seed(1)
    a = torch.randn(B, L, Cc, generator=g).half().to(DEV)
    w = (torch.randn(k * Cc, Cc, generator=g) / (Cc * k) ** 0.5).half().to(DEV)
    bias = torch.randn(Cc, generator=g).to(DEV)
    r = torch.randn(B, L, Cc, generator=g).half().to(DEV)
    sm = torch.randn(B, L, Cc, generator=g).half().to(DEV)
    shifts = [(i - (k - 1) // 2) * dil for i in range(k)]
    conv = F.conv1d(a.float().transpose(1, 2), w.view(k, Cc, Cc).float().permute(1, 2, 0).contiguous(), bias,
                    padding=(k - 1) // 2 * dil, dilation=dil).transpose(1, 2)
    rr = r.float()
    ref = conv + torch.where(rr > 0, rr, rr * 10.0) + sm.float()
    ref = torch.where(ref > 0, ref, ref * 0.1)
    outs = []
    for _ in range(2):
        ob = torch.full((B, L, Cc), 7.0, dtype=torch.float16, device=DEV)
        umma(a, w, bias, shifts, Cc, res=r, res_inv=10.0, out_buf=ob, sum_h=sm, out_slope=0.1)
        outs.append(ob)
    assert torch.equal(outs[0], outs[1])                                   # repeatable
    assert float((outs[0].float() - ref).abs().max()) <= 2e-2              # fp16 output of values up to ~10


@pytest.mark.gpu
@pytest.mark.parametrize("B,L", [(6, 236), (8, 801)])
def test_vocoder_is_deterministic_and_batch_invariant(B, L):
    from cmtts_b200 import synthetic
    from cmtts_b200.config import HifiGanSpec
    from cmtts_b200.vocoder import Generator
    ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=7)
    voc = Generator(hspec=HifiGanSpec(), precision="tc").load_state_dict(ck["generator"]).to(DEV)
    mel = synthetic.make_mels(B, 80, L, seed=1).transpose(1, 2).contiguous().to(DEV)
    full = voc.run(mel, want_float=True, want_int16=True)
    full = (full[0].clone(), full[1].clone())
    again = voc.run(mel, want_float=True, want_int16=True)
    assert torch.equal(full[0], again[0]) and torch.equal(full[1], again[1])
    h = B // 2
    # utterances are independent: a sub-batch must reproduce its rows of the batched run bit for bit
    lo = voc.run(mel[:h].contiguous(), want_float=True, want_int16=True)
    assert torch.equal(full[0][:h], lo[0]) and torch.equal(full[1][:h], lo[1])
    hi = voc.run(mel[h:].contiguous(), want_float=True, want_int16=True)
    assert torch.equal(full[0][h:], hi[0]) and torch.equal(full[1][h:], hi[1])


@pytest.mark.gpu
@pytest.mark.parametrize("B,L", [(3, 40), (4, 801)])
def test_paired_row_resblock_kernel_against_the_plain_one(B, L):
    """The 32-channel level runs on the two-time-steps-per-row ResBlock kernel (csrc/umma_resblock.cu, MODE 1 / 2); bit
    1024 of the debug word sends it back to the plain C = 32 kernel.  Same fp16 operands, fp32 accumulation in another
    order: the two wavs agree to fp16-storage noise, and each is within the vocoder bound of the fp32 FFMA path."""
    from cmtts_b200 import _lib, synthetic
    from cmtts_b200.config import HifiGanSpec
    from cmtts_b200.vocoder import Generator
    lib = _lib.load()
    ck = synthetic.make_hifigan_checkpoint(HifiGanSpec(), seed=7)
    voc = Generator(hspec=HifiGanSpec(), precision="tc").load_state_dict(ck["generator"]).to(DEV)
    ref = Generator(hspec=HifiGanSpec(), precision="fp32").load_state_dict(ck["generator"]).to(DEV)
    mel = synthetic.make_mels(B, 80, L, seed=3).transpose(1, 2).contiguous().to(DEV)
    n0 = lib.cmtts_launch_count()
    paired = voc.run(mel, want_float=True, want_int16=False)[0].clone()
    n_paired = lib.cmtts_launch_count() - n0
    try:
        lib.cmtts_debug_set(1024, -1)
        plain = voc.run(mel, want_float=True, want_int16=False)[0].clone()
    finally:
        lib.cmtts_debug_set(-1, -1)
    exact = ref.run(mel, want_float=True, want_int16=False)[0]
    torch.cuda.synchronize()
    d_pp = float((paired - plain).abs().max())
    d_p = float((paired - exact).abs().max())
    d_q = float((plain - exact).abs().max())
    print(f"paired-row ResBlock kernel B={B} L={L}: paired vs plain {d_pp:.2e}, paired vs fp32 {d_p:.2e}, plain vs fp32 {d_q:.2e} "
          f"({n_paired} launches)")
    assert torch.isfinite(paired).all()
    assert d_pp <= 2e-3 and d_p <= 4e-3 and d_p <= 2.0 * d_q + 1e-4
