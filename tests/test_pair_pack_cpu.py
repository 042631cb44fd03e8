"""Host-side check of the pair-packed weights of the two-time-steps-per-row ResBlock kernel (csrc/umma_resblock.cu,
MODE 1 / 2): the kernel's operand addressing (units of 64-byte half rows at byte offsets (u + 1) * 64 resp.
(1 + j d) * 64 / (2 + j d) * 64 of the halo tile, weight blocks of weights.py: pair_pack_d1 / pair_pack_taps) is
replayed in torch on the CPU and compared with F.conv1d (hifigan/models.py:96-103)."""
import pytest
import torch
import torch.nn.functional as F

from cmtts_b200.weights import conv_w_nk, pair_pack_d1, pair_pack_taps


def _paired_conv(x, w, d):
    """x (L, 32) fp64, w (32, 32, k) -> (L, 32) computed the way the paired kernel does."""
    L = x.shape[0]
    k = w.shape[2]
    h = (k - 1) // 2
    hq1 = (h * d + 1) // 2                       # pair rows of halo either side (RbCfg.p1d)
    w_nk = conv_w_nk(w)                          # [k][Cout][Cin]
    # halo tile of ONE big m-tile: pair rows -hq1 .. L/2 + hq1, as 64-byte "positions" (one time step each)
    pos = torch.zeros(L + 4 * hq1 + 4, 32, dtype=x.dtype)
    pos[2 * hq1:2 * hq1 + L] = x                 # position index X <-> time 2 (R - hq1) + X for accumulator row R = 0
    out = torch.zeros(L // 2, 64, dtype=x.dtype)  # accumulator rows: [t = 2R | t = 2R + 1]
    R = torch.arange(L // 2)
    if d == 1:
        wp = pair_pack_d1(w_nk)                  # [(k+1)/2][64][64]
        for u in range(k + 1):
            a = pos[2 * R + (u + 1)]                                   # byte offset (u + 1) * 64 from the row of R
            b = wp[u >> 1][:, (u & 1) * 32:(u & 1) * 32 + 32]          # [64 N][32 K]
            out += a @ b.t()
    else:
        wp = pair_pack_taps(w_nk)                # [(k+1)/2][32][64]
        for j in range(k):
            b = wp[j >> 1][:, (j & 1) * 32:(j & 1) * 32 + 32]          # [32 N][32 K]
            out[:, :32] += pos[2 * R + (1 + j * d)] @ b.t()
            out[:, 32:] += pos[2 * R + (2 + j * d)] @ b.t()
    return out.reshape(L, 32)


@pytest.mark.parametrize("k", [3, 7, 11])
@pytest.mark.parametrize("d", [1, 3, 5])
def test_paired_addressing_matches_conv1d(k, d):
    g = torch.Generator().manual_seed(100 * k + d)
    L = 64
    x = torch.randn(L, 32, generator=g, dtype=torch.float64)
    w = torch.randn(32, 32, k, generator=g, dtype=torch.float64)
    ref = F.conv1d(x.t()[None], w, padding=(k - 1) // 2 * d, dilation=d)[0].t()
    got = _paired_conv(x, w, d)
    assert torch.allclose(got, ref, atol=1e-10), float((got - ref).abs().max())


def test_pack_shapes_and_zero_blocks():
    w = torch.randn(7, 32, 32)
    p1 = pair_pack_d1(w)
    assert p1.shape == (4, 64, 64)
    assert float(p1[0, 32:, :32].abs().max()) == 0.0          # unit 0 has no tap for the odd output
    assert float(p1[3, :32, 32:].abs().max()) == 0.0          # unit k has no tap for the even output
    assert torch.equal(p1[1, :32, :32], w[2]) and torch.equal(p1[1, 32:, :32], w[1])
    p2 = pair_pack_taps(w)
    assert p2.shape == (4, 32, 64) and torch.equal(p2[3, :, :32], w[6]) and float(p2[3, :, 32:].abs().max()) == 0.0
