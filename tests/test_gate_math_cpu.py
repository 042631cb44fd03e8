"""The branch-free gate math of cmtts_b200/csrc/umma_gate.cu (gate_fast), restated in float32 numpy.

sigmoid(g) * tanh(f) = (1 - e^{-2f}) / ((1 + e^{-g}) (1 + e^{-2f})) with clamped arguments: the kernel evaluates it with
ex2.approx / rcp.approx (2 ulp each); this test pins the FORMULA (clamps, cancellation near f = 0, saturation) against
the reference's torch.sigmoid * torch.tanh (model/blocks.py:679-681) in float64.
"""
import numpy as np


def gate_fast_f32(g, f):
    one = np.float32(1)
    a = np.clip(g, np.float32(-30), np.float32(30)) * np.float32(-1.4426950408889634)
    b = np.clip(f, np.float32(-15), np.float32(15)) * np.float32(-2.8853900817779268)
    eg = np.exp2(a).astype(np.float32)
    ef = np.exp2(b).astype(np.float32)
    return ((one - ef) * (one / ((one + eg) * (one + ef))).astype(np.float32)).astype(np.float32)


def ref(g, f):
    g = g.astype(np.float64); f = f.astype(np.float64)
    with np.errstate(over="ignore"):                 # exp(1e4) = inf -> sigmoid = 0 exactly, which is the right limit
        return (1.0 / (1.0 + np.exp(-g))) * np.tanh(f)


def test_gate_math_matches_sigmoid_tanh():
    rng = np.random.default_rng(0)
    g = rng.normal(0, 3, 1_000_000).astype(np.float32)
    f = rng.normal(0, 2, 1_000_000).astype(np.float32)
    f[:2000] = rng.normal(0, 1e-3, 2000).astype(np.float32)          # cancellation region of 1 - e^{-2f}
    err = np.abs(gate_fast_f32(g, f).astype(np.float64) - ref(g, f))
    assert err.max() <= 3e-7, err.max()


def test_gate_math_saturates_without_overflow():
    g = np.array([-1e4, -88, -30, 30, 88, 1e4, 0, 0, 50, -50], np.float32)
    f = np.array([0.5, 0.5, 0.5, 0.5, 0.5, 0.5, -1e4, 1e4, 40, -40], np.float32)
    out = gate_fast_f32(g, f)
    assert np.all(np.isfinite(out))
    assert np.abs(out.astype(np.float64) - ref(g, f)).max() <= 3e-7
