"""CPU-side checks: the C-ABI library loads and exports every symbol include/cmtts_b200.h declares,
host-side packing helpers are correct, and the product package never touches the oracle."""
import os
import re

import pytest
import torch

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "cmtts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmtts_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from cmtts_b200 import build, _lib
    build.build()
    lib = _lib.load()
    assert lib.cmtts_abi_version() == 2
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cmtts_b200.h but not exported"
    # every bound prototype is declared in the header
    assert set(_lib.PROTOTYPES) <= set(syms)


def test_sass_is_sm100a():
    from cmtts_b200 import build
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    from cmtts_b200.model import CMTotalTTS
    from cmtts_b200.config import ModelSpec
    from cmtts_b200 import _lib
    m = CMTotalTTS(spec=ModelSpec.preset("LJSpeech"))
    with pytest.raises(_lib.CmttsError):
        m.to("cpu")
    with pytest.raises(_lib.CmttsError):
        _lib.ptr(torch.zeros(4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cmtts_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "/root/reference" not in txt, f


def test_conv_transpose_packing_matches_torch():
    from cmtts_b200.weights import pack_conv_transpose
    g = torch.Generator().manual_seed(0)
    for (cin, cout, k, u) in [(6, 4, 16, 8), (8, 4, 4, 2)]:
        w = torch.randn(cin, cout, k, generator=g)
        b = torch.randn(cout, generator=g)
        x = torch.randn(2, cin, 7, generator=g)
        ref = torch.nn.functional.conv_transpose1d(x, w, b, stride=u, padding=(k - u) // 2)
        pk, pb, d0 = pack_conv_transpose(w, b, u, (k - u) // 2)
        xt = x.transpose(1, 2)
        out = torch.zeros(2, 7, u * cout)
        for i in range(pk.shape[0]):
            dl = d0 + i
            sh = torch.zeros_like(xt)
            if dl < 0:
                sh[:, -dl:] = xt[:, :dl]
            elif dl > 0:
                sh[:, :-dl] = xt[:, dl:]
            else:
                sh = xt
            out += sh @ pk[i]
        out = (out + pb).view(2, 7 * u, cout).transpose(1, 2)
        assert (out - ref).abs().max() < 1e-5


def test_gate_permutation_pairs_channels():
    from cmtts_b200.weights import gate_permutation
    p = gate_permutation(256).tolist()
    assert sorted(p) == list(range(512))
    for tile in range(4):
        for i in range(64):
            assert p[tile * 128 + i] == tile * 64 + i            # gate c
            assert p[tile * 128 + 64 + i] == 256 + tile * 64 + i  # filter c


def test_sinusoid_table_matches_oracle():
    from cmtts_b200.weights import sinusoid_table
    from oracle.cmtts_oracle import sinusoid_table as ref
    assert torch.equal(sinusoid_table(300, 256), ref(300, 256))
    assert torch.equal(sinusoid_table(50, 128), ref(50, 128))


def test_sigma_plan_matches_reference_values():
    """SURVEY.md §0.5 / §8 S4 [probed]: c_skip 3.906e-5, c_out 0.49998, c_in 0.0124998, t' 1095.5067."""
    from cmtts_b200.model import KarrasDenoiser
    from cmtts_b200.sampler import get_sigmas_karras
    d = KarrasDenoiser(distillation=True)
    c_skip, c_out, c_in, t = d.scalar_plan(80.0)
    assert abs(c_skip - 3.906e-5) < 1e-8 and abs(c_out - 0.49998) < 1e-5 and abs(c_in - 0.0124998) < 1e-7
    assert abs(t - 1095.5067) < 1e-3
    s = get_sigmas_karras(2, 0.002, 80.0)
    assert abs(float(s[0]) - 79.99998474121094) < 1e-9 and float(s[2]) == 0.0


def test_custom_ops_are_registered_and_have_no_cpu_fallback():
    """north_star: kernels bound as custom ops through the C ABI.  Every op of cmtts_b200/ops.py exists under
    torch.ops.cmtts_b200 with a CUDA kernel only: CPU tensors are refused by the dispatcher, not silently computed."""
    import pytest
    import torch
    from cmtts_b200 import ops
    for name in ops.OPS:
        assert hasattr(torch.ops.cmtts_b200, name), name
        assert torch._C._dispatch_has_kernel_for_dispatch_key(f"cmtts_b200::{name}", "CUDA")
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(f"cmtts_b200::{name}", "CPU")
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.cmtts_b200.renoise(torch.zeros(4), torch.zeros(4), 1.0, 0.85)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.cmtts_b200.split_f16(torch.zeros(2, 8))
