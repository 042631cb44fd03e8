"""Host logic of the tensor-core denoiser: the y-recurrence operands (cmtts_b200/weights.py,
fused_recurrence_weights) reproduce the reference's residual stack (model/blocks.py:667-686,
model/modules.py:626-634) — checked in fp64 on the CPU, no kernels involved."""
import math

import torch
import torch.nn.functional as F

from cmtts_b200 import synthetic
from cmtts_b200.config import ModelSpec
from cmtts_b200.weights import TC_W_SCALE, fused_recurrence_weights, skip_stack_weights, split_f16


def _k3(sd, y, l, C):
    w = sd[f"net.residual_layers.{l}.conv_layer.conv.weight"].double()
    b = sd[f"net.residual_layers.{l}.conv_layer.conv.bias"].double()
    o = F.conv1d(y.transpose(1, 2), w, b, padding=1).transpose(1, 2)
    return torch.sigmoid(o[..., :C]) * torch.tanh(o[..., C:])


def test_y_recurrence_matches_residual_stack():
    spec = ModelSpec.preset("VCTK")
    sd = synthetic.make_acoustic_state_dict(spec, 0)
    C, H, NL = spec.res_channels, spec.hidden, spec.res_layers
    g = torch.Generator().manual_seed(0)
    B, L = 2, 37
    cond = torch.randn(B, L, H, generator=g, dtype=torch.float64)
    x0 = torch.randn(B, L, C, generator=g, dtype=torch.float64).relu()
    ds = torch.randn(B, NL, C, generator=g, dtype=torch.float64)           # diffusion_projection(step) per layer
    dsp = ds + torch.randn(B, NL, C, generator=g, dtype=torch.float64)     # + speaker_projection
    r = 1 / math.sqrt(2)

    def wb(l, name):
        p = f"net.residual_layers.{l}.{name}.conv."
        return sd[p + "weight"][:, :, 0].double(), sd[p + "bias"].double()

    # the reference's order of operations
    x, skip = x0.clone(), 0
    for l in range(NL):
        wc, bc = wb(l, "conditioner_projection")
        wo, bo = wb(l, "output_projection")
        y = x + cond @ wc.t() + bc + dsp[:, l][:, None]
        o = _k3(sd, y, l, C) @ wo.t() + bo
        x = (o[..., :C] + ds[:, l][:, None] + x) * r
        skip = skip + o[..., C:]

    # the recurrence, with the operands exactly as the kernels see them (fp16 hi/lo pairs)
    wc, bc = wb(0, "conditioner_projection")
    y = x0 + cond @ wc.t() + bc + dsp[:, 0][:, None]
    yc = r * ds[:, :-1] + dsp[:, 1:] - r * dsp[:, :-1]                       # rowops.cu: dn_fuse_steps_kernel
    gs = []
    for l in range(NL):
        gg = _k3(sd, y, l, C)
        gs.append(gg)
        if l + 1 < NL:
            wf, bf = fused_recurrence_weights(sd, l, C, H)
            assert torch.equal(wf[:, C:2 * C], r * torch.eye(C, dtype=torch.float64))   # block-diagonal y segment
            hi, lo = split_f16(wf)
            wf = (hi.double() + lo.double()) / TC_W_SCALE
            y = torch.cat([gg, y, cond], -1) @ wf.t() + bf + yc[:, l][:, None]
    # the skip sum as one GEMM over the stacked gate outputs
    wsk, bsk = skip_stack_weights(sd, NL, C)
    hi, lo = split_f16(wsk)
    wsk = (hi.double() + lo.double()) / TC_W_SCALE
    skip2 = torch.cat(gs, -1) @ wsk.reshape(NL, C, C).permute(1, 0, 2).reshape(C, NL * C).t() + bsk
    err = float((skip - skip2).abs().max())
    assert err <= 1e-5 * float(skip.abs().max()), err
