"""Seeded, *calibrated* synthetic checkpoint in the reference's key schema.  TEST INFRASTRUCTURE.

A randomly initialised MobileNetV2 in eval mode with arbitrary BatchNorm statistics forgets its
input after a few depthwise layers (kaiming fan_out weights shrink the signal by ~C/2 per
depthwise conv), so every crop would give the same 62 parameters and parity tests would only
exercise the bias path.  Here each BatchNorm's running statistics are set to the float64 batch
statistics of its own input on a small calibration batch (what training-mode BN would have
converged to), then perturbed, so that the signal survives all 52 convolutions like in a
trained checkpoint.  All randomness comes from seeded CPU generators; the statistics are computed
in float64 so that the fp32 result is reproducible across hosts.

Follows the module structure of reference backbone_nets/mobilenetv2_backbone.py:104-158.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from synergynet_b200 import synthetic
from synergynet_b200.backbone import conv_plan

_CACHE: Dict[int, Dict[str, torch.Tensor]] = {}


@torch.no_grad()
def build_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Full 445-key state dict (CPU fp32) of ``synergy3DMM.SynergyNet`` with synthetic 3DMM
    buffers (seed 0 model), seeded weights and calibrated BatchNorm statistics."""
    if seed in _CACHE:
        return _CACHE[seed]
    from synergynet_b200.params import ParamsPack, set_param_pack
    from synergynet_b200 import synergy3DMM
    pack = ParamsPack(arrays=synthetic.make_3dmm(seed=0))
    set_param_pack(pack)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = synergy3DMM.SynergyNet()
    synthetic.seeded_init_(model, seed)
    synthetic.randomize_batchnorm_(model, seed)          # gamma/beta (and the PointNet BN stats)
    sd = model.state_dict()
    g = torch.Generator().manual_seed(5000 + seed)
    x = synthetic.normalize_crops(synthetic.make_structured_crops_u8(8, seed=100 + seed)).double()
    pre = 'I2P.backbone.'
    block_in = None
    for spec in conv_plan():
        if spec.kind in ('expand', 'dw') and (spec.kind == 'expand' or spec.cin == 32 and spec.block == 1):
            block_in = x
        w = sd[pre + spec.conv_key + '.weight'].double()
        y = F.conv2d(x, w, None, spec.stride, (spec.ksize - 1) // 2, 1, spec.groups)
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        n = spec.cout
        mean = mean + 0.1 * var.sqrt() * torch.randn(n, generator=g, dtype=torch.float64)
        var = var * (0.8 + 0.4 * torch.rand(n, generator=g, dtype=torch.float64)) + 1e-6
        sd[pre + spec.bn_key + '.running_mean'].copy_(mean.float())
        sd[pre + spec.bn_key + '.running_var'].copy_(var.float())
        gamma = sd[pre + spec.bn_key + '.weight'].double()
        beta = sd[pre + spec.bn_key + '.bias'].double()
        rm = sd[pre + spec.bn_key + '.running_mean'].double()
        rv = sd[pre + spec.bn_key + '.running_var'].double()
        y = (y - rm.view(1, -1, 1, 1)) / torch.sqrt(rv.view(1, -1, 1, 1) + 1e-5) * gamma.view(1, -1, 1, 1) \
            + beta.view(1, -1, 1, 1)
        if spec.relu6:
            y = y.clamp(0, 6)
        if spec.residual:
            y = y + block_in
        x = y
    _calibrate_pointnet(sd, x.float().mean(dim=(2, 3)), seed)
    _CACHE[seed] = sd
    return sd


@torch.no_grad()
def _calibrate_pointnet(sd: Dict[str, torch.Tensor], pooled: torch.Tensor, seed: int) -> None:
    """Same treatment for the BatchNorm1d layers of forwardDirection / reverseDirection (reference
    backbone_nets/pointnet_backbone.py:7-106): with arbitrary statistics every ReLU of the heads is dead (the
    landmark coordinates are ~100 px, the random convs scale them further) and the refinement would be identically
    zero.  Each BN gets the float64 batch statistics of its own input on a calibration batch of landmarks, perturbed,
    so that about half of the units are active at every layer."""
    import numpy as np
    from oracle import reference_port as rp
    g = torch.Generator().manual_seed(7000 + seed)
    basis = rp.gather_sparse_basis(synthetic.make_3dmm(0))
    heads = [F.linear(pooled, sd[f'I2P.backbone.{h}.1.weight'], sd[f'I2P.backbone.{h}.1.bias'])
             for h in ('classifier_ori', 'classifier_shape', 'classifier_exp')]
    attr = torch.cat(heads, 1)
    lmk = torch.from_numpy(rp.reconstruct_vertex_62(attr.numpy(), basis)).double()       # (8,3,68)

    def layer(pre, x, conv, bn):
        y = F.conv1d(x, sd[f'{pre}{conv}.weight'].double(), sd[f'{pre}{conv}.bias'].double())
        n = y.shape[1]
        mean, var = y.mean(dim=(0, 2)), y.var(dim=(0, 2), unbiased=False)
        mean = mean + 0.1 * var.sqrt() * torch.randn(n, generator=g, dtype=torch.float64)
        var = var * (0.8 + 0.4 * torch.rand(n, generator=g, dtype=torch.float64)) + 1e-6
        sd[f'{pre}{bn}.running_mean'].copy_(mean.float())
        sd[f'{pre}{bn}.running_var'].copy_(var.float())
        rm, rv = sd[f'{pre}{bn}.running_mean'].double(), sd[f'{pre}{bn}.running_var'].double()
        gamma, beta = sd[f'{pre}{bn}.weight'].double(), sd[f'{pre}{bn}.bias'].double()
        y = (y - rm.view(1, -1, 1)) / torch.sqrt(rv.view(1, -1, 1) + 1e-5) * gamma.view(1, -1, 1) + beta.view(1, -1, 1)
        return y.clamp(min=0)

    for pre in ('forwardDirection.', 'reverseDirection.'):
        out = lmk
        for i in range(1, 6):
            out = layer(pre, out, f'conv{i}', f'bn{i}')
            if i == 2:
                pf = out
        glob = out.max(dim=2, keepdim=True).values
        if pre == 'forwardDirection.':
            rep = lambda t: t.repeat(1, 1, 68)
            pool = torch.randn(lmk.shape[0], 1280, 1, generator=g, dtype=torch.float64).abs() * 0.5
            cat = torch.cat([pf, rep(glob), rep(pool), rep(attr[:, 12:52].double().unsqueeze(2)),
                             rep(attr[:, 52:62].double().unsqueeze(2))], 1)
            out = layer(pre, cat, 'conv6', 'bn6')
            for i in (7, 8, 9):
                out = layer(pre, out, f'conv{i}', f'bn{i}')
        else:
            for i in (1, 2, 3):
                layer(pre, glob, f'conv6_{i}', f'bn6_{i}')


@torch.no_grad()
def build_resnet50_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded state dict of ``resnet_backbone.resnet50()`` (reference backbone_nets/resnet_backbone.py:148-249 key
    schema, keys without prefix): kaiming convs like the reference's own init, randomised BatchNorm affine parameters
    and running statistics so that BN folding is exercised.  ReLU networks with residual connections keep their signal
    without calibration; magnitudes grow to a few hundred, which is exactly what the dynamic row scaling of the GEMM
    layers is for."""
    key = ('resnet50', seed)
    if key in _CACHE:
        return _CACHE[key]
    from synergynet_b200 import backbone
    m = backbone.resnet50()
    synthetic.seeded_init_(m, 300 + seed)
    synthetic.randomize_batchnorm_(m, 300 + seed)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    _CACHE[key] = sd
    return sd
