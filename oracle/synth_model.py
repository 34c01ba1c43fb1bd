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
    _CACHE[seed] = sd
    return sd
