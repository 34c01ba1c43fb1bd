#!/usr/bin/env python
"""quick_variant_check.py plus the per-kernel CUDA-event times (syn_set_timing) of one step: one line per library /
environment, for A/B runs of kernel variants (SYN_LIB_PATH, SYN_FUSED_WARPS, SYN_FUSED_WARPS_MAP)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from synergynet_b200 import synthetic  # noqa: E402


def main():
    model = bench.build_model('cuda:0')
    eng = model._engine(torch.device('cuda', 0))
    x = synthetic.make_inputs(1024, 3).cuda()
    eng.set_engine(0)
    ref = model.forward_test(x[:256]).float().cpu()
    eng.set_engine(2)
    got = model.forward_test(x).float().cpu()
    err = (got[:256] - ref).abs().max().item() / ref.abs().max().item()
    for _ in range(5):
        eng.forward_landmarks(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        eng.forward_landmarks(x)
    e1.record()
    torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / 40
    eng.set_timing(True)
    acc = {}
    for _ in range(5):
        eng.forward_landmarks(x)
        for name, ms in eng.timings():
            acc[name] = acc.get(name, 0.0) + ms / 5
    eng.set_timing(False)
    flag = eng.poll_error() if hasattr(eng, 'poll_error') else 0
    tag = os.path.basename(os.environ.get('SYN_LIB_PATH', 'default')) + ' W=' + os.environ.get('SYN_FUSED_WARPS', '16') + ' ' + os.environ.get('SYN_FUSED_WARPS_MAP', '')
    top = ' '.join(f'{k.replace("fused_", "")}={v:.3f}' for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:8])
    print(f'{tag}: err {err:.2e} flag {flag} step {step:.4f} ms | {top}')


if __name__ == '__main__':
    main()
