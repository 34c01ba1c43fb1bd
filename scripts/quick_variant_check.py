#!/usr/bin/env python
"""Cheapest possible check of a kernel variant (SYN_LIB_PATH=...): fused engine vs the CUDA-core fp32 engine on
the same library at B=1024 (max relative error of the 62 parameters), then the device-resident step time."""
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
    got_all = model.forward_test(x).float().cpu()
    err = (got_all[:256] - ref).abs().max().item() / ref.abs().max().item()
    tail = model.forward_test(x[768:]).float().cpu()          # same faces in a different tile plan
    err2 = (got_all[768:] - tail).abs().max().item() / ref.abs().max().item()
    for _ in range(5):
        eng.forward_landmarks(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        eng.forward_landmarks(x)
    e1.record()
    torch.cuda.synchronize()
    print(f'{os.environ.get("SYN_LIB_PATH", "default")}: fused vs fp32 engine {err:.2e}, tile-plan consistency {err2:.2e}, '
          f'{e0.elapsed_time(e1) / 30:.4f} ms/step')


if __name__ == '__main__':
    main()
