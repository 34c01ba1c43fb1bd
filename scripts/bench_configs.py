#!/usr/bin/env python
"""Secondary measurements (BASELINE.json configs[2], per-stage timings).  Not the headline bench.

  python scripts/bench_configs.py dense   # config 3: params (B=1024) -> dense (B,3,53215) vertices
Prints one JSON line per measurement; CUDA events, >= 5 warm-up iterations.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def time_cuda(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'dense'
    peaks = bench.load_peaks()
    model = bench.build_model('cuda:0')
    model.set_engine(int(os.environ.get('SYN_ENGINE', '2')))
    eng = model._engine(torch.device('cuda', 0))
    B = 1024
    from synergynet_b200 import synthetic
    if what == 'dense':
        params = model.forward_test(synthetic.make_inputs(B, 0).cuda())
        out = [None]

        def fn():
            out[0] = eng.reconstruct(params, dense=True)
        ms = time_cuda(fn)
        nbytes = B * 3 * 53215 * 4
        print(json.dumps({'config': 'configs[2]: batch=1024 params -> dense (B,3,53215) vertices', 'ms': ms,
                          'faces_per_s': B / ms * 1e3, 'roofline': {'bound': 'hbm', 'achieved': nbytes / ms / 1e6,
                          'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': nbytes / ms / 1e6 / peaks['hbm'],
                          'what': '638,580 B written per face x 1024 / CUDA-event time; ' + peaks['source']}}))
    elif what == 'h2d':
        for mb in (44, 177):
            n = mb * 1000 * 1000 // 4
            src = torch.empty(n, dtype=torch.float32).pin_memory()
            dst = torch.empty(n, dtype=torch.float32, device='cuda')
            ms = time_cuda(lambda: dst.copy_(src, non_blocking=True), iters=10, warmup=3)
            print(json.dumps({'h2d_pinned_MB': mb, 'ms': ms, 'GBps': n * 4 / ms / 1e6}))
    elif what == 'sparse':
        params = model.forward_test(synthetic.make_inputs(B, 0).cuda())
        ms = time_cuda(lambda: eng.reconstruct(params, dense=False))
        print(json.dumps({'config': 'params -> 68 landmarks, B=1024', 'ms': ms}))


if __name__ == '__main__':
    main()
