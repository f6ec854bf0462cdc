#!/usr/bin/env python
"""Product backward kernel, best-of-n launch time per team size (full waves of 148 problems).
usage: [DPILQR_B200_LIB=dpilqr_b200/lib/variants/X.so] backward_time.py 6 8 10 12 14 15"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpilqr_b200 as dp  # noqa: E402
from dpilqr_b200 import scenarios  # noqa: E402

B = 148 * 2
for a in [int(v) for v in sys.argv[1:]] or [10]:
    specs, x0, U0 = scenarios.quad12_batch(0, B, a)
    batch = dp.CompiledBatch(specs, 50)
    X, J = batch.rollout(x0, U0)
    stage, _ = batch.linearize_quadraticize(X, U0)
    best = 1e9
    for rep in range(8):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        batch.backward(stage, 1.0)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{os.environ.get('DPILQR_B200_LIB', 'product')[-24:]:24s} a={a:2d}: {best:8.3f} ms for {B} problems = {best * 1e-3 * 1.965e9 / 50 / 2:.0f} cycles/step", flush=True)
    del batch, stage, X
    torch.cuda.empty_cache()
