#!/usr/bin/env python
"""Product rollout / line-search kernel: best-of-n time of one forward_pass call per candidate count.
usage: [DPILQR_B200_LIB=...] linesearch_time.py [agents] [problems] [n_alpha ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpilqr_b200 as dp  # noqa: E402
from dpilqr_b200 import scenarios  # noqa: E402

a = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
specs, x0, U0 = scenarios.quad12_batch(0, B, a)
batch = dp.CompiledBatch(specs, 50)
X, J = batch.rollout(x0, U0)
stage, _ = batch.linearize_quadraticize(X, U0)
K, d, st = batch.backward(stage, 1.0)
for NA in [int(v) for v in sys.argv[3:]] or [1, 2, 10]:
    alphas = [1.1 ** (-k * k) for k in range(NA)]
    best = 1e9
    for rep in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Xc, Uc, Jc = batch.forward_pass(X, U0, K, d, alphas)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{os.environ.get('DPILQR_B200_LIB', 'product')[-24:]:24s} a={a} B={B} candidates={NA:2d}: {best:8.3f} ms  J[0]={Jc.flatten()[:NA].sum().item():.12e}", flush=True)
