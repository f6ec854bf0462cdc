#!/usr/bin/env python
"""Per-phase cycle counts of the rollout / line-search kernel (CTA 0, thread 0).
usage: forward_phases.py [agents] [problems] [n_alpha]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpilqr_b200 as dp  # noqa: E402
from dpilqr_b200 import _native, scenarios  # noqa: E402

a = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 592
NA = int(sys.argv[3]) if len(sys.argv) > 3 else 10
specs, x0, U0 = scenarios.quad12_batch(0, B, a)
batch = dp.CompiledBatch(specs, 50)
X, J = batch.rollout(x0, U0)
stage, _ = batch.linearize_quadraticize(X, U0)
K, d, st = batch.backward(stage, 1.0)
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
_native.lib().dpilqr_debug_backward_timing(ctypes.c_void_p(buf.data_ptr()))
alphas = [1.1 ** (-k * k) for k in range(NA)]
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Xc, Uc, Jc = batch.forward_pass(X, U0, K, d, alphas)
    e1.record()
    torch.cuda.synchronize()
print(f"line-search launch: {e0.elapsed_time(e1):.3f} ms for {B} problems (a={a}, {NA} candidates)")
c = buf.cpu().numpy()[24:32]
names = ["P1 sum/dx/X out/pairs", "barrier 1", "P2 gains", "barrier 2", "fetch/prefetch/U out", "P3 cost", "P3 integrate", "barrier 3"]
for k, nm in enumerate(names):
    print(f"  {nm:24s} {c[k] / 51:10.0f} cycles/step {100 * c[k] / max(c.sum(), 1):5.1f}%")
print(f"  total {c.sum() / 51:.0f} cycles/step")
_native.lib().dpilqr_debug_backward_timing(None)
