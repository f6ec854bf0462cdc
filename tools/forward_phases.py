#!/usr/bin/env python
"""Per-phase cycle counts of the rollout / line-search kernel (CTA 0, thread 0)."""
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
specs, x0, U0 = scenarios.quad12_batch(0, B, a)
batch = dp.CompiledBatch(specs, 50)
X, J = batch.rollout(x0, U0)
stage, _ = batch.linearize_quadraticize(X, U0)
K, d, st = batch.backward(stage, 1.0)
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
_native.lib().dpilqr_debug_backward_timing(ctypes.c_void_p(buf.data_ptr()))
for rep in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Xc, Uc, Jc = batch.forward_pass(X, U0, K, d)
    e1.record()
    torch.cuda.synchronize()
print(f"line-search launch: {e0.elapsed_time(e1):.3f} ms for {B} problems (a={a}, 10 candidates)")
c = buf.cpu().numpy()[24:30]
names = ["load/store X", "gain", "store U", "agents (cost+RK4)", "pairs", "sum"]
for k, nm in enumerate(names):
    print(f"  {nm:18s} {c[k] / 51:10.0f} cycles/step {100 * c[k] / max(c.sum(), 1):5.1f}%")
print(f"  total {c.sum() / 51:.0f} cycles/step")
_native.lib().dpilqr_debug_backward_timing(None)
