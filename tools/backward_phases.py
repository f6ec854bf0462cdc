#!/usr/bin/env python
"""Per-phase cycle counts of the backward kernel (CTA 0), via dpilqr_debug_backward_timing."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpilqr_b200 as dp  # noqa: E402
from dpilqr_b200 import _native, scenarios  # noqa: E402

a = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
specs, x0, U0 = scenarios.quad12_batch(0, B, a)
batch = dp.CompiledBatch(specs, 50)
X, J = batch.rollout(x0, U0)
stage, _ = batch.linearize_quadraticize(X, U0)
# product (uninstrumented) kernel first: best of a few launches, clocks warm
best = 1e9
for rep in range(12):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    batch.backward(stage, 1.0)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"product kernel: best of 12 launches {best:.3f} ms for {B} problems = {best * 1e-3 * 1.965e9 / 50:.0f} cycles/step at 1965 MHz")
buf = torch.zeros(96, dtype=torch.int64, device="cuda")
_native.lib().dpilqr_debug_backward_timing(ctypes.c_void_p(buf.data_ptr()))
tbest = 1e9
for rep in range(6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K, d, st = batch.backward(stage, 1.0)
    e1.record()
    torch.cuda.synchronize()
    tbest = min(tbest, e0.elapsed_time(e1))
print(f"instrumented kernel: best of 6 launches {tbest:.3f} ms for {B} problems (a={a})")
c = buf.cpu().numpy()
names = ["load", "phaseA", "LU", "join", "D-out", "pack", "Epre", "E", "F", "regul.", "D-trsm", "prefetch"]
tot = c[:12].sum()
for k, nm in enumerate(names):
    print(f"  thread0 {nm:7s} {c[k] / 50:10.0f} cycles/step {100 * c[k] / max(tot, 1):5.1f}%")
print(f"  group2 phaseB {c[14] / 50:10.0f} cycles/step; total {tot / 50:.0f} cycles/step")
print(f"  group 2 detail: restore+barrier {c[15] / 50:.0f}  Q_xx blocks {c[16] / 50:.0f}  Q_ux tiles {c[14] / 50:.0f} cycles/step")
print(f"  LU detail: first panel {c[24] / 50:.0f}  U12+tiles+waits {c[25] / 50:.0f}  panels 1..4 {c[26] / 50:.0f} cycles/step")
print(f"  LU experiment (debug modes 32/256/512): before the recursion {c[27]}, top of the first step {c[28]}, after phase A of the second step {c[29]} cycles")
if c[32:48].any():
    nl = 6 * 50  # instrumented launches x steps
    print("  phase E per warp, cycles/step to the end of its own work:", " ".join(f"{v / nl:.0f}" for v in c[32:48]))
    print("  phase E per warp, cycles/step to the end of its pq:      ", " ".join(f"{v / nl:.0f}" for v in c[48:64]))
if c[64:67].any():
    nl = 6 * 50
    print(f"  LU between the panels (warp 0, cycles/step): barrier {c[64] / nl:.0f}  U12 {c[65] / nl:.0f}  look-ahead tile update {c[66] / nl:.0f}")
_native.lib().dpilqr_debug_backward_timing(None)
