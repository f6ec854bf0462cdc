#!/usr/bin/env python
"""Per-time-step error of one backward pass (K[t], d[t]) against the CPU oracle, from the oracle's second iterate of a
metric scenario.  Diagnostic for the GPU box:  [AGENTS=3] [DPILQR_BACKWARD_FORCE_BIG=1] python tests/probe_backward_steps.py [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpilqr_b200 as dp  # noqa: E402
from dpilqr_b200 import scenarios  # noqa: E402
from oracle import ilqr_oracle as O  # noqa: E402

a, T = int(os.environ.get("AGENTS", "10")), 50
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 41
it = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x0, xf, U0 = scenarios.quad12_inputs(seed, a, T)
prob = O.OracleProblem(["Quadcopter12D"] * a, 0.1, xf, np.eye(12), np.eye(4), 1000 * np.eye(12), 0.5, [3] * a, [100 + i for i in range(a)])
orc = O.OracleSolver(prob, T)
orc.solve(x0, U0.copy(), keep_gains=True)
rec = orc.trace[min(it, len(orc.trace) - 1)]
batch = dp.CompiledBatch([scenarios.quad12_spec(xf, a)], T)
stage, _ = batch.linearize_quadraticize(rec["X"][None], rec["U"][None])
K, d, _ = batch.backward(stage, rec["mu"])
K, d = K[0].cpu().numpy(), d[0].cpu().numpy()
for t in range(T - 1, -1, -1):
    eK = np.max(np.abs(K[t] - rec["K"][t])) / np.max(np.abs(rec["K"][t]))
    ed = np.max(np.abs(d[t] - rec["d"][t])) / np.max(np.abs(rec["d"][t]))
    if t > T - 8 or t % 8 == 0:
        print(f"t {t:2d} K {eK:.1e} d {ed:.1e}   |d| {np.max(np.abs(rec['d'][t])):.2e} |K| {np.max(np.abs(rec['K'][t])):.2e}")
