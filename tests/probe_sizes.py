#!/usr/bin/env python
"""Second-iterate backward-pass error against the oracle for several team sizes / seeds, with cond(Q_uu) and the oracle's
own movement under a reordered evaluation.  python tests/probe_sizes.py a:seed ..."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpilqr_b200 as dp  # noqa: E402
from dpilqr_b200 import scenarios  # noqa: E402
from oracle import ilqr_oracle as O  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


N = 50
for arg in sys.argv[1:]:
    a, seed = [int(v) for v in arg.split(":")]
    x0, xf, U0 = scenarios.quad12_inputs(seed, a, N)
    batch = dp.CompiledBatch([scenarios.quad12_spec(xf, a)], N)
    prob = O.OracleProblem(["Quadcopter12D"] * a, 0.1, xf, np.eye(12), np.eye(4), 1000 * np.eye(12), 0.5, [3] * a, [100 + i for i in range(a)])
    solver = O.OracleSolver(prob, N)
    Xs, Us, Js = solver.solve(x0, U0.copy(), n_lqr_iter=1)
    stage, _ = batch.linearize_quadraticize(Xs[None], Us[None])
    K, d, st = batch.backward(stage, solver.mu)
    solver.cond_log = []
    K2, d2 = solver.backward_pass(Xs, Us)
    cond = max(solver.cond_log)
    alt = O.OracleSolver(prob, N)
    alt.mu, alt.arith = solver.mu, 1
    K3, d3 = alt.backward_pass(Xs, Us)
    K = K[0].cpu().numpy()
    per_t = [rel(K[t], K2[t]) for t in range(N)]
    print(f"a={a} seed={seed} alpha {solver.trace[0]['alpha_index']} cond {cond:.1e}: K err {rel(K, K2):.1e} d err {rel(d[0].cpu().numpy(), d2):.1e} "
          f"(oracle reordered: K {rel(K3, K2):.1e} d {rel(d3, d2):.1e}) status {int(st[0])}; K err at t=49,40,25,0: "
          f"{per_t[49]:.1e} {per_t[40]:.1e} {per_t[25]:.1e} {per_t[0]:.1e}")
