#!/usr/bin/env python
"""Error budget of the CUDA path against the CPU oracle on one metric scenario, component by component and iteration by
iteration (teacher forced from the oracle's own iterates), with the oracle's OWN movement under an equally valid
evaluation order (OracleSolver.arith) beside it.  Diagnostic for the GPU box, not collected by pytest:
    python tests/probe_precision.py [seed ...]"""
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


a, T = int(os.environ.get("AGENTS", "10")), 50
for seed in [int(v) for v in sys.argv[1:]] or [41]:
    x0, xf, U0 = scenarios.quad12_inputs(seed, a, T)
    prob = O.OracleProblem(["Quadcopter12D"] * a, 0.1, xf, np.eye(12), np.eye(4), 1000 * np.eye(12), 0.5, [3] * a, [100 + i for i in range(a)])
    orc = O.OracleSolver(prob, T)
    Xo, Uo, Jo = orc.solve(x0, U0.copy(), keep_gains=True)
    batch = dp.CompiledBatch([scenarios.quad12_spec(xf, a)], T)
    out = batch.solve(x0[None], U0[None], trace=True)
    print(f"== seed {seed}: oracle {len(orc.trace)} its, gpu {int(out['iters'][0])}; whole solve X err {rel(out['X'][0].cpu().numpy(), Xo):.2e}")
    alphas = O.alpha_table()
    for i, rec in enumerate(orc.trace):
        X, U, mu = rec["X"], rec["U"], rec["mu"]
        stage, _ = batch.linearize_quadraticize(X[None], U[None])
        A, Bm, Lx, Lu, Lxx, Luu = [v[0].cpu().numpy() for v in batch.stage_to_dense(stage)]
        eA = eL = 0.0
        for t in (0, T // 2, T - 1):
            Ao, Bo = prob.linearize(X[t], U[t])
            lx, lu, lxx, luu, _ = prob.quadraticize(X[t], U[t])
            eA = max(eA, rel(A[t], Ao), rel(Bm[t], Bo))
            eL = max(eL, rel(Lx[t], lx), rel(Lxx[t], lxx), rel(Lu[t], lu))
        K, d, _ = batch.backward(stage, mu)
        K, d = K[0].cpu().numpy(), d[0].cpu().numpy()
        alt = O.OracleSolver(prob, T)
        alt.mu, alt.arith = mu, 1
        Ka, da = alt.backward_pass(X, U)
        orc2 = O.OracleSolver(prob, T)
        orc2.mu = mu
        orc2.cond_log = []
        orc2.backward_pass(X, U)
        k = rec["alpha_index"]
        line = (f"  it {i:2d} mu {mu:.1e} cond(Quu) max {max(orc2.cond_log):.1e} | lin {eA:.1e} quad {eL:.1e} | K gpu {rel(K, rec['K']):.1e} "
                f"oracle-reordered {rel(Ka, rec['K']):.1e} | d gpu {rel(d, rec['d']):.1e} reordered {rel(da, rec['d']):.1e}")
        if k >= 0:
            Xc, Uc, Jc = batch.forward_pass(X[None], U[None], rec["K"][None], rec["d"][None], alphas=[float(alphas[k])])
            Xn, Un, Jn = orc.forward_pass(X, U, rec["K"], rec["d"], alphas[k])
            line += f" | fwd (oracle gains) X {rel(Xc[0, 0].cpu().numpy(), Xn):.1e} J {abs(float(Jc[0, 0]) - Jn) / abs(Jn):.1e}"
            Xc2, _, _ = batch.forward_pass(X[None], U[None], K[None], d[None], alphas=[float(alphas[k])])
            Xr, _, _ = orc.forward_pass(X, U, Ka, da, alphas[k])
            line += f" | fwd X with own gains: gpu {rel(Xc2[0, 0].cpu().numpy(), Xn):.1e} reordered oracle {rel(Xr, Xn):.1e}"
        print(line)
