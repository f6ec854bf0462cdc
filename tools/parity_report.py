#!/usr/bin/env python
"""Per-iteration parity report of the CUDA solve against the golden fixtures (GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden, product_problem, rel_err, solve_case_names  # noqa: E402

import dpilqr_b200 as dp  # noqa: E402

for name in solve_case_names():
    case = golden(f"solve_{name}.npz")
    solver = dp.ilqrSolver(product_problem(case), int(case["N"]))
    X, U, J = solver.solve(case["x0"], case["U0"].copy(), n_lqr_iter=int(case["n_lqr_iter"]), tol=float(case["tol"]), verbose=False)
    tr = solver.last_trace
    n_ref = len(case["trace_mu"])
    print(f"== {name}: iters {tr['iters']} (ref {n_ref}) alpha {tr['alpha_index'].tolist()} ref {case['trace_alpha'].tolist()}")
    print(f"   X {rel_err(X, case['X']):.2e} U {rel_err(U, case['U']):.2e} J {abs(J - case['J']) / abs(case['J']):.2e}")
    Js = float(case["J0"])
    for i in range(min(tr["iters"], n_ref)):
        k = int(case["trace_alpha"][i])
        got, ref = tr["J_tried"][i], case["trace_J"][i]
        sl = slice(0, k + 1) if k >= 0 else slice(0, 10)
        err = np.abs(got[sl] - ref[sl]) / np.abs(ref[sl])
        acc_err = err[k] if k >= 0 else float("nan")
        rej = err[:k] if k >= 0 else err
        dJ = abs((Js - ref[k]) / Js) if k >= 0 else float("nan")
        print(f"   it {i:2d} acc {k:2d} accepted-J err {acc_err:.1e} max rejected err {(rej.max() if rej.size else 0):.1e} |dJ/J| {dJ:.3e}")
        if k >= 0:
            Js = ref[k]
