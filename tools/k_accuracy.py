#!/usr/bin/env python
"""Gain accuracy of the backward kernel against the golden fixtures, iteration by iteration (teacher forced)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden, product_problem, rel_err  # noqa: E402

import dpilqr_b200 as dp  # noqa: E402

names = sys.argv[1:] or ["quad12_a10_s0", "quad12_a10_s1", "quad12_a10_s2"]
for name in names:
    case = golden(f"solve_{name}.npz")
    batch = dp.CompiledBatch([dp.spec_from_problem(product_problem(case))], int(case["N"]))
    errs = []
    for i in (0, len(case["trace_mu"]) - 1):
        stage, _ = batch.linearize_quadraticize(case["iter_X"][i][None], case["iter_U"][i][None])
        K, d, _ = batch.backward(stage, float(case["trace_mu"][i]))
        Kref = case["K_first"] if i == 0 else case["K_last_iter"]
        dref = case["d_first"] if i == 0 else case["d_last_iter"]
        errs.append((rel_err(K[0].cpu().numpy()[case["K_first_steps"]], Kref), rel_err(d[0].cpu().numpy(), dref)))
    solver = dp.ilqrSolver(product_problem(case), int(case["N"]))
    X, U, J = solver.solve(case["x0"], case["U0"].copy(), verbose=False)
    print(f"{name}: K/d err first it {errs[0][0]:.1e}/{errs[0][1]:.1e} last it {errs[1][0]:.1e}/{errs[1][1]:.1e}  "
          f"solve X {rel_err(X, case['X']):.2e} U {rel_err(U, case['U']):.2e} (ref sens {float(case['sens_X']):.1e})")
