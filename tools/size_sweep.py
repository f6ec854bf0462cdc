#!/usr/bin/env python
"""Per-kernel time per problem as a function of the team size (Quadcopter12D): the cost table of the DP-iLQR
sub-problem bins.  usage: size_sweep.py [B] [sizes...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpilqr_b200 as dp  # noqa: E402
from dpilqr_b200 import _native, scenarios  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sizes = [int(v) for v in sys.argv[2:]] or [1, 2, 3, 4, 5, 6, 8, 10]


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


print(f"{'a':>3s} {'linquad':>9s} {'backward':>9s} {'rollout1':>9s} {'search10':>9s}   us per problem (B={B});  backward MFLOP, TFLOP/s")
for a in sizes:
    specs, x0, U0 = scenarios.quad12_batch(0, B, a)
    batch = dp.CompiledBatch(specs, 50)
    x0, U0 = torch.as_tensor(x0).cuda(), torch.as_tensor(U0).cuda()
    X, J = batch.rollout(x0, U0)
    t_lq, (stage, _) = timed(lambda: batch.linearize_quadraticize(X, U0))
    t_bw, (K, d, st) = timed(lambda: batch.backward(stage, 1.0))
    t_r1, _ = timed(lambda: batch.forward_pass(X, U0, K, d, [1.0]))
    t_ls, _ = timed(lambda: batch.forward_pass(X, U0, K, d))
    n, m, s = 12 * a, 4 * a, 12
    fl = 50 * (4 * n * n * s + 4 * m * n * s + 2 * m * m * s + 2 * m ** 3 / 3 + 2 * m * m * (n + 1) + 2 * n * m * m + 4 * n * n * m + 8 * n * m + 2 * n * 16)
    line = f"{a:3d} {1e3 * t_lq / B:9.2f} {1e3 * t_bw / B:9.2f} {1e3 * t_r1 / B:9.2f} {1e3 * t_ls / B:9.2f}   {fl / 1e6:8.2f} {fl * B / (t_bw * 1e-3) / 1e12:6.2f}"
    # inside the solver loop (odd teams: records padded by a phantom agent, tensor-path kernel of the next even size)
    _native.get_profile(reset=True)
    out = batch.solve(x0, U0, n_lqr_iter=50, tol=1e-3, profile=True)
    prof = _native.get_profile(reset=True)
    bms, _, bunits = prof["backward"]
    lms, _, lunits = prof["linesearch"]
    qms, _, qunits = prof["linquad"]
    line += (f"   | in the solver: {out['total_iters']} problem-iterations, backward {1e3 * bms / max(bunits, 1):.2f} line search "
             f"{1e3 * lms / max(lunits, 1):.2f} linquad {1e3 * qms / max(qunits, 1):.2f} us each")
    print(line)
    del batch, stage, K, d, X
    torch.cuda.empty_cache()
