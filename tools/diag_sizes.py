import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import dpilqr_b200 as dp
from dpilqr_b200 import scenarios
from oracle import ilqr_oracle as O
from helpers import rel_err
for a, seed in [(8,3),(10,3),(11,3),(12,3),(12,5),(13,3),(14,3),(15,3)]:
    N=50
    x0, xf, U0 = scenarios.quad12_inputs(seed, a, N)
    batch = dp.CompiledBatch([scenarios.quad12_spec(xf, a)], N)
    X, J = batch.rollout(x0[None], U0[None])
    stage, _ = batch.linearize_quadraticize(X, U0[None])
    prob = O.OracleProblem(["Quadcopter12D"]*a, 0.1, xf, np.eye(12), np.eye(4), 1000*np.eye(12), 0.5, [3]*a, [100+i for i in range(a)])
    solver = O.OracleSolver(prob, N)
    Xo, Jo = solver.rollout(x0, U0)
    for mu in (1.0, 0.0):
        K, d, st = batch.backward(stage, mu)
        solver.mu = mu
        Ko, do = solver.backward_pass(Xo, U0)
        print(a, seed, "mu", mu, "K err %.1e d err %.1e status %d" % (rel_err(K[0].cpu().numpy(), Ko), rel_err(d[0].cpu().numpy(), do), int(st[0])))
