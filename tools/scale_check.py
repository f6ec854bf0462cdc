#!/usr/bin/env python
"""Resident solve throughput of the Quad12D batch at several batch sizes (GPU box): tail/quantisation check."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
import dpilqr_b200 as dp
from dpilqr_b200 import scenarios
for B in (592, 1184, 2368, 4096):
    specs, x0, U0 = scenarios.quad12_batch(0, B, 10)
    batch = dp.CompiledBatch(specs, 50)
    X, J = batch.rollout(x0, U0)
    stage, _ = batch.linearize_quadraticize(X, U0)
    K, d, st = batch.backward(stage, 1.0)
    Xd, Ud = X.contiguous(), torch.as_tensor(U0).cuda()
    out = None
    for rep in range(4):
        del out
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = batch.forward_pass(Xd, Ud, K, d); e1.record(); torch.cuda.synchronize()
        t_f = e0.elapsed_time(e1)
    e0.record(); K2 = batch.backward(stage, 1.0); e1.record(); torch.cuda.synchronize()
    print(B, "linesearch %.2f ms (%.2f us/problem)  backward %.2f ms (%.2f us/problem)" % (t_f, 1e3*t_f/B, e0.elapsed_time(e1), 1e3*e0.elapsed_time(e1)/B))
    del out, K, d, K2, stage, X, batch
    torch.cuda.empty_cache()
