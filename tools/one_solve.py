"""One batched solve of the metric workload (for ncu launch lists / captures): python tools/one_solve.py [B] [agents]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dpilqr_b200 as dp
from dpilqr_b200 import scenarios

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
a = int(sys.argv[2]) if len(sys.argv) > 2 else 10
specs, x0, U0 = scenarios.quad12_batch(0, B, a, 50)
batch = dp.CompiledBatch(specs, 50)
x0d, U0d = torch.as_tensor(x0).cuda(), torch.as_tensor(U0).cuda()
out = batch.solve(x0d, U0d, n_lqr_iter=50, tol=1e-3)
torch.cuda.synchronize()
print("total_iters", out["total_iters"])
