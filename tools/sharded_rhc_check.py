#!/usr/bin/env python
"""Multi-GPU check (run under torchrun with >= 2 ranks, NCCL): a decentralized receding-horizon run whose agents'
sub-problems are sharded over the ranks, with one all-gather per round, must reproduce the single-GPU result bit for
bit on every rank (independent units + a pure data exchange: sharding must not change numerics)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dpilqr_b200 as dp  # noqa: E402
from dpilqr_b200 import scenarios  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
a, N = 15, 50
x0, xf, U0 = scenarios.quad12_inputs(0, a, N)
dp._reset_ids()
ids = [100 + i for i in range(a)]
dyn = dp.MultiDynamicalModel([dp.QuadcopterDynamics12D(0.1, id_) for id_ in ids])
costs = [dp.ReferenceCost(xf[12 * i:12 * i + 12], np.eye(12), np.eye(4), 1000 * np.eye(12), id_) for i, id_ in enumerate(ids)]
prob = dp.ilqrProblem(dyn, dp.GameCost(costs, dp.ProximityCost([12] * a, 0.5, [3] * a)))
kw = dict(n_d=3, step_size=5, dist_converge=0.2, t_diverge=1.0, U0=U0, n_lqr_iter=8)
Xs, Us, Js = dp.solve_rhc(prob, x0, N, 0.5, [], centralized=False, sharded=True, **kw)
Xr, Ur, Jr = dp.solve_rhc(prob, x0, N, 0.5, [], centralized=False, sharded=False, **kw)
same = np.array_equal(Xs, Xr) and np.array_equal(Us, Ur) and Js == Jr
flag = torch.tensor([1 if same else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded RHC over {dist.get_world_size()} ranks: {Xs.shape[0]} steps, J={Js:.6f}, identical to single-GPU on all ranks: {bool(flag.item())}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
