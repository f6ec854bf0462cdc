"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: scenario sharding and the receding-horizon
all-gather.  The CUDA solver is replaced by the CPU oracle through the injection points of
dpilqr_b200.parallel.solve_distributed_sharded, so the orchestration is exercised end to end against the
golden output of the reference's solve_distributed."""

import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_solve_fn(specs, x0s, U0s, N, **kw):
    from oracle import ilqr_oracle as O

    names = {v: k for k, v in O.MODEL_IDS.items()}
    out = []
    for sp, x0, U0 in zip(specs, x0s, U0s):
        prob = O.OracleProblem([names[m] for m in sp.models], sp.dt, sp.xf, sp.Q, sp.R, sp.Qf, sp.radius, sp.n_dims, sp.ids,
                               sp.weights[0], sp.weights[1], True)
        solver = O.OracleSolver(prob, N)
        X, U, J = solver.solve(np.asarray(x0), np.asarray(U0), **kw)
        out.append({"X": X, "U": U})
    return out


def _worker(rank, world, port, case_name, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from helpers import golden
    from oracle import ilqr_oracle as O

    from dpilqr_b200 import parallel
    from dpilqr_b200.engine import ProblemSpec

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        case = golden(f"dist_{case_name}.npz")
        a = len(case["ids"])
        models = [O.MODEL_IDS[str(m)] for m in case["models"]]
        s, c = O.MODEL_DIMS[models[0]]
        spec = ProblemSpec(models, float(case["dt"]), s, c, [int(v) for v in case["n_dims"]], list(case["Q"]), list(case["R"]),
                           list(case["Qf"]), case["xf"], float(case["radius"]), (1.0, 200.0), True, [int(v) for v in case["ids"]])
        X_dec, U_dec, graph = parallel.solve_distributed_sharded(
            spec, case["X_in"], case["U0"], float(case["radius_graph"]), solve_fn=_oracle_solve_fn,
            graph_fn=O.inter_graph_threshold, n_lqr_iter=int(case["n_lqr_iter"]), tol=float(case["tol"]))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), X=X_dec, U=U_dec, owned=np.array(parallel.owned_agents(a, rank, world)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name", ["cfg2_uni4_a5_crowded", "cfg3_q6q6h6_wide"])
def test_sharded_dp_ilqr_round_matches_reference(case_name, tmp_path):
    from helpers import golden, rel_err

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case_name, str(tmp_path)), nprocs=world, join=True)
    case = golden(f"dist_{case_name}.npz")
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    # both ranks hold the same, complete result; their shards partition the agents
    assert np.array_equal(outs[0]["X"], outs[1]["X"]) and np.array_equal(outs[0]["U"], outs[1]["U"])
    assert sorted(outs[0]["owned"].tolist() + outs[1]["owned"].tolist()) == list(range(len(case["ids"])))
    tol = max(1e-9, 1000 * float(case["sens_X"]))
    assert rel_err(outs[0]["X"], case["X_dec"]) < tol and rel_err(outs[0]["U"], case["U_dec"]) < max(1e-9, 1000 * float(case["sens_U"]))


def test_shard_bounds_partition():
    from dpilqr_b200.parallel import owned_agents, shard_bounds

    for n in (0, 1, 7, 15, 4096):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert owned_agents(15, 7, 8) == [14]
