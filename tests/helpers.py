"""Shared test helpers: load golden cases and build the same problem for the
oracle (oracle/ilqr_oracle.py) and for the product API (dpilqr_b200)."""

import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def solve_case_names():
    return sorted(os.path.basename(p)[len("solve_"):-4] for p in glob.glob(os.path.join(GOLDEN, "solve_*.npz")))


def dist_case_names():
    return sorted(os.path.basename(p)[len("dist_"):-4] for p in glob.glob(os.path.join(GOLDEN, "dist_*.npz")))


def oracle_problem(case, backend="auto"):
    from oracle.ilqr_oracle import OracleProblem

    return OracleProblem(
        [str(m) for m in case["models"]], float(case["dt"]), case["xf"], list(case["Q"]), list(case["R"]), list(case["Qf"]),
        float(case["radius"]), [int(v) for v in case["n_dims"]], [int(v) for v in case["ids"]], backend=backend,
    )


def product_problem(case):
    """The same problem through the drop-in API (names as in reference dpilqr/__init__.py)."""
    import dpilqr_b200 as dp

    classes = {
        "DoubleInt4D": dp.DoubleIntDynamics4D, "DoubleInt6D": dp.DoubleIntDynamics6D, "Car3D": dp.CarDynamics3D,
        "Unicycle4D": dp.UnicycleDynamics4D, "Quadcopter6D": dp.QuadcopterDynamics6D, "Human6D": dp.HumanDynamics6D,
        "HumanLin6D": dp.HumanDynamicsLin6D, "Quadcopter12D": dp.QuadcopterDynamics12D, "Bike5D": dp.BikeDynamics5D,
    }
    dp._reset_ids()
    models = [str(m) for m in case["models"]]
    ids = [int(v) for v in case["ids"]]
    dt = float(case["dt"])
    dyn = dp.MultiDynamicalModel([classes[m](dt, id_) for m, id_ in zip(models, ids)])
    s = dyn.x_dims[0]
    x_dims = [s] * len(models)
    costs = [
        dp.ReferenceCost(xf_i, case["Q"][i].copy(), case["R"][i].copy(), case["Qf"][i].copy(), id_)
        for i, (xf_i, id_) in enumerate(zip(dp.split_agents_gen(case["xf"], x_dims), ids))
    ]
    prox = dp.ProximityCost(x_dims, float(case["radius"]), [int(v) for v in case["n_dims"]])
    return dp.ilqrProblem(dyn, dp.GameCost(costs, prox))


def rel_err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def graph_to_adj(graph, ids):
    a = len(ids)
    adj = np.zeros((a, a), dtype=np.int8)
    for i, id_ in enumerate(ids):
        for other in graph[id_]:
            adj[i, ids.index(int(other))] = 1
    return adj
