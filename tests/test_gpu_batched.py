"""Device-resident batched drivers (dpilqr_b200/batched.py) against the per-scenario drop-in paths and the
reference goldens: one DP-iLQR round for many scenarios, the batched receding-horizon loop, trajectory metrics."""

import io

import numpy as np
import pytest

from helpers import dist_case_names, golden, product_problem, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _tol(sens):
    return max(TOL, 1000.0 * float(sens))


@pytest.mark.parametrize("name", dist_case_names())
def test_distributed_round_vs_reference_golden(name):
    import dpilqr_b200 as dp

    case = golden(f"dist_{name}.npz")
    if len(case["ignore_ids"]):
        pytest.skip("ignore_ids case")
    prob = product_problem(case)
    N = int(case["N"])
    batch = dp.CompiledBatch([dp.spec_from_problem(prob)] * 3, N)
    X_in = np.stack([case["X_in"]] * 3)
    out = dp.solve_distributed_round(batch, X_in, np.stack([case["U0"]] * 3), float(case["radius_graph"]),
                                     n_lqr_iter=int(case["n_lqr_iter"]), tol=float(case["tol"]))
    ids = [int(v) for v in case["ids"]]
    adj = out["adjacency"].cpu().numpy()
    for k in range(3):
        got = np.array([[(int(adj[k, i]) >> j) & 1 for j in range(len(ids))] for i in range(len(ids))], dtype=np.int8)
        assert np.array_equal(got, case["adjacency"])
        assert out["sub_iters"][k].cpu().numpy().tolist() == case["sub_iters"].tolist()
    X, U, J = out["X_dec"].cpu().numpy(), out["U_dec"].cpu().numpy(), out["J_full"].cpu().numpy()
    assert rel_err(X[0], case["X_dec"]) < _tol(case["sens_X"]) and rel_err(U[0], case["U_dec"]) < _tol(case["sens_U"])
    assert abs(J[0] - case["J_full"]) <= _tol(case["sens_J"]) * abs(case["J_full"])
    assert np.array_equal(X[1], X[0]) and np.array_equal(X[2], X[0]) and np.array_equal(U[2], U[0])
    # bitwise the per-scenario drop-in
    X2, U2, J2, _ = dp.solve_distributed(prob, case["X_in"], case["U0"], float(case["radius_graph"]), [], None, False,
                                         n_lqr_iter=int(case["n_lqr_iter"]), tol=float(case["tol"]))
    assert np.array_equal(X2, X[0]) and np.array_equal(U2, U[0]) and J2 == J[0]
    assert out["total_iters"] == 3 * int(case["sub_iters"].sum())


@pytest.mark.parametrize("name", ["cfg1_dint4_a3_central", "cfg1_dint4_a3_dec", "cfg4_quad12_a15_dec"])
def test_rhc_batch_vs_reference_golden(name):
    """The whole receding-horizon run, several copies of the scenario advanced together, against the unmodified
    reference's run (tests/golden/rhc_*.npz) and bitwise against the per-scenario drop-in solve_rhc."""
    import dpilqr_b200 as dp

    g = golden(f"rhc_{name}.npz")
    prob = product_problem(g)
    N, reps = int(g["N"]), 2
    batch = dp.CompiledBatch([dp.spec_from_problem(prob)] * reps, N)
    kw = dict(centralized=bool(g["centralized"]), n_d=int(g["n_d"]), step_size=int(g["step_size"]),
              dist_converge=float(g["dist_converge"]), t_diverge=float(g["t_diverge"]),
              n_lqr_iter=int(g["n_lqr_iter"]), tol=float(g["tol"]))
    log = []
    out = dp.solve_rhc_batch(batch, np.stack([g["x0"]] * reps), radius=float(g["radius_graph"]), U0=np.stack([g["U_init"]] * reps),
                             log=log, model_name=str(g["models"][0]), ids=[int(v) for v in g["ids"]], **kw)
    tol = _tol(g["sens_X"])
    for k in range(reps):
        X, U = out["X_full"][k].cpu().numpy(), out["U_full"][k].cpu().numpy()
        assert X.shape == g["X_full"].shape and U.shape == g["U_full"].shape
        assert rel_err(X, g["X_full"]) < tol and rel_err(U, g["U_full"]) < tol
        assert abs(float(out["J_full"][k]) - float(g["J_full"])) <= tol * abs(float(g["J_full"]))
        assert int(out["rounds"][k]) == len(g["round_J"])
    assert len(log) == reps * len(g["round_J"]) and log[0].count(",") >= 13
    args = () if kw["centralized"] else (float(g["radius_graph"]), [])
    X1, U1, J1 = dp.solve_rhc(prob, g["x0"], N, *args, U0=g["U_init"], **kw)
    assert np.array_equal(X1, out["X_full"][0].cpu().numpy()) and np.array_equal(U1, out["U_full"][0].cpu().numpy())
    assert J1 == float(out["J_full"][0])


def test_rhc_batch_many_scenarios_match_single_runs():
    """Different scenarios with different numbers of rounds in one batch: each must equal its own solve_rhc run."""
    import dpilqr_b200 as dp
    from dpilqr_b200 import scenarios

    a, N, B = 3, 20, 6
    specs, x0, U0 = scenarios.quad12_batch(0, B, a, N)
    batch = dp.CompiledBatch(specs, N)
    kw = dict(n_d=3, step_size=4, dist_converge=0.4, t_diverge=2.0, n_lqr_iter=10)
    out = dp.solve_rhc_batch(batch, x0, radius=0.5, centralized=False, U0=U0, **kw)
    for k in range(B):
        dp._reset_ids()
        dyn = dp.MultiDynamicalModel([dp.QuadcopterDynamics12D(0.1, 100 + i) for i in range(a)])
        xf = specs[k].xf
        costs = [dp.ReferenceCost(xf[12 * i:12 * i + 12], np.eye(12), np.eye(4), 1000 * np.eye(12), 100 + i) for i in range(a)]
        prob = dp.ilqrProblem(dyn, dp.GameCost(costs, dp.ProximityCost([12] * a, 0.5, [3] * a)))
        X1, U1, J1 = dp.solve_rhc(prob, x0[k], N, 0.5, [], centralized=False, U0=U0[k], **kw)
        Xb, Ub = out["X_full"][k].cpu().numpy(), out["U_full"][k].cpu().numpy()
        assert X1.shape == Xb.shape, k
        assert np.array_equal(X1, Xb) and np.array_equal(U1, Ub) and J1 == float(out["J_full"][k]), k
    assert len(set(int(r) for r in out["rounds"].tolist())) >= 1


def test_trajectory_metrics_vs_numpy():
    import torch

    import dpilqr_b200 as dp

    rng = np.random.default_rng(3)
    B, rows, a, s = 5, 21, 4, 6
    X = rng.normal(size=(B, rows, a * s))
    m = dp.trajectory_metrics(torch.as_tensor(X).cuda(), a, s, 0.8, n_d=3)
    for k in range(B):
        d = dp.compute_pairwise_distance(X[k], [s] * a, 3)
        assert np.allclose(m["pairwise"][k].cpu().numpy(), d, rtol=1e-13)
        assert abs(float(m["min_separation"][k]) - d.min()) < 1e-13
        assert int(m["violations"][k]) == int((d < 0.8).sum())


@pytest.mark.parametrize("a,s,n_d,var,energy", [(10, 12, 3, 5.0, 30.0), (3, 12, 3, 1.5, 9.0), (15, 12, 3, 7.5, 45.0),
                                                (5, 4, 2, 3.0, 10.0), (9, 6, 3, 2.0, None), (2, 4, 2, 1.0, 2.0)])
def test_random_setup_on_the_device_is_bit_identical_to_the_host(a, s, n_d, var, energy):
    """Scenario generation on the device (reference util.py:125-217, random=True): the kernel runs NumPy's legacy
    MT19937 stream and NumPy's reduction orders, so x0 and xf equal the host's np.random.seed(k); random_setup(...) bit
    for bit, for every seed of the batch."""
    from dpilqr_b200 import scenarios
    from dpilqr_b200.util import random_setup

    first, count = 7, 300
    x0, xf = scenarios.random_setup_batch(first, count, a, s, n_d=n_d, var=var, energy=energy)
    x0, xf = x0.cpu().numpy(), xf.cpu().numpy()
    for k in range(count):
        np.random.seed(first + k)
        h0, hf = random_setup(a, s, is_rotation=False, rel_dist=a, var=var, n_d=n_d, random=True, energy=energy)
        assert np.array_equal(x0[k], h0.reshape(-1)) and np.array_equal(xf[k], hf.reshape(-1)), (k, np.max(np.abs(x0[k] - h0.reshape(-1))))


def test_metric_batch_built_on_the_device_solves_like_the_host_built_one():
    """quad12_batch_device: descriptor and inputs built by kernels and tensor ops only -- same solve, bit for bit."""
    import dpilqr_b200 as dp
    from dpilqr_b200 import scenarios

    specs, x0, U0 = scenarios.quad12_batch(40, 24, 10, 50)
    host = dp.CompiledBatch(specs, 50).solve(x0, U0)
    batch, x0d, U0d = scenarios.quad12_batch_device(40, 24, 10, 50)
    assert np.array_equal(x0d.cpu().numpy(), x0) and np.array_equal(U0d.cpu().numpy(), U0)
    dev = batch.solve(x0d, U0d)
    assert torch_equal(dev["X"], host["X"]) and torch_equal(dev["U"], host["U"]) and torch_equal(dev["iters"], host["iters"])


def torch_equal(a, b):
    import torch

    return bool(torch.equal(a, b))
