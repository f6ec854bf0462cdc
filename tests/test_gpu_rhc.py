"""Receding-horizon and warm-start drop-ins on the GPU against the CPU oracle on identical inputs
(reference distributed.py:106-221, problem.py:66-91).  Well-conditioned double-integrator teams, so the 1e-9 bar applies."""

import numpy as np
import pytest

from helpers import golden, oracle_problem, product_problem, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _case():
    return golden("solve_cfg1_dint4_a3.npz")


@pytest.mark.parametrize("centralized", [True, False])
def test_solve_rhc_matches_oracle(centralized):
    import dpilqr_b200 as dp
    from oracle import ilqr_oracle as O

    case = _case()
    N, step = 20, 3
    x0 = case["x0"]
    U_init = np.zeros((N, 6))
    kw = dict(n_d=2, step_size=step, dist_converge=0.3, t_diverge=6 * 0.1)
    Xo, Uo, Jo = O.solve_rhc(oracle_problem(case), x0, N, radius=0.5, ignore_ids=[], centralized=centralized, U_init=U_init, **kw)
    prob = product_problem(case)
    X, U, J = dp.solve_rhc(prob, x0, N, 0.5, [], centralized=centralized, U0=U_init, **kw)
    assert X.shape == Xo.shape and U.shape == Uo.shape
    assert rel_err(X, Xo) < TOL and rel_err(U, Uo) < TOL and abs(J - Jo) <= TOL * abs(Jo)


def test_solve_rhc_consumes_the_global_rng_like_the_reference():
    """Without U0 the warm start is np.random.rand(N, n_u) * 0.01 drawn from the global NumPy RNG (distributed.py:152)."""
    import dpilqr_b200 as dp
    from oracle import ilqr_oracle as O

    case = _case()
    N = 12
    kw = dict(n_d=2, step_size=2, dist_converge=0.5, t_diverge=0.3)
    np.random.seed(11)
    Xo, Uo, Jo = O.solve_rhc(oracle_problem(case), case["x0"], N, centralized=True, **kw)
    np.random.seed(11)
    X, U, J = dp.solve_rhc(product_problem(case), case["x0"], N, centralized=True, **kw)
    assert rel_err(X, Xo) < TOL and rel_err(U, Uo) < TOL
    assert np.random.rand() == np.random.RandomState(11).rand(N * 6 + 1)[-1]


def test_selfish_warmstart_matches_single_agent_solves():
    import dpilqr_b200 as dp
    from oracle import ilqr_oracle as O

    case = _case()
    prob = product_problem(case)
    N = int(case["N"])
    U_warm = prob.selfish_warmstart(case["x0"], N)
    full = oracle_problem(case)
    for i, id_ in enumerate(full.ids):
        sub = full.subproblem([id_])
        _, Ui, _ = O.OracleSolver(sub, N).solve(case["x0"][4 * i:4 * i + 4])
        assert rel_err(U_warm[:, 2 * i:2 * i + 2], Ui) < TOL


def test_solve_distributed_ignore_ids_and_info():
    import dpilqr_b200 as dp

    case = golden("dist_cfg2_uni4_a5_crowded.npz")
    prob = product_problem(case)
    ids = [int(v) for v in case["ids"]]
    X, U, J, info = dp.solve_distributed(prob, case["X_in"], case["U0"], float(case["radius_graph"]), [ids[1]], None, False)
    assert np.all(X[:, 4:8] == 0.0) and np.all(U[:, 2:4] == 0.0)     # ignored agents keep zero columns (distributed.py:50-63)
    assert ids[1] not in info and set(info) == set(ids) - {ids[1]}
    assert all(sorted(v[1]) == v[1] and k in v[1] for k, v in info.items())
    with pytest.raises(ValueError):
        dp.solve_distributed(prob, case["X_in"], case["U0"], 0.5, [999], None, False)  # distributed.py:38-39


def test_single_agent_reference_cost_problem():
    """ilqrProblem(model, ReferenceCost) without GameCost (reference scripts/examples.py:26-70)."""
    import dpilqr_b200 as dp
    from oracle import ilqr_oracle as O

    dt, N = 0.05, 30
    x0 = np.array([-1.0, 1.0, 0.5, 0.0])
    Q, Qf, R = np.diag([1.0, 1, 0, 0]), 1000 * np.eye(4), np.eye(2)
    prob = dp.ilqrProblem(dp.UnicycleDynamics4D(dt), dp.ReferenceCost(np.zeros((1, 4)), Q, R, Qf))
    X, U, J = dp.ilqrSolver(prob, N).solve(x0, verbose=False)
    oracle = O.OracleSolver(O.OracleProblem(["Unicycle4D"], dt, np.zeros(4), [Q], [R], [Qf], game=False), N)
    Xo, Uo, Jo = oracle.solve(x0)
    assert rel_err(X, Xo) < 1e-8 and rel_err(U, Uo) < 1e-8 and abs(J - Jo) <= 1e-8 * abs(Jo)


@pytest.mark.parametrize("a,seed", [(1, 3), (2, 3), (7, 3), (12, 5), (15, 3)])
def test_backward_pass_all_sizes_vs_oracle(a, seed):
    """Every code path of the backward kernel (tensor path for even agent counts with compile-time sizes, generic
    DFMA path, and the L2-scratch path for teams too large for shared memory) against the oracle's
    _backward_pass on the hover rollout of a random Quadcopter12D scenario."""
    import dpilqr_b200 as dp
    from dpilqr_b200 import scenarios
    from oracle import ilqr_oracle as O

    N = 50
    # (seed 3 with 12 agents is a crowded draw on which the reference's own gains amplify rounding 1e3-fold)
    x0, xf, U0 = scenarios.quad12_inputs(seed, max(a, 2), N)
    if a == 1:  # random_setup normalises a lone agent onto the origin (0/0): take agent 0 of the two-agent scenario
        x0, xf, U0 = x0[:12], xf[:12], U0[:, :4]
    spec = scenarios.quad12_spec(xf, a)
    batch = dp.CompiledBatch([spec], N)
    X, J = batch.rollout(x0[None], U0[None])
    stage, _ = batch.linearize_quadraticize(X, U0[None])
    K, d, st = batch.backward(stage, 1.0)
    prob = O.OracleProblem(["Quadcopter12D"] * a, 0.1, xf, np.eye(12), np.eye(4), 1000 * np.eye(12), 0.5, [3] * a,
                           [100 + i for i in range(a)])
    solver = O.OracleSolver(prob, N)
    Xo, Jo = solver.rollout(x0, U0)
    Ko, do = solver.backward_pass(Xo, U0)
    assert rel_err(X[0].cpu().numpy(), Xo) < 1e-12 and abs(float(J[0]) - Jo) <= 1e-12 * abs(Jo)
    assert rel_err(K[0].cpu().numpy(), Ko) < TOL and rel_err(d[0].cpu().numpy(), do) < TOL
    # and one full iteration through the solver loop (line search included)
    out = batch.solve(x0[None], U0[None], n_lqr_iter=1, trace=True)
    Xs, Us, Js = solver.solve(x0, U0.copy(), n_lqr_iter=1)
    assert int(out["trace_alpha"][0, 0]) == solver.trace[0]["alpha_index"]
    assert rel_err(out["X"][0].cpu().numpy(), Xs) < TOL and rel_err(out["U"][0].cpu().numpy(), Us) < TOL
