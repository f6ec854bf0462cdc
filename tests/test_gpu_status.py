"""Error behaviour of the reference reproduced from the per-problem status words: the Point.ndim assertion of
quadraticize_distance (reference cost.py:279, util.py:28-30), np.linalg.solve's LinAlgError for a singular Q_uu
(reference control.py:141-142), NaN from coincident agents (d == 0, cost.py:292).  Also the conditioning test of the
blocked LU of the metric path (lu.cuh pivot tie band) on crowded 10-drone scenarios."""

import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _quad6_pair(z0, z1, gap=0.2):
    import dpilqr_b200 as dp

    dp._reset_ids()
    dt, N = 0.05, 10
    dyn = dp.MultiDynamicalModel([dp.QuadcopterDynamics6D(dt, 0), dp.QuadcopterDynamics6D(dt, 1)])
    x0 = np.array([0.0, 0.0, z0, 0, 0, 0, gap, 0.0, z1, 0, 0, 0])
    xf = np.array([1.0, 0.0, 1.0, 0, 0, 0, -1.0, 0.0, 1.0, 0, 0, 0])
    costs = [dp.ReferenceCost(xf[6 * i:6 * i + 6], np.eye(6), np.eye(3), 100 * np.eye(6), i) for i in range(2)]
    prob = dp.ilqrProblem(dyn, dp.GameCost(costs, dp.ProximityCost([6, 6], 0.5, [3, 3])))
    U0 = np.tile([9.80665, 0, 0], (N, 2))
    return prob, x0, U0, N


def test_point_ndim_mismatch_raises_assertion_error():
    """Exactly one of two 3-D agents at z == 0.0: the reference asserts (cost.py:279) whatever the distance."""
    import dpilqr_b200 as dp
    from dpilqr_b200 import _native

    prob, x0, U0, N = _quad6_pair(0.0, 1.0)
    solver = dp.ilqrSolver(prob, N)
    with pytest.raises(AssertionError):
        solver.solve(x0, U0, verbose=False)
    X, _ = solver._rollout(x0, U0)
    with pytest.raises(AssertionError):
        solver._backward_pass(X, U0)
    # status word through the batched front door; on_error="raise" (default) raises like the reference
    spec = dp.spec_from_problem(prob)
    res = dp.solve_specs([spec], [x0], [U0], N, on_error="status")
    assert int(res[0]["status"]) & _native.ST_POINT_NDIM
    with pytest.raises(AssertionError):
        dp.solve_specs([spec], [x0], [U0], N)
    with pytest.raises(AssertionError):
        dp.solve_distributed(prob, x0.reshape(1, -1), U0, 0.5, [], None, False)
    # both at z == 0 (both "2-D" points) or both off it: no assertion
    for z in (0.0, 0.7):
        prob2, x02, U02, _ = _quad6_pair(z, z)
        dp.ilqrSolver(prob2, N).solve(x02, U02, verbose=False)


def test_singular_quu_raises_linalg_error():
    """HumanLin6D's third control has no effect (B column 2 == 0); with R = diag(1, 1, 0) Q_uu has a zero row and
    np.linalg.solve raises LinAlgError in the reference (control.py:141).  Also through solve_distributed, whose
    sub-problems all go through ilqrSolver.solve in the reference."""
    import dpilqr_b200 as dp
    from dpilqr_b200 import _native

    dp._reset_ids()
    dt, N = 0.05, 10
    R = np.diag([1.0, 1.0, 0.0])
    prob1 = dp.ilqrProblem(dp.HumanDynamicsLin6D(dt), dp.ReferenceCost(np.zeros(6), np.eye(6), R, 100 * np.eye(6)))
    with pytest.raises(np.linalg.LinAlgError):
        dp.ilqrSolver(prob1, N).solve(np.array([1.0, 1, 0, 0, 0, 0]), verbose=False)
    dp._reset_ids()
    dyn = dp.MultiDynamicalModel([dp.HumanDynamicsLin6D(dt, 0), dp.HumanDynamicsLin6D(dt, 1)])
    xf = np.array([1.0, 0, 0, 0, 0, 0, -1.0, 0, 0, 0, 0, 0])
    costs = [dp.ReferenceCost(xf[6 * i:6 * i + 6], np.eye(6), R, 100 * np.eye(6), i) for i in range(2)]
    prob = dp.ilqrProblem(dyn, dp.GameCost(costs, dp.ProximityCost([6, 6], 0.5, [2, 2])))
    x0 = np.array([-1.0, 0.1, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 0])
    with pytest.raises(np.linalg.LinAlgError):
        dp.solve_distributed(prob, x0.reshape(1, -1), np.zeros((N, 6)), 0.5, [], None, False)
    res, _ = dp.solve_distributed_batch([prob], [x0.reshape(1, -1)], [np.zeros((N, 6))], 0.5, None, on_error="status")
    assert res[0][0].shape == (N + 1, 12)
    out = dp.solve_specs([dp.spec_from_problem(prob)], [x0], [np.zeros((N, 6))], N, on_error="status")
    assert int(out[0]["status"]) & _native.ST_SINGULAR


def test_coincident_agents_give_nan_like_the_reference():
    """d == 0 between two agents: 2 (d - r) / d * 0 = NaN in the reference (cost.py:292); the solve does not raise,
    the line search fails on NaN costs and J comes back NaN (control.py:179-198, 225)."""
    import dpilqr_b200 as dp
    from dpilqr_b200 import _native

    prob, x0, U0, N = _quad6_pair(1.0, 1.0, gap=0.0)
    solver = dp.ilqrSolver(prob, N)
    X, U, J = solver.solve(x0, U0, verbose=False)
    assert np.isnan(J)
    assert solver.last_trace["status"] & _native.ST_NONFINITE and solver.last_trace["status"] & _native.ST_LS_FAILED
    assert solver.last_trace["iters"] == 1 and np.array_equal(U, U0)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_crowded_ten_drone_backward_pass_vs_oracle(seed):
    """The blocked LU of the metric path (a == 10 typed kernel) picks pivots within a 2^-13 tie band of the column
    maximum (lu.cuh).  Crowded scenarios (energy 10: drones inside each other's radius, Q_uu ill-conditioned and
    indefinite away from the first iterate) are where a pivoting difference would show: gains against the oracle's
    dgesv-based backward pass, error reported next to cond(Q_uu)."""
    import random

    import dpilqr_b200 as dp
    from dpilqr_b200 import scenarios
    from dpilqr_b200.util import random_setup
    from oracle import ilqr_oracle as O

    a, N = 10, 50
    np.random.seed(seed)
    random.seed(seed)
    x0, xf = random_setup(a, 12, is_rotation=False, rel_dist=a, var=a / 2, n_d=3, random=True, energy=10.0)
    x0, xf = x0.reshape(-1), xf.reshape(-1)
    U0 = np.tile([0.0, 0.0, 0.0, scenarios.HOVER_THRUST], (N, a))
    batch = dp.CompiledBatch([scenarios.quad12_spec(xf, a)], N)
    prob = O.OracleProblem(["Quadcopter12D"] * a, 0.1, xf, np.eye(12), np.eye(4), 1000 * np.eye(12), 0.5, [3] * a,
                           [100 + i for i in range(a)])
    solver = O.OracleSolver(prob, N)
    Xo, _ = solver.rollout(x0, U0)
    for mu in (1.0, 2.0 ** -6, 0.0):
        solver.mu = mu
        solver.cond_log = []
        Ko, do = solver.backward_pass(Xo, U0)
        cond = max(solver.cond_log)
        stage, _ = batch.linearize_quadraticize(Xo[None], U0[None])
        K, d, st = batch.backward(stage, mu)
        eK, ed = rel_err(K[0].cpu().numpy(), Ko), rel_err(d[0].cpu().numpy(), do)
        bar = max(1e-9, 50 * cond * 2.2e-16)  # two backward-stable LU solves agree to a few cond * eps
        print(f"crowded seed {seed} mu {mu:g}: max cond(Q_uu) {cond:.1e}  K err {eK:.1e}  d err {ed:.1e}  bar {bar:.1e}  status {int(st.item())}")
        assert np.isfinite(Ko).all()
        assert eK < bar and ed < bar
