"""Receding-horizon and warm-start drop-ins on the GPU against the CPU oracle on identical inputs
(reference distributed.py:106-221, problem.py:66-91).  Well-conditioned double-integrator teams, so the 1e-9 bar applies."""

import numpy as np
import pytest

from helpers import golden, graph_to_adj, oracle_problem, product_problem, rel_err

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


@pytest.mark.parametrize("a,seed", [(1, 3), (2, 3), (3, 1), (4, 3), (5, 2), (6, 3), (7, 3), (8, 3), (9, 3), (11, 1), (12, 5), (13, 1), (14, 5), (15, 3)])
def test_backward_pass_all_sizes_vs_oracle(a, seed):
    """Every code path of the backward kernel (small-problem kernels for up to five drones, tensor path with
    compile-time sizes for 6, 8, 10 drones and -- on the L2 scratch -- 12 and 14, generic DFMA path, generic L2-scratch
    path) against the oracle's _backward_pass on the hover rollout of a random Quadcopter12D scenario and on the
    iterate after one iLQR iteration (tilted drones, active proximity terms).  The stand-alone backward pass of an odd
    team takes the generic kernels; inside the solver loop (the one-iteration solve below, then two more iterations)
    odd teams of 5..15 run the tensor-path kernel of the next even size on records padded by a phantom agent."""
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
    if solver.trace[0]["alpha_index"] >= 0:  # the backward pass of the second iteration, from the oracle's iterate
        stage, _ = batch.linearize_quadraticize(Xs[None], Us[None])
        K, d, st = batch.backward(stage, solver.mu)
        K2, d2 = solver.backward_pass(Xs, Us)
        eK, ed = rel_err(K[0].cpu().numpy(), K2), rel_err(d[0].cpu().numpy(), d2)
        print(f"a={a}: second-iterate gains K err {eK:.1e} d err {ed:.1e}")
        assert eK < TOL and ed < TOL
    # three iterations through the solver loop: same accepted step sizes, same iterate
    out = batch.solve(x0[None], U0[None], n_lqr_iter=3, trace=True)
    X3, U3, J3 = solver.solve(x0, U0.copy(), n_lqr_iter=3)
    assert [int(v) for v in out["trace_alpha"][0, :int(out["iters"][0])]] == [r["alpha_index"] for r in solver.trace]
    e3 = max(rel_err(out["X"][0].cpu().numpy(), X3), rel_err(out["U"][0].cpu().numpy(), U3))
    print(f"a={a}: three iterations, X/U err {e3:.1e}")
    assert e3 < 1e-8


# --------------------------------------------------------------------------------------------
# Receding horizon on the BASELINE configurations, against runs of the UNMODIFIED reference
# (tests/golden/rhc_*.npz written by generate_golden.py gen_rhc; reference distributed.py:106-221)
# --------------------------------------------------------------------------------------------
def _rhc_names():
    import glob
    import os

    from helpers import GOLDEN

    return sorted(os.path.basename(p)[len("rhc_"):-4] for p in glob.glob(os.path.join(GOLDEN, "rhc_*.npz")))


def _rhc_tol(sens):
    return max(TOL, 1000.0 * float(sens))


@pytest.mark.parametrize("name", _rhc_names())
def test_solve_rhc_vs_reference_golden(name):
    """The whole receding-horizon run through the drop-in solve_rhc: config 1 (centralized and decentralised),
    config 3 as configured in the reference's example (2 x Quad6D + Human6D, dt 0.05, radius 0.3, n_dims [3,3,2],
    centralized=False, n_d=3, step_size=3, dist_converge=0.1, the reference's own seeded 0.01*rand warm start),
    config 4 (15 x Quad12D, dynamic interaction graphs, hover warm start, 4 rounds)."""
    import dpilqr_b200 as dp

    g = golden(f"rhc_{name}.npz")
    prob = product_problem(g)
    N = int(g["N"])
    kw = dict(centralized=bool(g["centralized"]), n_d=int(g["n_d"]), step_size=int(g["step_size"]),
              dist_converge=float(g["dist_converge"]), t_diverge=float(g["t_diverge"]),
              n_lqr_iter=int(g["n_lqr_iter"]), tol=float(g["tol"]))
    args = () if kw["centralized"] else (float(g["radius_graph"]), [])
    if bool(g["has_U_init"]):
        X, U, J = dp.solve_rhc(prob, g["x0"], N, *args, U0=g["U_init"], **kw)
    else:  # the reference's own draw from the global NumPy RNG (distributed.py:152)
        np.random.seed(int(g["seed"]))
        X, U, J = dp.solve_rhc(prob, g["x0"], N, *args, **kw)
    tol = _rhc_tol(g["sens_X"])
    print(f"rhc_{name}: reference sensitivity {float(g['sens_X']):.1e} -> bar {tol:.1e}; shapes {X.shape} vs {g['X_full'].shape}")
    if tol > 1e-3:
        # config 3: the reference's own run moves by O(1) under a 1e-15 perturbation of x0; only the structure of the
        # run is comparable as a whole -- every round is checked from the reference's iterate in the next test
        assert X.shape[1] == g["X_full"].shape[1] and U.shape[1] == g["U_full"].shape[1] and np.isfinite(J)
        return
    assert X.shape == g["X_full"].shape and U.shape == g["U_full"].shape
    ex, eu = rel_err(X, g["X_full"]), rel_err(U, g["U_full"])
    print(f"   achieved X {ex:.1e}  U {eu:.1e}  J {abs(J - float(g['J_full'])) / abs(float(g['J_full'])):.1e}")
    assert ex < tol and eu < tol
    assert abs(J - float(g["J_full"])) <= tol * abs(float(g["J_full"]))


@pytest.mark.parametrize("name", _rhc_names())
def test_rhc_rounds_from_the_reference_iterate(name):
    """Teacher-forced: every round of the reference's run is re-solved on the GPU from the reference's own round
    input (X, U warm start).  Interaction graphs must match exactly; trajectories at max(1e-9, 1000 x the reference's
    own sensitivity of that round), rounds whose bar would exceed 1e-3 are counted and reported, not compared."""
    import dpilqr_b200 as dp

    g = golden(f"rhc_{name}.npz")
    prob = product_problem(g)
    N = int(g["N"])
    ids = [int(v) for v in g["ids"]]
    solver_kw = dict(n_lqr_iter=int(g["n_lqr_iter"]), tol=float(g["tol"]))
    checked, skipped, worst = 0, 0, 0.0
    for k in range(len(g["round_J"])):
        X_in = g["round_X_in"][k, : int(g["round_rows_in"][k])]
        tol = _rhc_tol(g["round_sens_tf"][k])
        if bool(g["centralized"]):
            X, U, J = dp.ilqrSolver(prob, N).solve(X_in[0], g["round_U_in"][k].copy(), verbose=False, **solver_kw)
        else:
            X, U, J, info = dp.solve_distributed(prob, X_in, g["round_U_in"][k].copy(), float(g["radius_graph"]), [], None, False, **solver_kw)
            assert np.array_equal(graph_to_adj({i: v[1] for i, v in info.items()}, ids), g["round_adjacency"][k]), k
        if tol > 1e-3:
            skipped += 1
            continue
        err = max(rel_err(X, g["round_X"][k]), rel_err(U, g["round_U"][k]))
        worst = max(worst, err / tol)
        assert err < tol, (k, err, tol)
        if not bool(g["centralized"]):
            assert abs(J - g["round_J"][k]) <= tol * abs(g["round_J"][k]), k
        checked += 1
    print(f"rhc_{name}: {checked} rounds compared (worst error / bar = {worst:.2g}), {skipped} chaotic rounds only graph-checked")
    assert checked >= 1


def test_selfish_warmstart_vs_reference_golden():
    import dpilqr_b200 as dp

    g = golden("warmstart.npz")
    for tag in ("dint4", "uni4"):
        case = {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + "_")}
        prob = product_problem(case)
        U_warm = prob.selfish_warmstart(case["x0"], int(case["N"]))
        err = rel_err(U_warm, case["U_warm"])
        print(f"selfish_warmstart {tag}: {err:.1e}")
        assert err < (TOL if tag == "dint4" else 1e-6)  # single unicycles: 20+ iterations of an ill-conditioned solve


def test_receding_horizon_controller_vs_reference_golden():
    """RecedingHorizonController (reference control.py:253-326): generator protocol, shift of the warm start,
    J_converge stop, RuntimeError on a wrong warm-start shape."""
    import dpilqr_b200 as dp

    g = golden("warmstart.npz")
    case = {k[len("dint4") + 1:]: v for k, v in g.items() if k.startswith("dint4_")}
    prob = product_problem(case)
    N, step = int(g["rhc_N"]), int(g["rhc_step"])
    ctrl = dp.RecedingHorizonController(case["x0"].reshape(-1, 1), dp.ilqrSolver(prob, N), step_size=step)
    assert ctrl.N == N
    Xs, Us, Js = [], [], []
    for Xk, Uk, Jk in ctrl.solve(np.zeros((N, 6)), J_converge=float(g["rhc_J_converge"]), verbose=False):
        Xs.append(Xk), Us.append(Uk), Js.append(Jk)
        assert len(Js) <= len(g["rhc_J"])
    assert len(Js) == len(g["rhc_J"])
    assert rel_err(np.stack(Xs), g["rhc_X"]) < TOL and rel_err(np.stack(Us), g["rhc_U"]) < TOL
    assert rel_err(np.array(Js), g["rhc_J"]) < TOL
    with pytest.raises(RuntimeError):
        next(dp.RecedingHorizonController(case["x0"], dp.ilqrSolver(prob, N)).solve(np.zeros((N + 1, 6))))
