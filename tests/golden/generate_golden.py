#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference and oracle/_ref):

    bash oracle/build_ref.sh && python tests/golden/generate_golden.py

Everything written here is an *output of the reference* (reference dpilqr/*.py +
its compiled bbdynamics module) on seeded inputs; the fixtures are what pins the
oracle (oracle/ilqr_oracle.py) and, through it and directly, the CUDA path.
The per-iteration traces are obtained by wrapping ilqrSolver._backward_pass and
ilqrSolver._forward_pass at class level.
"""

import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

ref = load_reference()
OUT = os.path.dirname(os.path.abspath(__file__))
G = 9.80665

MODEL_CLASSES = {
    "DoubleInt4D": ref.DoubleIntDynamics4D,
    "DoubleInt6D": ref.DoubleIntDynamics6D,
    "Car3D": ref.CarDynamics3D,
    "Unicycle4D": ref.UnicycleDynamics4D,
    "Quadcopter6D": ref.QuadcopterDynamics6D,
    "Human6D": ref.HumanDynamics6D,
    "HumanLin6D": ref.HumanDynamicsLin6D,
    "Quadcopter12D": ref.QuadcopterDynamics12D,
    "Bike5D": ref.BikeDynamics5D,
}


# --------------------------------------------------------------------------- #
# instrumentation
# --------------------------------------------------------------------------- #
class Trace:
    def __init__(self):
        self.records = []

    def __enter__(self):
        cls = ref.ilqrSolver
        self._bp, self._fp, self._solve = cls._backward_pass, cls._forward_pass, cls.solve
        trace = self
        self.n_solves = 0

        def solve(solver, *a, **kw):
            trace.n_solves += 1
            return trace._solve(solver, *a, **kw)

        def bp(solver, X, U):
            K, d = trace._bp(solver, X, U)
            trace.records.append({"solver": trace.n_solves, "mu": solver.μ, "K": K, "d": d, "J": [], "alpha": [],
                                  "X": X.copy(), "U": U.copy()})
            return K, d

        def fp(solver, X, U, K, d, α):
            out = trace._fp(solver, X, U, K, d, α)
            trace.records[-1]["J"].append(out[2])
            trace.records[-1]["alpha"].append(float(α))
            return out

        cls._backward_pass, cls._forward_pass, cls.solve = bp, fp, solve
        return self

    def __exit__(self, *exc):
        ref.ilqrSolver._backward_pass, ref.ilqrSolver._forward_pass, ref.ilqrSolver.solve = self._bp, self._fp, self._solve


def trace_arrays(records, J0):
    """mu per iteration, tried-J table (NaN padded), accepted alpha index (-1: failed search)."""
    n = len(records)
    mu = np.array([r["mu"] for r in records])
    Jt = np.full((n, 10), np.nan)
    acc = np.full(n, -1, dtype=np.int64)
    J_star = J0
    for i, r in enumerate(records):
        Jt[i, : len(r["J"])] = r["J"]
        if r["J"][-1] < J_star:
            acc[i] = len(r["J"]) - 1
            J_star = r["J"][-1]
    return mu, Jt, acc


# --------------------------------------------------------------------------- #
# problem construction (mirrors scripts/analysis.py:35-79 and scripts/examples.py)
# --------------------------------------------------------------------------- #
def build_problem(models, dt, xf, Q, R, Qf, radius, n_dims, ids):
    a = len(models)
    dyn = ref.MultiDynamicalModel([MODEL_CLASSES[m](dt, id_) for m, id_ in zip(models, ids)])
    s = dyn.x_dims[0]
    x_dims = [s] * a
    costs = [
        ref.ReferenceCost(xf_i, Q[i].copy(), R[i].copy(), Qf[i].copy(), id_)
        for i, (xf_i, id_) in enumerate(zip(ref.split_agents_gen(xf.flatten(), x_dims), ids))
    ]
    prox = ref.ProximityCost(x_dims, radius, list(n_dims))
    return ref.ilqrProblem(dyn, ref.GameCost(costs, prox))


def case_dict(models, dt, N, x0, xf, Q, R, Qf, radius, n_dims, ids, U0, n_lqr_iter, tol):
    return dict(
        models=np.array(models), dt=dt, N=N, x0=np.asarray(x0, float).flatten(), xf=np.asarray(xf, float).flatten(),
        Q=np.stack(Q), R=np.stack(R), Qf=np.stack(Qf), radius=radius, n_dims=np.array(n_dims), ids=np.array(ids),
        U0=U0, n_lqr_iter=n_lqr_iter, tol=tol,
    )


def accepted_costs(records, J0):
    out, J_star = [], J0
    for r in records:
        if r["J"][-1] < J_star:
            J_star = r["J"][-1]
        out.append(J_star)
    return np.array(out)


def run_centralized(name, case, keep_K_steps=(0,)):
    ref._reset_ids()
    prob = build_problem(list(case["models"]), case["dt"], case["xf"], case["Q"], case["R"], case["Qf"],
                         case["radius"], case["n_dims"], list(case["ids"]))
    solver = ref.ilqrSolver(prob, case["N"])
    X0, J0 = solver._rollout(case["x0"].reshape(-1, 1), case["U0"])
    with Trace() as tr:
        X, U, J = solver.solve(case["x0"].reshape(-1, 1), case["U0"].copy(), n_lqr_iter=int(case["n_lqr_iter"]),
                               tol=float(case["tol"]), verbose=False)
    mu, Jt, acc = trace_arrays(tr.records, J0)
    # Sensitivity of THE REFERENCE ITSELF to a one-ulp-sized input perturbation (x0 * (1 +- 1e-15)): how far the
    # accepted cost of each iteration and the final trajectory move.  iLQR on the ill-conditioned configs is chaotic
    # (error x10 per iteration), so this is the yardstick for what any re-implementation can reproduce.
    signs = np.sign(np.random.default_rng(2024).normal(size=case["x0"].shape))
    with Trace() as tr2:
        X2, U2, J2 = solver.solve((case["x0"] * (1 + 1e-15 * signs)).reshape(-1, 1), case["U0"].copy(),
                                  n_lqr_iter=int(case["n_lqr_iter"]), tol=float(case["tol"]), verbose=False)
    Ja, Jb = accepted_costs(tr.records, J0), accepted_costs(tr2.records, J0)
    k = min(len(Ja), len(Jb))
    sens_J = np.full(len(Ja), np.inf)
    sens_J[:k] = np.abs(Ja[:k] - Jb[:k]) / np.abs(Ja[:k])
    sens_X = float(np.max(np.abs(X - X2)) / np.max(np.abs(X)))
    sens_U = float(np.max(np.abs(U - U2)) / np.max(np.abs(U)))
    out = dict(case)
    out.update(X0=X0, J0=J0, X=X, U=U, J=J, trace_mu=mu, trace_J=Jt, trace_alpha=acc,
               iter_X=np.stack([r["X"] for r in tr.records]), iter_U=np.stack([r["U"] for r in tr.records]),
               sens_J=sens_J, sens_X=sens_X, sens_U=sens_U,
               K_first=tr.records[0]["K"][list(keep_K_steps)], K_first_steps=np.array(keep_K_steps),
               d_first=tr.records[0]["d"], K_last_iter=tr.records[-1]["K"][list(keep_K_steps)], d_last_iter=tr.records[-1]["d"])
    np.savez_compressed(os.path.join(OUT, f"solve_{name}.npz"), **out)
    print(f"solve_{name}: iters={len(mu)} acc={acc.tolist()} J0={J0:.6g} J={J:.6g} sens_X={sens_X:.1e} sens_Jmax={sens_J.max():.1e}")
    return prob


def run_distributed(name, case, Xin, radius_graph, ignore_ids=()):
    ref._reset_ids()
    prob = build_problem(list(case["models"]), case["dt"], case["xf"], case["Q"], case["R"], case["Qf"],
                         case["radius"], case["n_dims"], list(case["ids"]))
    ids = list(case["ids"])
    x_dims = prob.game_cost.x_dims
    graph = ref.define_inter_graph_threshold(Xin, radius_graph, x_dims, ids)
    with Trace() as tr:
        X, U, J, info = ref.solve_distributed(prob, Xin, case["U0"].copy(), radius_graph, list(ignore_ids), None, False,
                                              n_lqr_iter=int(case["n_lqr_iter"]), tol=float(case["tol"]))
    # iterations per subproblem, in agent order (solver objects are created in order)
    order, iters = [], {}
    for r in tr.records:
        if r["solver"] not in iters:
            order.append(r["solver"])
            iters[r["solver"]] = 0
        iters[r["solver"]] += 1
    adj = np.zeros((len(ids), len(ids)), dtype=np.int8)
    for i, id_ in enumerate(ids):
        for other in graph[id_]:
            adj[i, ids.index(int(other))] = 1
    signs = np.sign(np.random.default_rng(2024).normal(size=Xin.shape))
    X2, U2, J2, _ = ref.solve_distributed(prob, Xin * (1 + 1e-15 * signs), case["U0"].copy(), radius_graph, list(ignore_ids), None,
                                          False, n_lqr_iter=int(case["n_lqr_iter"]), tol=float(case["tol"]))
    sens_X = float(np.max(np.abs(X - X2)) / np.max(np.abs(X)))
    sens_U = float(np.max(np.abs(U - U2)) / np.max(np.abs(U)))
    out = dict(case)
    out.update(sens_X=sens_X, sens_U=sens_U, sens_J=abs(J - J2) / abs(J), X_in=Xin, radius_graph=radius_graph, ignore_ids=np.array(list(ignore_ids), dtype=np.int64), adjacency=adj,
               X_dec=X, U_dec=U, J_full=J, sub_iters=np.array([iters[s] for s in order]))
    np.savez_compressed(os.path.join(OUT, f"dist_{name}.npz"), **out)
    print(f"dist_{name}: graph sizes={adj.sum(1).tolist()} sub_iters={out['sub_iters'].tolist()} J_full={J:.6g} sens_X={sens_X:.1e}")


def random_case(model, a, seed, energy, n_d, Q, R, Qf, dt=0.1, N=50, radius=0.5, U0=None, n_dims=None,
                n_lqr_iter=50, tol=1e-3):
    """random_setup exactly as scripts/analysis.py:45-54 / SURVEY.md section 8d."""
    ref._reset_ids()
    np.random.seed(seed)
    random.seed(seed)
    s = MODEL_CLASSES[model](dt).n_x
    c = MODEL_CLASSES[model](dt).n_u
    x0, xf = ref.random_setup(a, s, is_rotation=False, rel_dist=a, var=a / 2, n_d=n_d, random=True, energy=energy)
    ids = [100 + i for i in range(a)]
    if U0 is None:
        U0 = np.zeros((N, a * c))
    return case_dict([model] * a, dt, N, x0, xf, [Q] * a, [R] * a, [Qf] * a, radius,
                     n_dims if n_dims is not None else [n_d] * a, ids, U0, n_lqr_iter, tol)


def hover(N, a):
    return np.tile([0, 0, 0, G * 63 / 2000], (N, a))


# --------------------------------------------------------------------------- #
def gen_dynamics():
    rng = np.random.default_rng(1234)
    out = {}
    for name, cls in MODEL_CLASSES.items():
        model = cls(0.1)
        nx, nu = model.n_x, model.n_u
        xs = rng.normal(size=(16, nx)) * 0.6
        us = rng.normal(size=(16, nu)) * 0.4
        if name == "Quadcopter12D":
            us[:, 3] += G * 63 / 2000
            us[:, :3] *= 0.01
        f = np.stack([np.asarray(model.f(x.copy(), u.copy()), dtype=float).flatten() for x, u in zip(xs, us)])
        xn = np.stack([np.asarray(model(x.copy(), u.copy()), dtype=float).flatten() for x, u in zip(xs, us)])
        AB = [model.linearize(x.copy(), u.copy()) for x, u in zip(xs, us)]
        out[f"{name}_x"], out[f"{name}_u"] = xs, us
        out[f"{name}_f"], out[f"{name}_xn"] = f, xn
        out[f"{name}_A"] = np.stack([np.asarray(ab[0], dtype=float) for ab in AB])
        out[f"{name}_B"] = np.stack([np.asarray(ab[1], dtype=float) for ab in AB])
    # survey-time golden vector, Quad12D
    x = np.array([.3, -.2, 1.1, .05, -.04, .03, .5, -.3, .2, .1, -.2, .15])
    u = np.array([.01, -.02, .005, .31])
    out["survey_quad12_xn"] = ref.integrate(x, u, 0.1, ref.Model.Quadcopter12D)
    np.savez_compressed(os.path.join(OUT, "dynamics.npz"), **out)
    print("dynamics.npz written")


def gen_cost():
    """GameCost value + quadraticisation on crowded random points (cost.py:197-239)."""
    rng = np.random.default_rng(99)
    out = {}
    specs = {
        # name: (models, n_dims, radius)
        "quad12_3d": (["Quadcopter12D"] * 4, [3] * 4, 0.5),
        "unicycle_2d": (["Unicycle4D"] * 5, [2] * 5, 0.5),
        "hetero_q6h6": (["Quadcopter6D", "Quadcopter6D", "Human6D"], [3, 3, 2], 0.3),
    }
    for name, (models, n_dims, radius) in specs.items():
        ref._reset_ids()
        a = len(models)
        s = MODEL_CLASSES[models[0]](0.1).n_x
        c = MODEL_CLASSES[models[0]](0.1).n_u
        Q = rng.normal(size=(a, s, s))  # deliberately dense and asymmetric (cost.py:60-61)
        R = rng.normal(size=(a, c, c))
        Qf = rng.normal(size=(a, s, s))
        xf = rng.normal(size=a * s)
        prob = build_problem(models, 0.1, xf, Q, R, Qf, radius, n_dims, [100 + i for i in range(a)])
        xs = rng.normal(size=(12, a * s)) * 0.25  # crowded: many pairs inside the radius
        us = rng.normal(size=(12, a * c))
        gc = prob.game_cost
        out[f"{name}_Q"], out[f"{name}_R"], out[f"{name}_Qf"], out[f"{name}_xf"] = Q, R, Qf, xf
        out[f"{name}_x"], out[f"{name}_u"] = xs, us
        out[f"{name}_models"], out[f"{name}_n_dims"], out[f"{name}_radius"] = np.array(models), np.array(n_dims), radius
        for term in (False, True):
            tag = "T" if term else "R"
            out[f"{name}_L{tag}"] = np.array([np.asarray(gc(x, u, term)).item() for x, u in zip(xs, us)])
            quads = [gc.quadraticize(x, u, term) for x, u in zip(xs, us)]
            for k, key in enumerate(["Lx", "Lu", "Lxx", "Luu", "Lux"]):
                out[f"{name}_{key}{tag}"] = np.stack([np.asarray(q[k], dtype=float) for q in quads])
    np.savez_compressed(os.path.join(OUT, "cost.npz"), **out)
    print("cost.npz written")


def gen_graphs():
    """define_inter_graph_threshold (distributed.py:224-247) on random trajectories."""
    rng = np.random.default_rng(7)
    out = {}
    k = 0
    for a, s, rows in [(2, 4, 1), (5, 4, 51), (10, 12, 51), (15, 12, 51), (7, 6, 1), (4, 6, 7), (12, 12, 23), (6, 3, 100)]:
        for rep in range(3):
            X = np.cumsum(rng.normal(size=(rows, a * s)) * 0.15, axis=0) + rng.normal(size=(1, a * s)) * 1.2
            radius = float(rng.uniform(0.2, 0.9))
            ids = [100 + i for i in range(a)]
            graph = ref.define_inter_graph_threshold(X, radius, [s] * a, ids)
            adj = np.zeros((a, a), dtype=np.int8)
            for i, id_ in enumerate(ids):
                for other in graph[id_]:
                    adj[i, ids.index(int(other))] = 1
            out[f"g{k}_X"], out[f"g{k}_radius"], out[f"g{k}_s"], out[f"g{k}_adj"] = X, radius, s, adj
            k += 1
    out["count"] = k
    np.savez_compressed(os.path.join(OUT, "graphs.npz"), **out)
    print(f"graphs.npz written ({k} cases)")


def gen_solves():
    I = np.eye
    # config 1: 3 x DoubleInt4D centralized (scripts/analysis.py:35-79)
    c1 = random_case("DoubleInt4D", 3, 0, 10.0, 2, np.diag([1.0, 1, 0, 0]), I(2), 1000 * I(4))
    run_centralized("cfg1_dint4_a3", c1)
    # config 2: 5 x Unicycle4D, centralized + DP-iLQR split
    c2 = random_case("Unicycle4D", 5, 1, 10.0, 2, np.diag([1.0, 1, 0, 0]), I(2), 1000 * I(4))
    prob = run_centralized("cfg2_uni4_a5", c2)
    run_distributed("cfg2_uni4_a5_x0", c2, c2["x0"].reshape(1, -1), 0.5)
    Xroll, _ = ref.ilqrSolver(prob, 50)._rollout(c2["x0"].reshape(-1, 1), c2["U0"])
    run_distributed("cfg2_uni4_a5_traj", c2, Xroll, 0.5)
    # crowded unicycles so the graph has real structure
    c2b = random_case("Unicycle4D", 5, 3, 4.0, 2, np.diag([1.0, 1, 0, 0]), I(2), 1000 * I(4))
    run_distributed("cfg2_uni4_a5_crowded", c2b, c2b["x0"].reshape(1, -1), 0.5)
    # config 3: 2 x Quad6D + 1 x Human6D (scripts/examples.py:73-131, scenarios.py:145-152)
    x0 = np.array([-1.5, 0.1, 1, 0, 0, 0, 1.5, 0, 1, 0, 0, 0, 0, -1, 1.5, 0, 0, 0.0])
    xf = np.array([1.5, 0, 2, 0, 0, 0, -1.5, 0, 2, 0, 0, 0, 0.0, 2, 1.5, 0, 0, 0])
    Qq, Rq, Qfq = np.diag([1.0, 1, 1, 5, 5, 5]), np.diag([1.0, 1, 1]), 1e3 * I(6)
    Qh, Rh = np.diag([1.0, 1, 1, 0, 0, 0]), np.diag([1, 1, 1e-9])
    U0 = np.c_[np.tile([G, 0, 0], (50, 2)), np.ones((50, 3))]
    c3 = case_dict(["Quadcopter6D", "Quadcopter6D", "Human6D"], 0.05, 50, x0, xf, [Qq, Qq, Qh], [Rq, Rq, Rh],
                   [Qfq, Qfq, Qfq], 0.3, [3, 3, 2], [100, 101, 102], U0, 50, 1e-3)
    run_centralized("cfg3_q6q6h6", c3)
    run_distributed("cfg3_q6q6h6_x0", c3, x0.reshape(1, -1), 0.3)
    run_distributed("cfg3_q6q6h6_wide", c3, x0.reshape(1, -1), 1.2)
    # metric family: a x Quad12D, hover warm start, energy 3a
    for a, seeds in [(3, (0, 1)), (5, (0,)), (10, (0, 1, 2))]:
        for seed in seeds:
            c = random_case("Quadcopter12D", a, seed, 3.0 * a, 3, I(12), I(4), 1000 * I(12), U0=hover(50, a))
            prob = run_centralized(f"quad12_a{a}_s{seed}", c)
            if seed == 0:
                Xh, _ = ref.ilqrSolver(prob, 50)._rollout(c["x0"].reshape(-1, 1), c["U0"])
                run_distributed(f"quad12_a{a}_s{seed}", c, Xh, 0.5)
    # config 4: 15 x Quad12D, one decentralised round on the hover rollout
    c4 = random_case("Quadcopter12D", 15, 0, 45.0, 3, I(12), I(4), 1000 * I(12), U0=hover(50, 15), n_lqr_iter=8)
    prob = build_problem(list(c4["models"]), 0.1, c4["xf"], c4["Q"], c4["R"], c4["Qf"], 0.5, [3] * 15, list(c4["ids"]))
    Xh, _ = ref.ilqrSolver(prob, 50)._rollout(c4["x0"].reshape(-1, 1), c4["U0"])
    run_distributed("cfg4_quad12_a15", c4, Xh, 0.5)
    # single-agent GameCost problems of every remaining model class
    for model, Q, R, Qf, dt in [
        ("Car3D", I(3), I(2), 100 * I(3), 0.05),
        ("DoubleInt6D", I(6), I(3), 1000 * I(6), 0.05),
        ("HumanLin6D", I(6), 0.1 * I(3), 1e4 * I(6), 0.05),
        ("Bike5D", np.diag([1.0, 1, 0, 0, 0]), I(2), 1000 * I(5), 0.05),
    ]:
        c = random_case(model, 3, 5, 6.0, 2, Q, R, Qf, dt=dt, N=40)
        run_centralized(f"misc_{model}_a3", c)


# --------------------------------------------------------------------------- #
# receding horizon (distributed.py:106-221), selfish warm start (problem.py:66-91),
# RecedingHorizonController (control.py:253-326)
# --------------------------------------------------------------------------- #
class _FixedWarmStart:
    """Stands in for ``np.random.rand(N, n_u)`` inside the reference's solve_rhc (distributed.py:152): the product
    ``* 0.01`` returns the wanted warm start, so the UNMODIFIED reference runs from e.g. hover controls."""

    def __init__(self, U):
        self.U = U

    def __mul__(self, _):
        return self.U.copy()


class RoundRecorder:
    """Wraps the reference's solve_distributed / solve_centralized (looked up as module globals by solve_rhc) and
    records every round's inputs and outputs."""

    def __init__(self):
        self.rounds = []

    def __enter__(self):
        mod = sys.modules[ref.solve_rhc.__module__]
        self.mod, self._sd, self._sc = mod, mod.solve_distributed, mod.solve_centralized
        rec = self

        def sd(problem, X, U, *args, **kw):
            out = rec._sd(problem, X, U, *args, **kw)
            rec.rounds.append(dict(X_in=np.array(X), U_in=np.array(U), X=out[0].copy(), U=out[1].copy(), J=out[2],
                                   graph={k: list(v[1]) for k, v in out[3].items()}))
            return out

        def sc(solver, xi, U, ids, verbose, **kw):
            out = rec._sc(solver, xi, U, ids, verbose, **kw)
            rec.rounds.append(dict(X_in=np.array(xi).reshape(1, -1), U_in=np.array(U), X=out[0].copy(), U=out[1].copy(), J=out[2],
                                   graph={}))
            return out

        mod.solve_distributed, mod.solve_centralized = sd, sc
        return self

    def __exit__(self, *exc):
        self.mod.solve_distributed, self.mod.solve_centralized = self._sd, self._sc


def run_rhc(name, case, N, radius_graph, centralized, rhc_kw, U_init=None, seed=None, solver_kw=None):
    """One full solve_rhc of the unmodified reference + the same run from x0 * (1 +- 1e-15) (sensitivity)."""
    solver_kw = solver_kw or {}
    ids = list(case["ids"])

    def once(x0):
        ref._reset_ids()
        prob = build_problem(list(case["models"]), case["dt"], case["xf"], case["Q"], case["R"], case["Qf"],
                             case["radius"], case["n_dims"], ids)
        if seed is not None:
            np.random.seed(seed)
        saved = np.random.rand
        if U_init is not None:
            np.random.rand = lambda *shape: _FixedWarmStart(U_init)
        try:
            with RoundRecorder() as rr:
                args = () if centralized else (radius_graph, [])
                X, U, J = ref.solve_rhc(prob, x0.reshape(-1, 1), N, *args, centralized=centralized, **rhc_kw, **solver_kw)
        finally:
            np.random.rand = saved
        return X, U, J, rr.rounds, prob

    X, U, J, rounds, prob = once(case["x0"])
    signs = np.sign(np.random.default_rng(2024).normal(size=case["x0"].shape))
    X2, U2, J2, rounds2, _ = once(case["x0"] * (1 + 1e-15 * signs))
    same = X.shape == X2.shape
    sens_X = float(np.max(np.abs(X - X2)) / np.max(np.abs(X))) if same else np.inf
    sens_U = float(np.max(np.abs(U - U2)) / np.max(np.abs(U))) if same else np.inf
    R = len(rounds)
    a = len(ids)
    adj = np.zeros((R, a, a), dtype=np.int8)
    rows_in = np.array([r["X_in"].shape[0] for r in rounds])
    Xin = np.zeros((R, N + 1, X.shape[1]))
    round_sens = np.full(R, np.inf)
    for k, r in enumerate(rounds):
        Xin[k, : rows_in[k]] = r["X_in"]
        for i, id_ in enumerate(ids):
            for other in r["graph"].get(id_, []):
                adj[k, i, ids.index(int(other))] = 1
        if k < len(rounds2) and rounds2[k]["X"].shape == r["X"].shape:
            round_sens[k] = max(np.max(np.abs(r["X"] - rounds2[k]["X"])) / np.max(np.abs(r["X"])),
                                np.max(np.abs(r["U"] - rounds2[k]["U"])) / max(np.max(np.abs(r["U"])), 1e-300))
    # teacher-forced sensitivity of every round: the reference's own round solve from the SAME recorded input,
    # perturbed by 1e-15 relative (first row of X_in: only X_in[0] enters the sub-problem solves)
    round_sens_tf = np.full(R, np.inf)
    for k, r in enumerate(rounds):
        sg = np.sign(np.random.default_rng(7 + k).normal(size=r["X_in"].shape))
        Xp = r["X_in"] * (1 + 1e-15 * sg)
        if centralized:
            ref._reset_ids()
            Xk, Uk, _ = ref.ilqrSolver(prob, N).solve(Xp.reshape(-1, 1), r["U_in"].copy(), verbose=False, **solver_kw)
        else:
            Xk, Uk, _, _ = ref.solve_distributed(prob, Xp, r["U_in"].copy(), radius_graph, [], None, False, **solver_kw)
        round_sens_tf[k] = max(np.max(np.abs(r["X"] - Xk)) / np.max(np.abs(r["X"])),
                               np.max(np.abs(r["U"] - Uk)) / max(np.max(np.abs(r["U"])), 1e-300))
    out = dict(case)
    out.pop("U0", None)
    out.update(round_sens_tf=round_sens_tf)
    out.update(N=N, radius_graph=radius_graph if radius_graph is not None else 0.0, centralized=centralized,
               n_d=rhc_kw.get("n_d", 2), step_size=rhc_kw.get("step_size", 1),
               dist_converge=rhc_kw.get("dist_converge") or 0.0, J_converge=rhc_kw.get("J_converge") or 0.0,
               t_diverge=rhc_kw.get("t_diverge") or 0.0, seed=-1 if seed is None else seed,
               has_U_init=U_init is not None, U_init=U_init if U_init is not None else rounds[0]["U_in"],
               n_lqr_iter=solver_kw.get("n_lqr_iter", 50), tol=solver_kw.get("tol", 1e-3),
               X_full=X, U_full=U, J_full=J, sens_X=sens_X, sens_U=sens_U, sens_J=abs(J - J2) / abs(J),
               round_rows_in=rows_in, round_X_in=Xin, round_U_in=np.stack([r["U_in"] for r in rounds]),
               round_X=np.stack([r["X"] for r in rounds]), round_U=np.stack([r["U"] for r in rounds]),
               round_J=np.array([r["J"] for r in rounds]), round_adjacency=adj, round_sens=round_sens)
    np.savez_compressed(os.path.join(OUT, f"rhc_{name}.npz"), **out)
    print(f"rhc_{name}: rounds={R} X_full{X.shape} J_full={J:.6g} sens_X={sens_X:.1e} round_sens_max={round_sens.max():.1e} "
          f"teacher-forced round sens={np.array2string(round_sens_tf, precision=1)}")


def gen_rhc():
    I = np.eye
    # config 1 family: 3 x DoubleInt4D, centralized and decentralised receding horizon (well conditioned)
    c1 = random_case("DoubleInt4D", 3, 0, 10.0, 2, np.diag([1.0, 1, 0, 0]), I(2), 1000 * I(4))
    kw = dict(n_d=2, step_size=3, dist_converge=0.3, t_diverge=6 * 0.1)
    run_rhc("cfg1_dint4_a3_central", c1, 20, None, True, kw, seed=11)
    run_rhc("cfg1_dint4_a3_dec", c1, 20, 0.5, False, kw, seed=11)
    # (J_converge mode cannot be pinned: the unmodified reference raises NameError at distributed.py:131/189 there,
    # because n_agents / n_states are only bound in the dist_converge branch :141-142)
    # config 3 as configured (scripts/examples.py:86-130, scenarios.py:145-152): 2 x Quad6D + Human6D, decentralised,
    # n_d=3, step_size=3, dist_converge=0.1, the reference's own 0.01*rand warm start
    x0 = np.array([-1.5, 0.1, 1, 0, 0, 0, 1.5, 0, 1, 0, 0, 0, 0, -1, 1.5, 0, 0, 0.0])
    xf = np.array([1.5, 0, 2, 0, 0, 0, -1.5, 0, 2, 0, 0, 0, 0.0, 2, 1.5, 0, 0, 0])
    Qq, Rq, Qfq = np.diag([1.0, 1, 1, 5, 5, 5]), np.diag([1.0, 1, 1]), 1e3 * I(6)
    Qh, Rh = np.diag([1.0, 1, 1, 0, 0, 0]), np.diag([1, 1, 1e-9])
    c3 = case_dict(["Quadcopter6D", "Quadcopter6D", "Human6D"], 0.05, 50, x0, xf, [Qq, Qq, Qh], [Rq, Rq, Rh],
                   [Qfq, Qfq, Qfq], 0.3, [3, 3, 2], [100, 101, 102], np.zeros((50, 9)), 50, 1e-3)
    run_rhc("cfg3_q6q6h6_dec", c3, 50, 0.3, False, dict(n_d=3, step_size=3, dist_converge=0.1, t_diverge=50 * 0.05), seed=0)
    # config 4: 15 x Quad12D decentralised receding horizon with dynamic interaction graphs, hover warm start, 4 rounds
    c4 = random_case("Quadcopter12D", 15, 0, 45.0, 3, I(12), I(4), 1000 * I(12), U0=hover(50, 15))
    run_rhc("cfg4_quad12_a15_dec", c4, 50, 0.5, False, dict(n_d=3, step_size=5, dist_converge=0.1, t_diverge=3 * 5 * 0.1),
            U_init=hover(50, 15), solver_kw=dict(n_lqr_iter=8))


def gen_warmstart():
    """ilqrProblem.selfish_warmstart (problem.py:66-91) and RecedingHorizonController (control.py:253-326)."""
    import contextlib
    import io

    I = np.eye
    out = {}
    for tag, case in [("dint4", random_case("DoubleInt4D", 3, 0, 10.0, 2, np.diag([1.0, 1, 0, 0]), I(2), 1000 * I(4))),
                      ("uni4", random_case("Unicycle4D", 5, 1, 10.0, 2, np.diag([1.0, 1, 0, 0]), I(2), 1000 * I(4)))]:
        ref._reset_ids()
        prob = build_problem(list(case["models"]), case["dt"], case["xf"], case["Q"], case["R"], case["Qf"],
                             case["radius"], case["n_dims"], list(case["ids"]))
        with contextlib.redirect_stdout(io.StringIO()):
            U_warm = prob.selfish_warmstart(case["x0"].reshape(-1, 1), int(case["N"]))
        for k, v in case.items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_U_warm"] = U_warm
    # RecedingHorizonController on the 3 x DoubleInt4D problem
    case = random_case("DoubleInt4D", 3, 0, 10.0, 2, np.diag([1.0, 1, 0, 0]), I(2), 1000 * I(4))
    ref._reset_ids()
    prob = build_problem(list(case["models"]), case["dt"], case["xf"], case["Q"], case["R"], case["Qf"],
                         case["radius"], case["n_dims"], list(case["ids"]))
    N, step = 15, 2
    ctrl = ref.RecedingHorizonController(case["x0"].reshape(-1, 1), ref.ilqrSolver(prob, N), step_size=step)
    Xs, Us, Js = [], [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for k, (Xk, Uk, Jk) in enumerate(ctrl.solve(np.zeros((N, 6)), J_converge=250.0, verbose=False)):
            Xs.append(Xk), Us.append(Uk), Js.append(Jk)
            if k >= 7:
                break
    out.update(rhc_N=N, rhc_step=step, rhc_J_converge=250.0, rhc_X=np.stack(Xs), rhc_U=np.stack(Us), rhc_J=np.array(Js))
    np.savez_compressed(os.path.join(OUT, "warmstart.npz"), **out)
    print(f"warmstart.npz written (RecedingHorizonController horizons={len(Js)}, J={Js})")


if __name__ == "__main__":
    which = sys.argv[1:] or ["dynamics", "cost", "graphs", "solves", "rhc", "warmstart"]
    if "dynamics" in which:
        gen_dynamics()
    if "cost" in which:
        gen_cost()
    if "graphs" in which:
        gen_graphs()
    if "solves" in which:
        gen_solves()
    if "rhc" in which:
        gen_rhc()
    if "warmstart" in which:
        gen_warmstart()
