"""DP-iLQR orchestration (reference dpilqr/distributed.py) on the batched CUDA engine.

The interaction graph is a CUDA kernel (bit-exact neighbourhoods); all sub-problems of a
round -- of one scenario or of thousands -- are binned by neighbourhood size and solved as
a few batched launches instead of one Python solve per agent.
"""

import ctypes
import logging
from time import perf_counter as pc

import numpy as np
import torch

from . import _native
from .control import ilqrSolver
from .engine import CompiledBatch, default_device, solve_specs, spec_from_problem

g = 9.81


def inter_graph_batch(X, radius, n_agents, n_states, device=None):
    """Adjacency bit masks for many scenarios at once.

    X: [n_scen, rows, n_agents*n_states] (NumPy or tensor), radius: scalar or [n_scen].
    Returns a uint64 NumPy array [n_scen, n_agents]; bit j of entry (k, i) is set iff agent j is
    in agent i's neighbourhood (itself included)."""
    _native.require_device()
    dev = torch.device(device) if device is not None else default_device()
    Xt = X if isinstance(X, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64))
    Xt = Xt.to(device=dev, dtype=torch.float64).contiguous()
    n_scen, rows = Xt.shape[0], Xt.shape[1]
    rad = torch.as_tensor(np.broadcast_to(np.asarray(radius, dtype=np.float64), (n_scen,)).copy()).to(dev)
    adj = torch.zeros((n_scen, n_agents), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _native.check(_native.lib().dpilqr_inter_graph(
            ctypes.c_void_p(Xt.data_ptr()), n_scen, rows, n_agents, n_states, ctypes.c_void_p(rad.data_ptr()),
            ctypes.c_void_p(adj.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return adj.cpu().numpy().view(np.uint64)


def _masks_to_graph(masks, ids):
    return {id_: [ids[j] for j in range(len(ids)) if (int(masks[i]) >> j) & 1] for i, id_ in enumerate(ids)}


def define_inter_graph_threshold(X, radius, x_dims, ids):
    """Interaction graph by thresholding planar distances sampled along the trajectory
    (reference distributed.py:224-247): ``{id: sorted([id] + neighbours)}``."""
    assert len(set(x_dims)) == 1
    n_agents = len(x_dims)
    if n_agents == 1:
        raise ValueError("Can't compute pairwise distance for one agent.")  # reference util.py:55-56
    X = np.asarray(X, dtype=np.float64)
    X = X.reshape(-1, sum(x_dims))
    masks = inter_graph_batch(X[None], radius, n_agents, x_dims[0])[0]
    graph = _masks_to_graph(masks, list(ids))
    return {k: sorted(v) for k, v in graph.items()}


def solve_distributed_batch(problems, Xs, Us, radius, ignore_ids=None, **kwargs):
    """One DP-iLQR round for MANY scenarios at once (not in the reference, which loops).

    problems: list of ilqrProblem (or pre-compiled ProblemSpec); Xs[k]: [rows, n]; Us[k]: [N, m];
    ignore_ids: None or one list of ids per scenario.
    Returns a list of (X_dec, U_dec, J_full, solve_info) tuples, one per scenario, plus the total
    number of sub-problem iterations as the second return value."""
    from .engine import ProblemSpec

    specs = [p if isinstance(p, ProblemSpec) else spec_from_problem(p) for p in problems]
    n_scen = len(specs)
    radius = np.broadcast_to(np.asarray(radius, dtype=np.float64), (n_scen,))
    ignore = [[] for _ in range(n_scen)] if ignore_ids is None else [list(ig or []) for ig in ignore_ids]
    Xs = [np.asarray(X, dtype=np.float64).reshape(-1, sp.a * sp.s) for X, sp in zip(Xs, specs)]
    Us = [np.asarray(U, dtype=np.float64) for U in Us]
    t0 = pc()
    # ---- interaction graphs: one launch per (rows, a, s) group
    graphs = [None] * n_scen
    groups = {}
    for k, (X, sp) in enumerate(zip(Xs, specs)):
        groups.setdefault((X.shape[0], sp.a, sp.s), []).append(k)
    for (rows, a, s), idxs in groups.items():
        masks = inter_graph_batch(np.stack([Xs[k] for k in idxs]), radius[idxs], a, s)
        for j, k in enumerate(idxs):
            graphs[k] = masks[j]
    # ---- sub-problems of every scenario
    sub_specs, sub_x0, sub_U0, owner = [], [], [], []
    for k, sp in enumerate(specs):
        s, c = sp.s, sp.c
        for i, id_ in enumerate(sp.ids):
            if id_ in ignore[k]:
                continue
            members = [j for j in range(sp.a) if (int(graphs[k][i]) >> j) & 1]
            # split_graph orders columns by the *sorted id list* (reference util.py:107-115), the sub-models keep the
            # parent's order (reference dynamics.py:194); the two agree whenever ids ascend with position.
            cols = sorted(members, key=lambda j: sp.ids[j])
            sub_specs.append(sp.subset(members))
            sub_x0.append(np.concatenate([Xs[k][0, j * s:(j + 1) * s] for j in cols]))
            sub_U0.append(np.concatenate([Us[k][:, j * c:(j + 1) * c] for j in cols], axis=1))
            owner.append((k, i, sub_specs[-1].ids.index(id_)))
    N = Us[0].shape[0]
    results = solve_specs(sub_specs, sub_x0, sub_U0, N, **kwargs)
    elapsed = pc() - t0
    # ---- stitch each agent's own columns back and evaluate the joint cost
    X_dec = [np.zeros((N + 1, sp.a * sp.s)) for sp in specs]
    U_dec = [np.zeros((N, sp.a * sp.c)) for sp in specs]
    infos = [dict() for _ in specs]
    total_iters = 0
    for (k, i, pos), sub, res in zip(owner, sub_specs, results):
        s, c = specs[k].s, specs[k].c
        X_dec[k][:, i * s:(i + 1) * s] = res["X"][:, pos * s:(pos + 1) * s]
        U_dec[k][:, i * c:(i + 1) * c] = res["U"][:, pos * c:(pos + 1) * c]
        infos[k][specs[k].ids[i]] = (elapsed / max(len(owner), 1), sorted(sub.ids))
        total_iters += int(res["iters"])
    J_full = [None] * n_scen
    bins = {}
    for k, sp in enumerate(specs):
        bins.setdefault(sp.key, []).append(k)
    for key, idxs in bins.items():
        batch = CompiledBatch([specs[k] for k in idxs], N)
        _, J = batch.rollout(np.stack([Xs[k][0] for k in idxs]), np.stack([U_dec[k] for k in idxs]))
        for j, k in enumerate(idxs):
            J_full[k] = float(J[j].item())
    return [(X_dec[k], U_dec[k], J_full[k], infos[k]) for k in range(n_scen)], total_iters


def solve_distributed(problem, X, U, radius, ignore_ids=None, pool=None, verbose=True, **kwargs):
    """Solve the problem by splitting it into one sub-problem per agent over its neighbourhood
    (reference distributed.py:25-103).  ``pool`` is accepted for compatibility and ignored: the
    sub-problems run as batched GPU launches.  ``ignore_ids=None`` behaves like ``[]`` (the
    reference raises TypeError there, distributed.py:59)."""
    ids = problem.ids
    if ignore_ids and any(id_ not in ids for id_ in ignore_ids):
        raise ValueError(f"Some of {ignore_ids} not in {ids}.")
    results, _ = solve_distributed_batch([problem], [X], [U], radius, [list(ignore_ids or [])], **kwargs)
    X_dec, U_dec, J_full, info = results[0]
    if verbose:
        print("=" * 80 + f"\nInteraction Graph: { {k: v[1] for k, v in info.items()} }")
        for id_ in (ignore_ids or []):
            info[id_] = (0.0, [id_])
            print(f"Ignoring subproblem {id_}...")
    return X_dec, U_dec, J_full, info


def solve_centralized(solver, xi, U, ids, verbose, **kwargs):
    """Thin wrapper that times one centralized solve (reference distributed.py:250-258)."""
    t0 = pc()
    X, U, J = solver.solve(xi, U, verbose=verbose, **kwargs)
    Δt = pc() - t0
    return X, U, J, {id_: (Δt, ids) for id_ in ids}


def solve_rhc(problem, x0, N, *args, centralized=True, n_d=2, step_size=1, J_converge=None, dist_converge=None,
              t_diverge=None, i_trial=None, verbose=False, U0=None, sharded=False, **kwargs):
    """Receding-horizon loop, centralized or decentralized (reference distributed.py:106-221).

    ``U0`` (not in the reference) overrides the ``np.random.rand(N, n_u) * 0.01`` warm start the
    reference draws from the global NumPy RNG (:152); when it is None the same draw is made so a
    seeded run consumes the RNG identically.  ``sharded=True`` (decentralized mode, under torch.distributed): every
    rank runs this same loop, solves only its own agents' sub-problems each round, and the ranks exchange the new
    agent trajectories with one NCCL all-gather per round (dpilqr_b200.parallel)."""
    if (J_converge is None) == (dist_converge is None):
        raise ValueError("Must either specify a convergence cost or distance")
    xf = problem.game_cost.xf
    n_states = problem.dynamics.x_dims[0]
    n_agents = problem.n_agents

    def distance_to_goal(x):
        return np.linalg.norm((x - xf).reshape(n_agents, n_states)[:, :n_d], axis=1)

    if J_converge:
        def predicate(_, J):
            return J >= J_converge
    else:
        def predicate(x, _):
            return np.any(distance_to_goal(x) > dist_converge)

    n_x, n_u = problem.dynamics.n_x, problem.dynamics.n_u
    model_name = problem.dynamics.submodels[0].__class__.__name__
    xi = np.asarray(x0, dtype=np.float64).reshape(1, -1)
    X = xi.copy()
    U = np.random.rand(N, n_u) * 0.01 if U0 is None else np.asarray(U0, dtype=np.float64).copy()
    centralized_solver = ilqrSolver(problem, N)
    t = 0
    J = np.inf
    converged = True
    dt = problem.dynamics.dt
    ids = problem.ids.copy()
    X_full = np.zeros((0, n_x))
    U_full = np.zeros((0, n_u))
    times, subgraphs, distance_left = [], [], distance_to_goal(xi.flatten()).tolist()
    while predicate(xi.flatten(), J):
        if verbose:
            print(f"t: {t:.3g}")
        if centralized:
            X, U, J, solve_info = solve_centralized(centralized_solver, xi, U, ids, False, **kwargs)
        elif sharded:
            from .parallel import solve_distributed_sharded

            t0 = pc()
            X, U, graph = solve_distributed_sharded(problem, X, U, args[0], args[1] if len(args) > 1 else None, **kwargs)
            _, J = centralized_solver._rollout(xi, U)
            solve_info = {id_: (pc() - t0, members) for id_, members in graph.items()}
        else:
            X, U, J, solve_info = solve_distributed(problem, X, U, *args, verbose=False, **kwargs)
        xi = X[step_size]
        X_full = np.r_[X_full, X[:step_size]]
        U_full = np.r_[U_full, U[:step_size]]
        # warm start of the next round: shift and hold the last state (reference :184-185)
        X = np.r_[X[step_size:], np.tile(X[-1], (step_size, 1))]
        U = np.r_[U[step_size:], np.zeros((step_size, n_u))]
        times = [tup[0] for tup in solve_info.values()]
        subgraphs = [tup[1] for tup in solve_info.values()]
        distance_left = distance_to_goal(xi).tolist()
        logging.info(
            f'"{model_name}",{problem.n_agents},{i_trial},{centralized},{False},{t},{J},{N},{dt},{converged},"{ids}",'
            f'"{times}","{subgraphs}","{distance_left}"'
        )
        if t_diverge and t >= t_diverge:
            converged = False
            if verbose:
                print("Failed to converge within allotted time...")
            break
        t += step_size * dt
    if not X_full.size and not U_full.size:
        X_full = np.asarray(x0, dtype=np.float64).copy()
        U_full = np.zeros((1, n_u))
    _, J_full = centralized_solver._rollout(x0, U_full)
    tf = U_full.shape[0] * dt
    logging.info(
        f'"{model_name}",{problem.n_agents},{i_trial},{centralized},{True},{tf},{J_full},{N},{dt},{converged},"{ids}",'
        f'"{times}","{subgraphs}","{distance_left}"'
    )
    return X_full, U_full, J_full
