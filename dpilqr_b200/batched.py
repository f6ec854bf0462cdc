"""Device-resident batched DP-iLQR and receding-horizon drivers (not in the reference, which loops over
scenarios, rounds and agents in Python: reference distributed.py:25-221, scripts/analysis.py:35-174).

Thousands of scenarios that share one team shape (agents, model sizes, horizon, dt) are advanced together:

* :func:`solve_distributed_round` -- one DP-iLQR round for every scenario of a :class:`CompiledBatch`: the
  interaction-graph kernel, then the sub-problems of ALL scenarios binned by neighbourhood size and built with
  gathers on the device (no per-agent Python objects), one batched solve per bin, the agents' own columns
  scattered back, and the joint cost of the stitched controls (reference distributed.py:25-103).
* :func:`solve_rhc_batch` -- the receding-horizon loop of reference distributed.py:106-221 for every scenario at
  once: trajectories, warm starts (shift :184-185), the convergence predicate (:130-143) and the graphs stay on the
  device; per round the host reads one flag per scenario plus, if asked, the log rows of the reference's CSV
  schema (:190-194) in one transfer.
* :func:`trajectory_metrics` -- minimum pairwise separation and collision counts of whole trajectories (the data
  behind reference graphics.plot_pairwise_distances, graphics.py:146-156) as a batched reduction.

Sub-problem columns follow agent position; the reference orders them by sorted id (util.py:107-115), the two
agree when ids ascend with position, which :func:`solve_distributed_round` requires of its callers.
"""

import ctypes

import numpy as np
import torch

from . import _native
from .engine import CompiledBatch, raise_for_status


def _graph_masks(batch, X, radius):
    """Adjacency bit masks [B, a] (int64, device) of `define_inter_graph_threshold` for every scenario."""
    B, rows = X.shape[0], X.shape[1]
    adj = torch.zeros((B, batch.a), dtype=torch.int64, device=batch.device)
    rad = radius if isinstance(radius, torch.Tensor) else torch.full((B,), float(radius), dtype=torch.float64, device=batch.device)
    rad = rad.to(device=batch.device, dtype=torch.float64).contiguous()
    with torch.cuda.device(batch.device):
        _native.check(_native.lib().dpilqr_inter_graph(
            ctypes.c_void_p(X.data_ptr()), B, rows, batch.a, batch.s, ctypes.c_void_p(rad.data_ptr()),
            ctypes.c_void_p(adj.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(batch.device).cuda_stream)))
    return adj


def solve_distributed_round(batch, X, U, radius, ignore=None, on_error="raise", **solve_kw):
    """One DP-iLQR round (reference distributed.py:25-103) for all B scenarios of ``batch``.

    X [B, rows, n] (rows = 1 or N+1), U [B, N, m]: device tensors (or anything ``torch.as_tensor`` takes);
    radius: scalar or [B]; ignore: optional bool [B, a], True = the reference's ``ignore_ids``.
    Returns a dict of device tensors: X_dec [B, N+1, n], U_dec [B, N, m], J_full [B], adjacency [B, a] (bit masks),
    sub_iters [B, a] (iterations of every agent's sub-problem), status [B, a], plus ``total_iters`` and ``bins``
    ({neighbourhood size: number of sub-problems})."""
    dev, a, s, c, N = batch.device, batch.a, batch.s, batch.c, batch.N
    B = batch.B
    f64 = dict(dtype=torch.float64, device=dev)
    X = torch.as_tensor(X).to(**f64).reshape(B, -1, batch.n).contiguous()
    U = torch.as_tensor(U).to(**f64).reshape(B, N, batch.m).contiguous()
    if a == 1:
        raise ValueError("Can't compute pairwise distance for one agent.")  # reference util.py:55-56
    adj = _graph_masks(batch, X, radius)
    member = ((adj.unsqueeze(-1) >> torch.arange(a, device=dev)) & 1).bool()  # [B, a, a]: j in neighbourhood of i
    sizes = member.sum(-1)  # [B, a]
    if ignore is not None:
        sizes = torch.where(torch.as_tensor(ignore, device=dev).bool(), torch.zeros_like(sizes), sizes)
    X_dec = torch.zeros((B, N + 1, a, s), **f64)
    U_dec = torch.zeros((B, N, a, c), **f64)
    sub_iters = torch.zeros((B, a), dtype=torch.int32, device=dev)
    status = torch.zeros((B, a), dtype=torch.int32, device=dev)
    x0 = X[:, 0].reshape(B, a, s)
    U4 = U.reshape(B, N, a, c)
    xf3 = batch.t_xf.reshape(B, a, s)
    total, bins = 0, {}
    for k in torch.unique(sizes).tolist():  # host sync: the handful of neighbourhood sizes present
        if k == 0:
            continue
        b_idx, i_idx = torch.nonzero(sizes == k, as_tuple=True)
        n_k = int(b_idx.numel())
        bins[int(k)] = n_k
        rows_k = member[b_idx, i_idx]  # [n_k, a]
        # ascending positions of the members: a stable sort of the complement keeps index order
        mem = torch.argsort((~rows_k).to(torch.int8), dim=1, stable=True)[:, :k]  # [n_k, k]
        ar = torch.arange(n_k, device=dev)
        bb = b_idx[:, None]
        sub = CompiledBatch.from_tensors(
            N, k, s, c, batch.dt, batch.t_model[bb, mem], batch.t_ndims[bb, mem], batch.t_cidx[bb, mem], batch.t_Q, batch.t_R,
            batch.t_Qf, xf3[bb, mem].reshape(n_k, k * s), batch.t_radius[b_idx],
            torch.tensor([1.0, 200.0], **f64).expand(n_k, 2),  # a fresh GameCost carries the default weights (cost.py:185-186)
            torch.ones(n_k, dtype=torch.int32, device=dev), batch.model_hint, batch.costs_nonnegative, dev)
        x0_k = x0[bb, mem].reshape(n_k, k * s)
        U0_k = torch.gather(U4[b_idx], 2, mem[:, None, :, None].expand(n_k, N, k, c)).reshape(n_k, N, k * c)
        out = sub.solve(x0_k, U0_k, **solve_kw)
        total += out["total_iters"]
        pos = (mem < i_idx[:, None]).sum(1)  # where agent i sits among its neighbourhood (problem.extract)
        X_dec[b_idx, :, i_idx] = out["X"].reshape(n_k, N + 1, k, s)[ar, :, pos]
        U_dec[b_idx, :, i_idx] = out["U"].reshape(n_k, N, k, c)[ar, :, pos]
        sub_iters[b_idx, i_idx] = out["iters"]
        status[b_idx, i_idx] = out["status"]
    if on_error == "raise":
        bad = status & (_native.ST_POINT_NDIM | _native.ST_SINGULAR)
        if bool(bad.any()):
            raise_for_status(int(bad[bad != 0][0]))
    X_dec = X_dec.reshape(B, N + 1, batch.n)
    U_dec = U_dec.reshape(B, N, batch.m)
    _, J_full = batch.rollout(X[:, 0].contiguous(), U_dec)
    return dict(X_dec=X_dec, U_dec=U_dec, J_full=J_full, adjacency=adj, sub_iters=sub_iters, status=status,
                total_iters=int(total), bins=bins)


def masks_to_graphs(adj, ids):
    """Bit masks [a] -> the reference's ``{id: sorted neighbour ids}`` dictionary."""
    return {id_: [ids[j] for j in range(len(ids)) if (int(adj[i]) >> j) & 1] for i, id_ in enumerate(ids)}


LOG_HEADER = "dynamics,n_agents,trial,centralized,last,t,J,horizon,dt,converged,ids,times,subgraphs,dist_left"


def solve_rhc_batch(batch, x0, radius=None, centralized=True, n_d=2, step_size=1, dist_converge=None, t_diverge=None,
                    U0=None, max_rounds=None, log=None, model_name="", ids=None, trial0=0, seed=None, **solve_kw):
    """Receding-horizon simulation of ALL scenarios of ``batch`` at once (reference distributed.py:106-221,
    dist_converge mode).

    x0 [B, n]; U0 [B, N, m] or None (then ``np.random.rand(N, m) * 0.01`` is drawn per scenario from
    ``np.random.RandomState(seed + k)``); ``radius`` is the interaction-graph radius of the decentralised mode.
    Every scenario runs until all its agents are within ``dist_converge`` of their goals or ``t >= t_diverge``
    (checked after the round, like the reference); scenarios that are done leave the batch.
    ``log``: a file object (or list) that receives one row per scenario and round in the reference's CSV schema
    (LOG_HEADER; wall-clock ``times`` are the round's device time split evenly).
    Returns a dict: X_full / U_full (lists of [steps_k, n] / [steps_k, m] device tensors), J_full [B],
    rounds [B] (int), converged [B] (bool), total_iters."""
    if dist_converge is None:
        raise ValueError("Must either specify a convergence cost or distance")
    dev, a, s, c, N = batch.device, batch.a, batch.s, batch.c, batch.N
    B, n, m = batch.B, batch.n, batch.m
    f64 = dict(dtype=torch.float64, device=dev)
    dt = batch.dt
    if max_rounds is None:
        if not t_diverge:
            raise ValueError("solve_rhc_batch needs t_diverge or max_rounds to bound the simulation")
        max_rounds = int(np.ceil(t_diverge / (step_size * dt))) + 2
    x0 = torch.as_tensor(x0).to(**f64).reshape(B, n)
    if U0 is None:
        U_np = np.stack([np.random.RandomState(None if seed is None else seed + k).rand(N, m) * 0.01 for k in range(B)])
        U = torch.as_tensor(U_np).to(**f64)
    else:
        U = torch.as_tensor(U0).to(**f64).reshape(B, N, m).clone()
    xf3 = batch.t_xf.reshape(B, a, s)

    def dist_left(x, idx):  # [k, n] -> [k, a]: distance_to_goal of reference distributed.py:130-131
        return torch.linalg.vector_norm((x.reshape(-1, a, s) - xf3[idx])[:, :, :n_d], dim=2)

    active0 = (dist_left(x0, slice(None)) > dist_converge).any(1)
    X_full = torch.zeros((B, max_rounds * step_size, n), **f64)
    U_full = torch.zeros((B, max_rounds * step_size, m), **f64)
    rounds = torch.zeros(B, dtype=torch.int64, device=dev)
    converged = torch.ones(B, dtype=torch.bool, device=dev)
    alive = torch.nonzero(active0, as_tuple=True)[0]
    xi = x0[alive]
    X = xi[:, None, :].contiguous()
    U = U[alive]
    t, rnd, total = 0.0, 0, 0
    id_list = list(ids) if ids is not None else [100 + i for i in range(a)]
    while alive.numel() > 0 and rnd < max_rounds:
        sub = batch.select(alive)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if centralized:
            out = sub.solve(xi, U, **solve_kw)
            for st in out["status"].tolist() if bool((out["status"] & 6).any()) else []:
                raise_for_status(st)
            Xn, Un, J, adj = out["X"], out["U"], out["J"], None
            total += out["total_iters"]
        else:
            rad = radius[alive] if isinstance(radius, torch.Tensor) else radius
            out = solve_distributed_round(sub, X, U, rad, **solve_kw)
            Xn, Un, J, adj = out["X_dec"], out["U_dec"], out["J_full"], out["adjacency"]
            total += out["total_iters"]
        e1.record()
        xi = Xn[:, step_size].contiguous()
        X_full[alive, rnd * step_size:(rnd + 1) * step_size] = Xn[:, :step_size]
        U_full[alive, rnd * step_size:(rnd + 1) * step_size] = Un[:, :step_size]
        rounds[alive] += 1
        # warm start of the next round: shift and hold the last state (reference :184-185)
        X = torch.cat([Xn[:, step_size:], Xn[:, -1:].expand(-1, step_size, -1)], dim=1).contiguous()
        U = torch.cat([Un[:, step_size:], torch.zeros((Un.shape[0], step_size, m), **f64)], dim=1).contiguous()
        left = dist_left(xi, alive)
        keep = (left > dist_converge).any(1)
        timed_out = bool(t_diverge) and t >= t_diverge
        if log is not None:  # one transfer per round: the reference's log rows (:190-194)
            e1.synchronize()
            per = e0.elapsed_time(e1) * 1e-3 / max(int(alive.numel()), 1)
            host = torch.cat([J[:, None], left], dim=1).cpu().numpy()
            adj_h = adj.cpu().numpy() if adj is not None else None
            for row, k in enumerate(alive.tolist()):
                graphs = [id_list] * a if adj_h is None else list(masks_to_graphs(adj_h[row], id_list).values())
                line = (f'"{model_name}",{a},{trial0 + k},{centralized},{False},{t},{host[row, 0]},{N},{dt},{True},"{id_list}",'
                        f'"{[per] * a}","{graphs}","{host[row, 1:].tolist()}"')
                log.append(line) if isinstance(log, list) else log.write(line + "\\n")
        if timed_out:  # the reference flags the run before it looks at the predicate again (:196-200)
            converged[alive] = False
            keep = torch.zeros_like(keep)
        sel = torch.nonzero(keep, as_tuple=True)[0]
        alive, xi, X, U = alive[sel], xi[sel], X[sel], U[sel]
        t += step_size * dt
        rnd += 1
    if alive.numel() > 0:
        converged[alive] = False
    # joint cost of what was actually flown: group the scenarios by the number of steps they took
    J_full = torch.zeros(B, **f64)
    steps = rounds * step_size
    for L in torch.unique(steps).tolist():
        idx = torch.nonzero(steps == L, as_tuple=True)[0]
        if L == 0:  # immediate convergence: the reference rolls out a single zero control (:203-205)
            _, Jk = batch.select(idx, 1).rollout(x0[idx], torch.zeros((idx.numel(), 1, m), **f64))
        else:
            _, Jk = batch.select(idx, int(L)).rollout(x0[idx], U_full[idx, :int(L)].contiguous())
        J_full[idx] = Jk
    steps_h = steps.tolist()
    return dict(X_full=[X_full[k, :steps_h[k]] for k in range(B)], U_full=[U_full[k, :steps_h[k]] for k in range(B)],
                J_full=J_full, rounds=rounds, converged=converged, total_iters=int(total))


def trajectory_metrics(X, a, s, radius, n_d=2):
    """Collision statistics of whole trajectories X [B, rows, a*s] (device): minimum pairwise separation over time
    per scenario and the number of (time step, pair) samples closer than ``radius``
    (reference util.compute_pairwise_distance over the trajectory, graphics.py:146-156)."""
    X = torch.as_tensor(X)
    B, rows = X.shape[0], X.shape[1]
    P = X.reshape(B, rows, a, s)[..., :n_d]
    i, j = torch.triu_indices(a, a, offset=1, device=X.device)
    d = torch.linalg.vector_norm(P[:, :, i] - P[:, :, j], dim=-1)  # [B, rows, pairs]
    return dict(min_separation=d.amin(dim=(1, 2)), violations=(d < radius).sum(dim=(1, 2)), pairwise=d)
