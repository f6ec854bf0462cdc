"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink).

The hot path shards without any data-path collective: scenarios (Potential-iLQR) and
(scenario, agent) sub-problems (DP-iLQR) are independent.  The one exchange the algorithm has is
between receding-horizon rounds when the agents of ONE scenario are spread over ranks: every rank
must see all agents' new trajectories before the next round's interaction graph (reference
distributed.py:174-185).  That is a single all-gather of (T+1)*s + T*c doubles per agent.
"""

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of `n_items` independent units owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owned_agents(n_agents, rank, world):
    """Positions of the agents whose sub-problems `rank` solves in a sharded DP-iLQR round."""
    lo, hi = shard_bounds(n_agents, rank, world)
    return list(range(lo, hi))


def allgather_agent_trajectories(X_own, U_own, positions, n_agents, group=None, device=None):
    """Exchange agent trajectories between receding-horizon rounds.

    X_own: [k, T+1, s], U_own: [k, T, c] for the k agents at `positions` owned by this rank.
    Returns X_dec [T+1, n_agents*s], U_dec [T, n_agents*c] assembled from all ranks (NumPy).
    One all-gather per round: the payload is padded to the largest shard so a single collective suffices."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    X_own = np.asarray(X_own, dtype=np.float64)
    U_own = np.asarray(U_own, dtype=np.float64)
    T1, s = (X_own.shape[1], X_own.shape[2]) if X_own.size else (None, None)
    if world == 1:
        if len(positions) != n_agents:
            raise ValueError("a single rank must own every agent")
        order = np.argsort(positions)
        return (np.concatenate([X_own[i] for i in order], axis=1), np.concatenate([U_own[i] for i in order], axis=1))
    kmax = -(-n_agents // world)
    meta = torch.tensor([X_own.shape[1] if X_own.size else 0, X_own.shape[2] if X_own.size else 0,
                         U_own.shape[1] if U_own.size else 0, U_own.shape[2] if U_own.size else 0], dtype=torch.int64)
    backend = dist.get_backend(group)
    dev = torch.device(device) if device is not None else (torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))
    meta = meta.to(dev)
    dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=group)  # ranks that own no agent learn the shapes
    T1, s, T, c = (int(v) for v in meta.tolist())
    width = T1 * s + T * c
    payload = torch.zeros((kmax, width + 1), dtype=torch.float64, device=dev)
    payload[:, width] = -1.0  # position tag; -1 marks padding
    for j, pos in enumerate(positions):
        payload[j, : T1 * s] = torch.as_tensor(X_own[j].reshape(-1))
        payload[j, T1 * s: width] = torch.as_tensor(U_own[j].reshape(-1))
        payload[j, width] = float(pos)
    gathered = torch.empty((world * kmax, width + 1), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gathered, payload, group=group)
    gathered = gathered.cpu().numpy()
    X_dec = np.zeros((T1, n_agents * s))
    U_dec = np.zeros((T, n_agents * c))
    seen = set()
    for row in gathered:
        pos = int(row[width])
        if pos < 0:
            continue
        seen.add(pos)
        X_dec[:, pos * s:(pos + 1) * s] = row[: T1 * s].reshape(T1, s)
        U_dec[:, pos * c:(pos + 1) * c] = row[T1 * s: width].reshape(T, c)
    return X_dec, U_dec


def solve_distributed_sharded(problem, X, U, radius, ignore_ids=None, group=None, solve_fn=None, graph_fn=None, **kwargs):
    """One DP-iLQR round (reference distributed.py:25-103) with the agents' sub-problems sharded over the ranks
    of `group`: every rank builds the (bit-exact, hence identical) interaction graph, solves the sub-problems of
    its own agents on its GPU, and one all-gather stitches X_dec / U_dec back together on every rank.

    `solve_fn(specs, x0s, U0s, N, **kwargs)` and `graph_fn(X, radius, x_dims, ids)` default to the CUDA
    implementations; they are injection points for the CPU (gloo) tests of this orchestration."""
    from .distributed import define_inter_graph_threshold
    from .engine import solve_specs, spec_from_problem

    solve_fn = solve_fn or solve_specs
    graph_fn = graph_fn or define_inter_graph_threshold
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    spec = spec_from_problem(problem) if not hasattr(problem, "subset") else problem
    ignore_ids = list(ignore_ids or [])
    X = np.asarray(X, dtype=np.float64).reshape(-1, spec.a * spec.s)
    U = np.asarray(U, dtype=np.float64)
    N = U.shape[0]
    s, c = spec.s, spec.c
    graph = graph_fn(X, radius, [s] * spec.a, spec.ids)
    mine = [i for i in owned_agents(spec.a, rank, world) if spec.ids[i] not in ignore_ids]
    specs, x0s, U0s, pos_in_sub = [], [], [], []
    for i in mine:
        members = [spec.ids.index(int(id_)) for id_ in graph[spec.ids[i]]]
        cols = sorted(members, key=lambda j: spec.ids[j])
        sub = spec.subset(sorted(members))
        specs.append(sub)
        x0s.append(np.concatenate([X[0, j * s:(j + 1) * s] for j in cols]))
        U0s.append(np.concatenate([U[:, j * c:(j + 1) * c] for j in cols], axis=1))
        pos_in_sub.append(sub.ids.index(spec.ids[i]))
    results = solve_fn(specs, x0s, U0s, N, **kwargs) if specs else []
    X_own = np.stack([res["X"][:, k * s:(k + 1) * s] for res, k in zip(results, pos_in_sub)]) if results else np.zeros((0, N + 1, s))
    U_own = np.stack([res["U"][:, k * c:(k + 1) * c] for res, k in zip(results, pos_in_sub)]) if results else np.zeros((0, N, c))
    if world == 1 and len(mine) == spec.a:
        X_dec, U_dec = allgather_agent_trajectories(X_own, U_own, mine, spec.a, group)
    elif world == 1:
        X_dec, U_dec = np.zeros((N + 1, spec.a * s)), np.zeros((N, spec.a * c))
        for j, pos in enumerate(mine):
            X_dec[:, pos * s:(pos + 1) * s] = X_own[j]
            U_dec[:, pos * c:(pos + 1) * c] = U_own[j]
    else:
        X_dec, U_dec = allgather_agent_trajectories(X_own, U_own, mine, spec.a, group)
    return X_dec, U_dec, graph
