"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the batched iLQR hot path.

NumPy restatement of the reference's algorithm (Potential-iLQR / DP-iLQR),
function by function, with the reference file:line each one follows.  It is
the checker for ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; it is never
imported by the product package ``dpilqr_b200`` (which has no CPU path at all).

Parity status: PINNED.  ``tests/test_oracle.py`` checks this file against
(a) golden vectors produced by running the unmodified reference in the build
container (``tests/golden/*.npz``, generator ``tests/golden/generate_golden.py``)
and (b), when ``/root/reference`` is present, against the live reference.

The computational *structure* deliberately mirrors the reference (Python loop
over time steps, dense block-diagonal Jacobians, per-pair dense scatter in the
proximity quadraticisation, two LU solves per step) so that timing it is a fair
stand-in for timing the reference's CPU path.

Single-agent dynamics come from one of two native back ends:
  * ``oracle/_ref``  -- the reference's own Cython module compiled from its own
    sources by ``oracle/build_ref.sh`` (preferred when present: exact reference
    arithmetic and call overhead);
  * ``oracle/dynamics_oracle.c`` -- the plain-C restatement via ctypes.
"""

import ctypes
import itertools
import os
import subprocess
from time import perf_counter

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# Reference enum order, dpilqr/bbdynamicswrap.pyx:8-16 (+ Bike5D appended).
MODEL_IDS = {
    "DoubleInt4D": 0,
    "DoubleInt6D": 1,
    "Car3D": 2,
    "Unicycle4D": 3,
    "Quadcopter6D": 4,
    "Human6D": 5,
    "HumanLin6D": 6,
    "Quadcopter12D": 7,
    "Bike5D": 8,
}
MODEL_DIMS = {0: (4, 2), 1: (6, 3), 2: (3, 2), 3: (4, 2), 4: (6, 3), 5: (6, 3), 6: (6, 3), 7: (12, 4), 8: (5, 2)}


# --------------------------------------------------------------------------- #
# native back ends
# --------------------------------------------------------------------------- #
def build_c_oracle(force=False):
    """gcc the C restatement into oracle/_build/liboracle_dynamics.so."""
    out_dir = os.path.join(_HERE, "_build")
    out = os.path.join(out_dir, "liboracle_dynamics.so")
    src = os.path.join(_HERE, "dynamics_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", out, "-lm"]
        )
    return out


class _CDynamics:
    """ctypes view of oracle/dynamics_oracle.c."""

    name = "c-restatement"

    def __init__(self):
        lib = ctypes.CDLL(build_c_oracle())
        dp = ctypes.POINTER(ctypes.c_double)
        lib.oracle_f.argtypes = [ctypes.c_int, dp, dp, dp]
        lib.oracle_integrate.argtypes = [ctypes.c_int, ctypes.c_double, dp, dp, dp]
        lib.oracle_linearize.argtypes = [ctypes.c_int, ctypes.c_double, dp, dp, dp, dp]
        self.lib = lib
        self._dp = dp

    def _p(self, a):
        return a.ctypes.data_as(self._dp)

    def f(self, x, u, model):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty(MODEL_DIMS[model][0])
        self.lib.oracle_f(model, self._p(x), self._p(u), self._p(out))
        return out

    def integrate(self, x, u, dt, model):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty(MODEL_DIMS[model][0])
        self.lib.oracle_integrate(model, dt, self._p(x), self._p(u), self._p(out))
        return out

    def linearize(self, x, u, dt, model):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        nx, nu = MODEL_DIMS[model]
        A = np.empty((nx, nx))
        B = np.empty((nx, nu))
        self.lib.oracle_linearize(model, dt, self._p(x), self._p(u), self._p(A), self._p(B))
        return A, B


class _RefDynamics:
    """The reference's own compiled Cython module (oracle/_ref); Bike5D falls
    through to the C restatement because the reference implements it in sympy."""

    name = "reference-native"

    def __init__(self, module):
        self.m = module
        self.enum = {v.value: v for v in module.Model}
        self.c = _CDynamics()

    def f(self, x, u, model):
        if model == 8:
            return self.c.f(x, u, model)
        return self.m.f(np.ascontiguousarray(x), np.ascontiguousarray(u), self.enum[model])

    def integrate(self, x, u, dt, model):
        if model == 8:
            return self.c.integrate(x, u, dt, model)
        return self.m.integrate(x, u, dt, self.enum[model])

    def linearize(self, x, u, dt, model):
        if model == 8:
            return self.c.linearize(x, u, dt, model)
        return self.m.linearize(x, u, dt, self.enum[model])


_BACKENDS = {}


def dynamics_backend(kind="auto"):
    """kind: 'c', 'ref' or 'auto' (ref if oracle/_ref holds the module)."""
    if kind in _BACKENDS:
        return _BACKENDS[kind]
    be = None
    if kind in ("ref", "auto"):
        ref_dir = os.path.join(_HERE, "_ref")
        cands = [f for f in (os.listdir(ref_dir) if os.path.isdir(ref_dir) else []) if f.startswith("bbdynamicswrap") and f.endswith(".so")]
        if cands:
            import importlib.util

            spec = importlib.util.spec_from_file_location("bbdynamicswrap", os.path.join(ref_dir, cands[0]))
            try:
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                be = _RefDynamics(mod)
            except Exception:
                if kind == "ref":
                    raise
        elif kind == "ref":
            raise RuntimeError("oracle/_ref not built")
    if be is None:
        be = _CDynamics()
    _BACKENDS[kind] = be
    return be


# --------------------------------------------------------------------------- #
# problem description
# --------------------------------------------------------------------------- #
class OracleProblem:
    """Flat description of one (sub)problem: what ilqrProblem + MultiDynamicalModel
    + GameCost hold in the reference (problem.py:15-24, dynamics.py:133-146,
    cost.py:174-191)."""

    def __init__(self, models, dt, xf, Q, R, Qf, radius=0.0, n_dims=None, ids=None,
                 ref_weight=1.0, prox_weight=200.0, game=True, backend="auto"):
        self.models = [MODEL_IDS[m] if isinstance(m, str) else int(m) for m in models]
        self.a = len(self.models)
        self.s, self.c = MODEL_DIMS[self.models[0]]  # uniform strides from agent 0 (dynamics.py:165-166)
        self.dt = float(dt)
        self.n_x = self.a * self.s
        self.n_u = self.a * self.c
        self.xf = np.asarray(xf, dtype=float).flatten()

        def per_agent(M):
            if isinstance(M, np.ndarray) and M.ndim == 2:
                return [M] * self.a
            return [np.asarray(m, dtype=float) for m in M]

        self.Q = per_agent(Q)
        self.R = per_agent(R)
        self.Qf = per_agent(Qf)
        self.radius = float(radius)
        self.n_dims = list(n_dims) if n_dims is not None else [2] * self.a
        self.ids = list(ids) if ids is not None else list(range(self.a))
        self.ref_weight = ref_weight
        self.prox_weight = prox_weight
        self.game = game  # False: bare ReferenceCost + single model (examples.py:26-70)
        self.backend = backend
        self.dyn = dynamics_backend(backend)

    # ---- dynamics: MultiDynamicalModel.__call__/linearize (dynamics.py:159-186)
    def step(self, x, u):
        xn = np.zeros_like(x)
        s, c = self.s, self.c
        for i, model in enumerate(self.models):
            xn[i * s:(i + 1) * s] = self.dyn.integrate(x[i * s:(i + 1) * s], u[i * c:(i + 1) * c], self.dt, model)
        return xn

    def linearize(self, x, u):
        s, c = self.s, self.c
        subs = [
            self.dyn.linearize(x[i * s:(i + 1) * s].flatten(), u[i * c:(i + 1) * c].flatten(), self.dt, model)
            for i, model in enumerate(self.models)
        ]
        return block_diag_uniform([ab[0] for ab in subs]), block_diag_uniform([ab[1] for ab in subs])

    # ---- costs
    def ref_cost(self, i, x, u, terminal=False):
        """ReferenceCost.__call__ (cost.py:79-83)."""
        e = x - self.xf[i * self.s:(i + 1) * self.s]
        if not terminal:
            u = u.reshape(1, -1)
            return e @ self.Q[i] @ e.T + u @ self.R[i] @ u.T
        return e @ self.Qf[i] @ e.T

    def ref_quadraticize(self, i, x, u, terminal=False):
        """ReferenceCost.quadraticize (cost.py:85-101)."""
        x = x.flatten()
        u = u.flatten()
        e = x - self.xf[i * self.s:(i + 1) * self.s]
        nu, nx = self.R[i].shape[0], self.Q[i].shape[0]
        if terminal:
            QQ = self.Qf[i] + self.Qf[i].T
            return e.T @ QQ, np.zeros(nu), QQ, np.zeros((nu, nu)), np.zeros((nu, nx))
        QQ = self.Q[i] + self.Q[i].T
        RR = self.R[i] + self.R[i].T
        return e.T @ QQ, u.T @ RR, QQ, RR, np.zeros((nu, nx))

    def prox_cost(self, x):
        """ProximityCost.__call__ (cost.py:117-133): planar distance when all
        n_dims agree, min(n_dims) per pair otherwise."""
        if self.a == 1:
            return 0.0
        if len(set(self.n_dims)) == 1:
            dist = pairwise_distance(x, [self.s] * self.a)
        else:
            dist = pairwise_distance_nd(x.reshape(1, -1), [self.s] * self.a, self.n_dims)
        return (np.fmin(np.zeros(1), dist - self.radius) ** 2).sum()

    def prox_quadraticize(self, x):
        """ProximityCost.quadraticize (cost.py:135-171): per-pair dense scatter."""
        nx = self.n_x
        s = self.s
        L_x = np.zeros(nx)
        L_xx = np.zeros((nx, nx))
        for i in range(self.a):
            for j in range(i + 1, self.a):
                nd = min(self.n_dims[i], self.n_dims[j])
                gi = np.zeros(nx)
                Hi = np.zeros((nx, nx))
                ix, jx = s * i, s * j
                g_pair, H_pair = quadraticize_distance(x[ix:ix + nd], x[jx:jx + nd], self.radius, nd)
                gi[ix:ix + nd] = g_pair
                gi[jx:jx + nd] = -g_pair
                Hi[ix:ix + nd, ix:ix + nd] = H_pair
                Hi[jx:jx + nd, jx:jx + nd] = H_pair
                Hi[ix:ix + nd, jx:jx + nd] = -H_pair
                Hi[jx:jx + nd, ix:ix + nd] = -H_pair
                L_x += gi
                L_xx += Hi
        return L_x, L_xx

    def cost(self, x, u, terminal=False):
        """GameCost.__call__ (cost.py:197-206); bare ReferenceCost if not game."""
        if not self.game:
            return self.ref_cost(0, x, u, terminal)
        ref_total = 0.0
        for i in range(self.a):
            ref_total += self.ref_cost(i, x[i * self.s:(i + 1) * self.s], u[i * self.c:(i + 1) * self.c], terminal)
        return self.prox_weight * self.prox_cost(x) + self.ref_weight * ref_total

    def quadraticize(self, x, u, terminal=False):
        """GameCost.quadraticize (cost.py:208-239)."""
        if not self.game:
            return self.ref_quadraticize(0, x, u, terminal)
        parts = [
            self.ref_quadraticize(i, x[i * self.s:(i + 1) * self.s].flatten(), u[i * self.c:(i + 1) * self.c].flatten(), terminal)
            for i in range(self.a)
        ]
        L_x = self.ref_weight * np.hstack([p[0] for p in parts])
        L_u = self.ref_weight * np.hstack([p[1] for p in parts])
        L_xx = self.ref_weight * block_diag_uniform([p[2] for p in parts])
        L_uu = self.ref_weight * block_diag_uniform([p[3] for p in parts])
        L_ux = self.ref_weight * block_diag_uniform([p[4] for p in parts])
        if self.a > 1:
            g, H = self.prox_quadraticize(x)
            L_x += self.prox_weight * g
            L_xx += self.prox_weight * H
        return L_x, L_u, L_xx, L_uu, L_ux

    # ---- DP-iLQR split (problem.py:36-47, dynamics.py:188-198, cost.py:241-262)
    def subproblem(self, member_ids):
        keep = [i for i, id_ in enumerate(self.ids) if id_ in member_ids]
        return OracleProblem(
            [self.models[i] for i in keep], self.dt,
            np.concatenate([self.xf[i * self.s:(i + 1) * self.s] for i in keep]),
            [self.Q[i] for i in keep], [self.R[i] for i in keep], [self.Qf[i] for i in keep],
            self.radius, [self.n_dims[i] for i in keep], [self.ids[i] for i in keep],
            # a fresh GameCost carries the default weights (cost.py:185-186, 261)
            1.0, 200.0, True, self.backend,
        )


# --------------------------------------------------------------------------- #
# helpers (util.py)
# --------------------------------------------------------------------------- #
def block_diag_uniform(arrs):
    """uniform_block_diag (util.py:229-236)."""
    r, c = arrs[0].shape
    out = np.zeros((len(arrs) * r, len(arrs) * c))
    for i, arr in enumerate(arrs):
        out[r * i:r * (i + 1), c * i:c * (i + 1)] = arr
    return out


def pairwise_distance(X, x_dims, n_d=2):
    """compute_pairwise_distance (util.py:48-61)."""
    n_agents, n_states = len(x_dims), x_dims[0]
    if n_agents == 1:
        raise ValueError("Can't compute pairwise distance for one agent.")
    pairs = np.array(list(itertools.combinations(range(n_agents), 2)))
    Xa = X.reshape(-1, n_agents, n_states).swapaxes(0, 2)
    dX = Xa[:n_d, pairs[:, 0]] - Xa[:n_d, pairs[:, 1]]
    return np.linalg.norm(dX, axis=0).T


def pairwise_distance_nd(X, x_dims, n_dims):
    """compute_pairwise_distance_nd (util.py:64-87)."""
    if X.ndim == 1:
        X = X.reshape(1, -1)
    n_states, n_agents = x_dims[0], len(x_dims)
    cols = []
    for i, j in itertools.combinations(range(n_agents), 2):
        nd = min(n_dims[i], n_dims[j])
        cols.append(np.linalg.norm(X[:, i * n_states:i * n_states + nd] - X[:, j * n_states:j * n_states + nd], axis=1))
    return np.stack(cols, axis=1) if cols else np.zeros((X.shape[0], 0))


def quadraticize_distance(pa, pb, radius, nd):
    """quadraticize_distance (cost.py:269-315) incl. the Point.ndim assertion
    (util.py:28-30) and the cancelling distance formula in the cross terms."""
    ax, ay = pa[0], pa[1]
    bx, by = pb[0], pb[1]
    az = pa[2] if nd > 2 else 0
    bz = pb[2] if nd > 2 else 0
    assert (2 if az == 0 else 3) == (2 if bz == 0 else 3)
    g = np.zeros(3)
    H = np.zeros((3, 3))
    dx, dy, dz = ax - bx, ay - by, az - bz
    dist = np.sqrt(dx * dx + dy * dy + dz * dz)
    if dist > radius:
        return g[:nd], H[:nd, :nd]
    g = 2 * (dist - radius) / dist * np.array([dx, dy, dz])
    cross = 2 * radius / np.sqrt(
        ((ax ** 2 + ay ** 2 + az ** 2) + (bx ** 2 + by ** 2 + bz ** 2)) - 2 * (ax * bx + ay * by + az * bz)
    ) ** 3
    H[np.diag_indices(3)] = 2 * radius * np.array([dx, dy, dz]) ** 2 / dist ** 3 - 2 * radius / dist + 2
    H[np.tril_indices(3, -1)] = H[np.triu_indices(3, 1)] = np.array([dx * dy, dx * dz, dy * dz]) * cross
    return g[:nd], H[:nd, :nd]


def split_graph(Z, z_dims, graph):
    """split_graph (util.py:102-117)."""
    pos = {id_: i for i, id_ in enumerate(list(graph))}
    nz = z_dims[0]
    return [np.concatenate([Z[:, pos[id_] * nz:(pos[id_] + 1) * nz] for id_ in ids], axis=1) for ids in graph.values()]


# --------------------------------------------------------------------------- #
# solver (control.py)
# --------------------------------------------------------------------------- #
# alphas = 1.1 ** (-np.arange(10, dtype=np.float32) ** 2) -- a float32 table (control.py:162)
def alpha_table(n=10):
    return 1.1 ** (-np.arange(n, dtype=np.float32) ** 2)


class OracleSolver:
    DELTA_0 = 2.0
    MU_MIN = 1e-6
    N_LS_ITER = 10

    def __init__(self, problem, N):
        self.p = problem
        self.N = N
        self.mu = 1.0
        self.delta = self.DELTA_0
        self.trace = []  # one dict per outer iteration
        self.n_backward = 0
        self.cond_log = None  # set to a list to record cond(Q_uu) of every step of every backward pass
        # 0: the reference's own evaluation order.  1, 2, ...: the SAME formulas in another, equally valid floating-point
        # order (products re-associated, the linear system handed to LAPACK with its unknowns permuted) -- never a
        # parity target, only the yardstick for how far rounding alone moves the reference's result on a scenario
        self.arith = 0

    def rollout(self, x0, U):
        """_rollout (control.py:80-93)."""
        N = U.shape[0]
        X = np.zeros((N + 1, self.p.n_x))
        X[0] = x0.flatten()
        J = 0.0
        for t in range(N):
            X[t + 1] = self.p.step(X[t], U[t])
            J += float(np.asarray(self.p.cost(X[t], U[t])).item())
        J += float(np.asarray(self.p.cost(X[-1], np.zeros(self.p.n_u), terminal=True)).item())
        return X, J

    def forward_pass(self, X, U, K, d, alpha):
        """_forward_pass (control.py:95-114)."""
        Xn = np.zeros((self.N + 1, self.p.n_x))
        Un = np.zeros((self.N, self.p.n_u))
        Xn[0] = X[0]
        J = 0.0
        for t in range(self.N):
            dx = Xn[t] - X[t]
            du = K[t] @ dx + alpha * d[t]
            Un[t] = U[t] + du
            Xn[t + 1] = self.p.step(Xn[t], Un[t])
            J += float(np.asarray(self.p.cost(Xn[t], Un[t])).item())
        J += float(np.asarray(self.p.cost(Xn[-1], np.zeros(self.p.n_u), terminal=True)).item())
        return Xn, Un, J

    def backward_pass(self, X, U):
        """_backward_pass (control.py:116-148)."""
        self.n_backward += 1
        n_x, n_u = self.p.n_x, self.p.n_u
        K = np.zeros((self.N, n_u, n_x))
        d = np.zeros((self.N, n_u))
        reg = self.mu * np.eye(n_x)
        L_x, _, L_xx, _, _ = self.p.quadraticize(X[-1], np.zeros(n_u), terminal=True)
        p, P = L_x, L_xx
        for t in range(self.N - 1, -1, -1):
            L_x, L_u, L_xx, L_uu, L_ux = self.p.quadraticize(X[t], U[t])
            A, B = self.p.linearize(X[t], U[t])
            Q_x = L_x + A.T @ p
            Q_u = L_u + B.T @ p
            Q_xx = L_xx + A.T @ P @ A
            Q_uu = L_uu + B.T @ (P + reg) @ B
            Q_ux = L_ux + B.T @ (P + reg) @ A
            if self.cond_log is not None:
                self.cond_log.append(np.linalg.cond(Q_uu))
            if self.arith:
                Pr = P + reg
                Q_xx = L_xx + A.T @ (P @ A)
                Q_uu = L_uu + B.T @ (Pr @ B)
                Q_ux = L_ux + B.T @ (Pr @ A)
                perm = np.random.default_rng(self.arith).permutation(n_u)
                Mp = Q_uu[np.ix_(perm, perm)]
                K[t][perm] = -np.linalg.solve(Mp, Q_ux[perm])
                d[t][perm] = -np.linalg.solve(Mp, Q_u[perm])
                p = Q_x + K[t].T @ Q_uu @ d[t] + K[t].T @ Q_u + Q_ux.T @ d[t]
                P = Q_xx + K[t].T @ Q_uu @ K[t] + K[t].T @ Q_ux + Q_ux.T @ K[t]
                P = 0.5 * (P + P.T)
                continue
            K[t] = -np.linalg.solve(Q_uu, Q_ux)
            d[t] = -np.linalg.solve(Q_uu, Q_u)
            p = Q_x + K[t].T @ Q_uu @ d[t] + K[t].T @ Q_u + Q_ux.T @ d[t]
            P = Q_xx + K[t].T @ Q_uu @ K[t] + K[t].T @ Q_ux + Q_ux.T @ K[t]
            P = 0.5 * (P + P.T)
        return K, d

    def _decrease_regularization(self):
        """control.py:232-237."""
        self.delta = min(1.0, self.delta) / self.DELTA_0
        self.mu *= self.delta
        if self.mu <= self.MU_MIN:
            self.mu = 0.0

    def solve(self, x0, U=None, n_lqr_iter=50, tol=1e-3, keep_gains=False):
        """solve (control.py:150-225) without the wall-clock t_kill branch.
        Returns X, U, J (J of the LAST candidate tried, control.py:225)."""
        if U is None:
            U = np.zeros((self.N, self.p.n_u))
        if U.shape != (self.N, self.p.n_u):
            raise ValueError
        self.mu, self.delta = 1.0, self.DELTA_0
        self.trace = []
        x0 = x0.reshape(-1, 1)
        converged = False
        alphas = alpha_table(self.N_LS_ITER)
        X, J_star = self.rollout(x0, U)
        self.J0 = J_star
        J = None
        for _ in range(n_lqr_iter):
            accept = False
            rec = {"mu": self.mu, "J_tried": []}
            K, d = self.backward_pass(X, U)
            if keep_gains:
                rec["K"], rec["d"], rec["X"], rec["U"] = K, d, X, U
            for k, alpha in enumerate(alphas):
                Xn, Un, J = self.forward_pass(X, U, K, d, alpha)
                rec["J_tried"].append(J)
                if J < J_star:
                    if abs((J_star - J) / J_star) < tol:
                        converged = True
                    X, U, J_star = Xn, Un, J
                    self._decrease_regularization()
                    accept = True
                    rec["alpha_index"] = k
                    break
            if not accept:
                rec["alpha_index"] = -1
            rec["J_star"] = J_star
            self.trace.append(rec)
            if not accept or converged:
                break
        self.converged = converged
        return X, U, J


# --------------------------------------------------------------------------- #
# DP-iLQR (distributed.py)
# --------------------------------------------------------------------------- #
def inter_graph_threshold(X, radius, x_dims, ids):
    """define_inter_graph_threshold (distributed.py:224-247)."""
    planning_radii = 2 * radius
    rel = pairwise_distance(X, x_dims)
    N = X.shape[0]
    step = max(N // 10, 1)
    rows = slice(0, N + 1, step)
    graph = {id_: [id_] for id_ in ids}
    for k, (i, j) in enumerate(itertools.combinations(ids, 2)):
        if np.any(rel[rows, k] < planning_radii):
            graph[i].append(j)
            graph[j].append(i)
    return {i: sorted(v) for i, v in graph.items()}


def solve_distributed(problem, X, U, radius, ignore_ids=(), n_lqr_iter=50, tol=1e-3, count=None):
    """solve_distributed, serial branch (distributed.py:25-77,99-103)."""
    s, c, a = problem.s, problem.c, problem.a
    N = U.shape[0]
    ids = problem.ids
    graph = inter_graph_threshold(X, radius, [s] * a, ids)
    x0_split = split_graph(X[np.newaxis, 0], [s] * a, graph)
    U_split = split_graph(U, [c] * a, graph)
    X_dec = np.zeros((N + 1, a * s))
    U_dec = np.zeros((N, a * c))
    info = {}
    for i, (id_, x0i, Ui) in enumerate(zip(ids, x0_split, U_split)):
        if id_ in ignore_ids:
            continue
        sub = problem.subproblem(graph[id_])
        solver = OracleSolver(sub, N)
        t0 = perf_counter()
        Xi, Ui_sol, _ = solver.solve(x0i, Ui, n_lqr_iter=n_lqr_iter, tol=tol)
        if count is not None:
            count[0] += solver.n_backward
        k = sub.ids.index(id_)  # extract (problem.py:49-64)
        X_dec[:, i * s:(i + 1) * s] = Xi[:, k * s:(k + 1) * s]
        U_dec[:, i * c:(i + 1) * c] = Ui_sol[:, k * c:(k + 1) * c]
        info[id_] = (perf_counter() - t0, graph[id_])
    _, J_full = OracleSolver(problem, N).rollout(X[0], U_dec)
    return X_dec, U_dec, J_full, info


def solve_rhc(problem, x0, N, radius=None, ignore_ids=(), centralized=True, n_d=2, step_size=1,
              dist_converge=None, t_diverge=None, U_init=None, n_lqr_iter=50, tol=1e-3):
    """solve_rhc, dist_converge mode (distributed.py:106-221).  ``U_init`` replaces the
    reference's ``np.random.rand(N, n_u) * 0.01`` draw (:152) when given."""
    xf = problem.xf
    a, s = problem.a, problem.s

    def dist_left(x):
        return np.linalg.norm((x - xf).reshape(a, s)[:, :n_d], axis=1)

    n_x, n_u = problem.n_x, problem.n_u
    xi = x0.reshape(1, -1)
    X = xi.copy()
    U = np.random.rand(N, n_u) * 0.01 if U_init is None else U_init.copy()
    solver = OracleSolver(problem, N)
    t = 0
    dt = problem.dt
    X_full = np.zeros((0, n_x))
    U_full = np.zeros((0, n_u))
    rounds = 0
    while np.any(dist_left(xi) > dist_converge):
        if centralized:
            X, U, J = solver.solve(xi, U, n_lqr_iter=n_lqr_iter, tol=tol)
        else:
            X, U, J, _ = solve_distributed(problem, X, U, radius, ignore_ids, n_lqr_iter=n_lqr_iter, tol=tol)
        rounds += 1
        xi = X[step_size]
        X_full = np.r_[X_full, X[:step_size]]
        U_full = np.r_[U_full, U[:step_size]]
        X = np.r_[X[step_size:], np.tile(X[-1], (step_size, 1))]
        U = np.r_[U[step_size:], np.zeros((step_size, n_u))]
        if t_diverge and t >= t_diverge:
            break
        t += step_size * dt
    if not X_full.size and not U_full.size:
        X_full = x0.copy()
        U_full = np.zeros((1, n_u))
    _, J_full = solver.rollout(x0, U_full)
    return X_full, U_full, J_full
