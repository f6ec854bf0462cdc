"""Cost objects of the drop-in API (reference dpilqr/cost.py).

The objects only *describe* the cost (goal, weights, radius); values and quadraticisations
are computed by the CUDA kernels through :class:`dpilqr_b200.engine.CompiledBatch`
(``dpilqr_game_cost`` and the fused linearise+quadraticise kernel).
"""

import abc

import numpy as np

from .util import uniform_block_diag


class Cost(abc.ABC):
    """Abstract cost (reference cost.py:19-34)."""

    @abc.abstractmethod
    def __call__(self, *args):
        pass

    @abc.abstractmethod
    def quadraticize():
        pass


def _single_batch(spec):
    from .engine import CompiledBatch

    return CompiledBatch([spec], 1)


def _quadraticize_spec(spec, x, u, terminal):
    """(L_x, L_u, L_xx, L_uu, L_ux) of one problem at one point via the fused kernel: a horizon-1
    trajectory [x, x] gives the running record at t=0 and the terminal record at t=1."""
    batch = _single_batch(spec)
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    u = np.asarray(u, dtype=np.float64).reshape(-1)
    stage, status = batch.linearize_quadraticize(np.stack([x, x])[None], u[None, None])
    if int(status.item()) & 4:
        # the reference trips `assert point_a.ndim == point_b.ndim` here (cost.py:279)
        raise AssertionError
    _, _, Lx, Lu, Lxx, Luu = batch.stage_to_dense(stage)
    t = 1 if terminal else 0
    n, m = batch.n, batch.m
    return (Lx[0, t].cpu().numpy(), Lu[0, t].cpu().numpy(), Lxx[0, t].cpu().numpy(), Luu[0, t].cpu().numpy(),
            np.zeros((m, n)))


class ReferenceCost(Cost):
    """Quadratic distance to a goal state (reference cost.py:37-107)."""

    _id = 0

    def __init__(self, xf, Q, R, Qf=None, id=None):
        if Qf is None:
            Qf = np.eye(Q.shape[0])
        if not id:
            id = ReferenceCost._id
            ReferenceCost._id += 1
        self.xf = np.asarray(xf).flatten()
        self.Q, self.R, self.Qf = Q, R, Qf
        self.id = id
        self.Q_plus_QT = Q + Q.T
        self.R_plus_RT = R + R.T
        self.nx = Q.shape[0]
        self.nu = R.shape[0]

    @property
    def x_dim(self):
        return self.Q.shape[0]

    @property
    def u_dim(self):
        return self.R.shape[0]

    @classmethod
    def _reset_ids(cls):
        cls._id = 0

    def _spec(self):
        from .dynamics import Model
        from .engine import ProblemSpec

        # any model with matching sizes serves: the cost kernels only use the sizes
        lib_sizes = {(4, 2): Model.DoubleInt4D, (6, 3): Model.DoubleInt6D, (3, 2): Model.Car3D, (12, 4): Model.Quadcopter12D,
                     (5, 2): Model.Bike5D}
        model = lib_sizes.get((self.nx, self.nu))
        if model is None:
            raise ValueError(f"no device model with sizes ({self.nx}, {self.nu})")
        f64 = lambda M: np.asarray(M, dtype=np.float64)
        return ProblemSpec([model.value], 1.0, self.nx, self.nu, [2], [f64(self.Q)], [f64(self.R)], [f64(self.Qf)], self.xf,
                           0.0, (1.0, 0.0), False, [self.id])

    def __call__(self, x, u, terminal=False):
        x = np.asarray(x, dtype=np.float64).reshape(1, 1, -1)
        u = np.asarray(u, dtype=np.float64).reshape(1, 1, -1)
        val = _single_batch(self._spec()).cost(x, u, terminal).cpu().numpy()
        # the reference returns a (1, 1) array for running costs and a scalar for terminal ones (cost.py:81-83)
        return val[0, 0] if terminal else val.reshape(1, 1)

    def quadraticize(self, x, u, terminal=False):
        return _quadraticize_spec(self._spec(), x, u, terminal)

    def __repr__(self):
        return f"ReferenceCost(\n\tQ: {self.Q},\n\tR: {self.R},\n\tQf: {self.Qf},\n\tid: {self.id}\n)"


class ProximityCost(Cost):
    """Pairwise thresholded-distance penalty (reference cost.py:110-171)."""

    def __init__(self, x_dims, radius, n_dims):
        self.x_dims = x_dims
        self.radius = radius
        self.n_dims = n_dims
        self.n_agents = len(x_dims)

    def _spec(self):
        from .dynamics import Model
        from .engine import ProblemSpec

        s = self.x_dims[0]
        sizes = {4: (Model.DoubleInt4D, 2), 6: (Model.DoubleInt6D, 3), 3: (Model.Car3D, 2), 12: (Model.Quadcopter12D, 4),
                 5: (Model.Bike5D, 2)}
        if s not in sizes:
            raise ValueError(f"no device model with {s} states")
        model, c = sizes[s]
        a = self.n_agents
        zero_s, zero_c = np.zeros((s, s)), np.zeros((c, c))
        # unit proximity weight, zero reference weight: the kernels then return the bare proximity terms
        return ProblemSpec([model.value] * a, 1.0, s, c, list(self.n_dims), [zero_s] * a, [zero_c] * a, [zero_s] * a,
                           np.zeros(a * s), self.radius, (0.0, 1.0), True, list(range(a))), c

    def __call__(self, x):
        if len(self.x_dims) == 1:
            return 0.0
        spec, c = self._spec()
        x = np.asarray(x, dtype=np.float64).reshape(1, 1, -1)
        return float(_single_batch(spec).cost(x, np.zeros((1, 1, self.n_agents * c)), False).item())

    def quadraticize(self, x):
        spec, c = self._spec()
        Lx, _, Lxx, _, _ = _quadraticize_spec(spec, x, np.zeros(self.n_agents * c), False)
        return Lx, Lxx


class GameCost(Cost):
    """Weighted sum of per-agent reference costs and the proximity cost (reference cost.py:174-266)."""

    def __init__(self, reference_costs, proximity_cost=None):
        if not proximity_cost:

            def proximity_cost(_):
                return 0.0

        self.ref_costs = reference_costs
        self.prox_cost = proximity_cost
        self.REF_WEIGHT = 1.0
        self.PROX_WEIGHT = 200.0
        self.x_dims = [rc.x_dim for rc in self.ref_costs]
        self.u_dims = [rc.u_dim for rc in self.ref_costs]
        self.ids = [rc.id for rc in self.ref_costs]
        self.n_agents = len(reference_costs)

    @property
    def xf(self):
        return np.concatenate([rc.xf for rc in self.ref_costs])

    def _spec(self):
        from .dynamics import Model
        from .engine import ProblemSpec

        s, c = self.x_dims[0], self.u_dims[0]
        sizes = {(4, 2): Model.DoubleInt4D, (6, 3): Model.DoubleInt6D, (3, 2): Model.Car3D, (12, 4): Model.Quadcopter12D,
                 (5, 2): Model.Bike5D}
        if (s, c) not in sizes:
            raise ValueError(f"no device model with sizes ({s}, {c})")
        has_prox = isinstance(self.prox_cost, ProximityCost)
        f64 = lambda M: np.asarray(M, dtype=np.float64)
        return ProblemSpec(
            [sizes[(s, c)].value] * self.n_agents, 1.0, s, c,
            list(self.prox_cost.n_dims) if has_prox else [2] * self.n_agents,
            [f64(rc.Q) for rc in self.ref_costs], [f64(rc.R) for rc in self.ref_costs], [f64(rc.Qf) for rc in self.ref_costs],
            self.xf, self.prox_cost.radius if has_prox else 0.0, (self.REF_WEIGHT, self.PROX_WEIGHT), has_prox, self.ids)

    def __call__(self, x, u, terminal=False):
        x = np.asarray(x, dtype=np.float64).reshape(1, 1, -1)
        u = np.asarray(u, dtype=np.float64).reshape(1, 1, -1)
        return _single_batch(self._spec()).cost(x, u, terminal).cpu().numpy().reshape(1, 1)

    def quadraticize(self, x, u, terminal=False):
        return _quadraticize_spec(self._spec(), x, u, terminal)

    def split(self, graph):
        """One GameCost per graph entry (reference cost.py:241-262)."""
        n_states = self.ref_costs[0].x_dim
        radius = self.prox_cost.radius
        n_dims = self.prox_cost.n_dims
        out = []
        for members in graph.values():
            refs, dims = [], []
            for n_dim, rc in zip(n_dims, self.ref_costs):
                if rc.id in members:
                    refs.append(rc)
                    dims.append(n_dim)
            out.append(GameCost(refs, ProximityCost([n_states] * len(members), radius, dims)))
        return out

    def __repr__(self):
        return f"GameCost(\n\tids: {[rc.id for rc in self.ref_costs]},\n\tprox_cost: {self.prox_cost}\n)"


def quadraticize_distance(point_a, point_b, radius, n_d):
    """Gradient / Hessian of the thresholded distance between two points (reference cost.py:269-315),
    evaluated by the proximity branch of the fused kernel on a two-agent problem."""
    assert point_a.ndim == point_b.ndim
    x = np.array([point_a.x, point_a.y, point_a.z, point_b.x, point_b.y, point_b.z], dtype=np.float64)
    Lx, Lxx = ProximityCost([3, 3], radius, [n_d, n_d]).quadraticize(x)
    return Lx[:n_d], Lxx[:n_d, :n_d]


def quadraticize_finite_difference(cost, x, u, terminal=False, jac_eps=None):
    """Finite-difference quadraticisation (reference cost.py:318-349).  Test helper; SciPy."""
    from scipy.optimize import approx_fprime

    if not jac_eps:
        jac_eps = np.sqrt(np.finfo(float).eps)
    hess_eps = np.sqrt(jac_eps)
    n_x, n_u = x.shape[0], u.shape[0]

    def scalar(x, u):
        return float(np.asarray(cost(x, u, terminal)).item())

    def Lx(x, u):
        return approx_fprime(x, lambda x: scalar(x, u), jac_eps)

    def Lu(x, u):
        return approx_fprime(u, lambda u: scalar(x, u), jac_eps)

    L_xx = np.vstack([approx_fprime(x, lambda x: Lx(x, u)[i], hess_eps) for i in range(n_x)])
    L_uu = np.vstack([approx_fprime(u, lambda u: Lu(x, u)[i], hess_eps) for i in range(n_u)])
    L_ux = np.vstack([approx_fprime(x, lambda x: Lu(x, u)[i], hess_eps) for i in range(n_u)])
    return Lx(x, u), Lu(x, u), L_xx, L_uu, L_ux
