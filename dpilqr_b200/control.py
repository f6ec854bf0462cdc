"""ilqrSolver drop-in (reference dpilqr/control.py) on top of the batched CUDA engine.

``solve`` is a batch-of-one call of ``dpilqr_solve_batch``; the three internal hooks
``_rollout`` / ``_forward_pass`` / ``_backward_pass`` that the reference's own callers and
tests use (reference distributed.py:100-101,211) are single launches of the corresponding
kernels.
"""

import numpy as np

from . import _native
from .engine import CompiledBatch, raise_for_status, spec_from_problem


class ilqrSolver:
    """Iterative LQR solver for one (possibly multi-agent) problem (reference control.py:15-249)."""

    DELTA_0 = 2.0
    MU_MIN = 1e-6
    MU_MAX = 1e3
    N_LS_ITER = 10

    def __init__(self, problem, N=10):
        self.problem = problem
        self.N = N
        self._batch = None
        self._batch_key = None
        self._batches = {}
        self._reset_regularization()

    # ---- reference properties (control.py:58-78)
    @property
    def cost(self):
        return self.problem.game_cost

    @property
    def dynamics(self):
        return self.problem.dynamics

    @property
    def n_x(self):
        return self.problem.dynamics.n_x

    @property
    def n_u(self):
        return self.problem.dynamics.n_u

    @property
    def dt(self):
        return self.problem.dynamics.dt

    # ---- engine plumbing
    def _compiled(self, N=None):
        """Device descriptor of the problem for horizon N.  The reference re-reads problem.dynamics / problem.game_cost on
        every call; here the flattened spec is fingerprinted (goals, weights, radius, n_dims, cost matrices, dt, models), so a
        caller that moves the goals or edits a cost between solves gets a fresh descriptor, and one descriptor is kept per
        horizon (solve_rhc alternates between the planning horizon and its final rollout)."""
        N = self.N if N is None else N
        spec = spec_from_problem(self.problem)
        finger = hash((spec.key, tuple(spec.models), tuple(spec.n_dims), spec.xf.tobytes(), float(spec.radius), tuple(spec.weights),
                       bool(spec.has_prox), tuple(np.asarray(M).tobytes() for M in spec.Q), tuple(np.asarray(M).tobytes() for M in spec.R),
                       tuple(np.asarray(M).tobytes() for M in spec.Qf)))
        if self._batch_key != finger:
            self._batches, self._batch_key = {}, finger
        batch = self._batches.get(N)
        if batch is None:
            batch = self._batches[N] = CompiledBatch([spec], N)
        self._batch = batch
        return batch

    def _rollout(self, x0, U):
        """Roll the controls out from x0 (reference control.py:80-93)."""
        U = np.asarray(U, dtype=np.float64)
        batch = self._compiled(U.shape[0])
        X, J = batch.rollout(np.asarray(x0, dtype=np.float64).reshape(1, -1), U[None])
        return X[0].cpu().numpy(), float(J[0].item())

    def _forward_pass(self, X, U, K, d, α):
        """One line-search candidate (reference control.py:95-114)."""
        batch = self._compiled()
        Xc, Uc, Jc = batch.forward_pass(np.asarray(X)[None], np.asarray(U)[None], np.asarray(K)[None], np.asarray(d)[None], [float(α)])
        return Xc[0, 0].cpu().numpy(), Uc[0, 0].cpu().numpy(), float(Jc[0, 0].item())

    def _backward_pass(self, X, U):
        """Riccati recursion around (X, U) with the current μ (reference control.py:116-148)."""
        batch = self._compiled()
        stage, st1 = batch.linearize_quadraticize(np.asarray(X)[None], np.asarray(U)[None])
        K, d, st2 = batch.backward(stage, self.μ)
        raise_for_status(int(st1.item()) | int(st2.item()))  # reference cost.py:279, control.py:141
        return K[0].cpu().numpy(), d[0].cpu().numpy()

    def solve(self, x0, U=None, n_lqr_iter=50, tol=1e-3, t_kill=None, verbose=True):
        """Solve from x0 with warm start U; returns (X, U, J) like reference control.py:150-225.

        J is the cost of the last candidate tried, as in the reference (control.py:225).  The
        per-iteration trace is kept in ``self.last_trace``."""
        if U is None:
            U = np.zeros((self.N, self.n_u))
        U = np.asarray(U)
        if U.shape != (self.N, self.n_u):
            raise ValueError
        if n_lqr_iter < 1:
            raise UnboundLocalError("cannot access local variable 'J'")  # what the reference does for n_lqr_iter=0
        self._reset_regularization()
        batch = self._compiled()
        out = batch.solve(np.asarray(x0, dtype=np.float64).reshape(1, -1), U[None].astype(np.float64), n_lqr_iter=n_lqr_iter,
                          tol=tol, t_kill=t_kill, trace=True)
        status = int(out["status"][0].item())
        raise_for_status(status)
        iters = int(out["iters"][0].item())
        acc = out["trace_alpha"][0, :iters].cpu().numpy()
        mus = out["trace_mu"][0, :iters].cpu().numpy()
        Jt = out["trace_J"][0, :iters].cpu().numpy()
        self.last_trace = {"alpha_index": acc, "mu": mus, "J_tried": Jt, "iters": iters, "status": status}
        # leave μ, Δ where the reference's schedule would have left them (control.py:232-237)
        for k in acc:
            if k >= 0:
                self._decrease_regularization()
        if verbose:
            J_run = None
            for i in range(iters):
                if acc[i] >= 0:
                    J_run = Jt[i, acc[i]]
                    if i + 1 < iters or not (status & _native.ST_CONVERGED):
                        print(f"{i+1}/{n_lqr_iter}\tJ: {J_run:g}\tμ: {mus[i]:g}")
        return out["X"][0].cpu().numpy(), out["U"][0].cpu().numpy(), float(out["J"][0].item())

    # ---- regularisation schedule (reference control.py:227-242)
    def _reset_regularization(self):
        self.μ = 1.0
        self.Δ = self.DELTA_0

    def _decrease_regularization(self):
        self.Δ = min(1.0, self.Δ) / self.DELTA_0
        self.μ *= self.Δ
        if self.μ <= self.MU_MIN:
            self.μ = 0.0

    def _increase_regularization(self):
        self.Δ = max(1.0, self.Δ) * self.DELTA_0
        self.μ = max(self.MU_MIN, self.μ * self.Δ)

    def __repr__(self):
        return (f"iLQR(\n\tdynamics: {self.dynamics},\n\tcost: {self.cost},\n\tN: {self.N},\n\tdt: {self.dt},"
                f"\n\tμ: {self.μ},\n\tΔ: {self.Δ}\n)")


class RecedingHorizonController:
    """Legacy generator-style receding horizon wrapper (reference control.py:253-326)."""

    def __init__(self, x0, controller, step_size=1):
        self.x = x0
        self._controller = controller
        self.step_size = step_size

    @property
    def N(self):
        return self._controller.N

    def solve(self, U0, J_converge=1.0, **kwargs):
        i = 0
        U = U0
        while True:
            print("-" * 50 + f"\nHorizon {i}")
            i += 1
            if U.shape != (self._controller.N, self._controller.n_u):
                raise RuntimeError
            X, U, J = self._controller.solve(self.x, U, **kwargs)
            self.x = X[self.step_size]
            yield X[: self.step_size], U[: self.step_size], J
            U = np.vstack([U[self.step_size:], np.zeros((self.step_size, self._controller.n_u))])
            if J < J_converge:
                print("Converged!")
                break
