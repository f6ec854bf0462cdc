"""Batched device engine: problem compiler + thin wrappers over the C ABI.

This is the one place where the reference's object graph (ilqrProblem ->
MultiDynamicalModel + GameCost, reference problem.py:15-24) is flattened into the
``dpilqr_batch`` descriptor the CUDA kernels consume.  PyTorch is used only as the
device-memory / stream plumbing; all arithmetic happens in libdpilqr_b200.so.

Not in the reference (which solves one problem at a time): :class:`CompiledBatch` is the
batched front door that the drop-in single-problem API routes through with batch size 1.
"""

import ctypes

import numpy as np
import torch

from . import _native
from ._native import BatchStruct, NativeError, SolveOpts

N_LS_ITER = 10  # ilqrSolver.N_LS_ITER, reference control.py:51


def default_device():
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device visible; dpilqr_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class ProblemSpec:
    """Flat host-side description of one (sub)problem."""

    __slots__ = ("models", "dt", "s", "c", "n_dims", "Q", "R", "Qf", "xf", "radius", "weights", "has_prox", "ids")

    def __init__(self, models, dt, s, c, n_dims, Q, R, Qf, xf, radius, weights, has_prox, ids):
        self.models = [int(m) for m in models]
        self.dt = float(dt)
        self.s, self.c = int(s), int(c)
        self.n_dims = [int(v) for v in n_dims]
        self.Q, self.R, self.Qf = Q, R, Qf
        self.xf = np.ascontiguousarray(xf, dtype=np.float64).reshape(-1)
        self.radius = float(radius)
        self.weights = (float(weights[0]), float(weights[1]))
        self.has_prox = bool(has_prox)
        self.ids = list(ids)

    @property
    def a(self):
        return len(self.models)

    @property
    def key(self):
        return (self.a, self.s, self.c, self.dt)

    def subset(self, keep):
        """Sub-problem over the agents at positions ``keep`` (original order): what
        MultiDynamicalModel.split / GameCost.split build for one graph entry (reference
        dynamics.py:188-198, cost.py:241-262).  A fresh GameCost carries the default weights."""
        return ProblemSpec(
            [self.models[i] for i in keep], self.dt, self.s, self.c, [self.n_dims[i] for i in keep],
            [self.Q[i] for i in keep], [self.R[i] for i in keep], [self.Qf[i] for i in keep],
            np.concatenate([self.xf[i * self.s:(i + 1) * self.s] for i in keep]),
            self.radius, (1.0, 200.0), True, [self.ids[i] for i in keep],
        )


def spec_from_problem(problem):
    """Compile an ``ilqrProblem``-like object (duck typed) into a :class:`ProblemSpec`.

    Anything that cannot be expressed for the kernels raises: there is no CPU fallback."""
    from .cost import GameCost, ProximityCost, ReferenceCost
    from .dynamics import DynamicalModel, MultiDynamicalModel

    dyn, cost = problem.dynamics, problem.game_cost
    if isinstance(dyn, MultiDynamicalModel):
        submodels = list(dyn.submodels)
    elif isinstance(dyn, DynamicalModel):
        submodels = [dyn]
    else:
        raise TypeError(f"unsupported dynamics object {type(dyn).__name__}")
    for sm in submodels:
        if getattr(sm, "model", None) is None:
            raise TypeError(f"{type(sm).__name__} has no native model id; only the built-in model classes run on the GPU")
    s, c = submodels[0].n_x, submodels[0].n_u
    if any(sm.n_x != s or sm.n_u != c for sm in submodels):
        raise ValueError("all agents must share per-agent state/control sizes (zero-pad heterogeneous teams)")
    dts = {float(sm.dt) for sm in submodels}
    if len(dts) != 1:
        raise ValueError("all agents must share one dt")
    if isinstance(cost, GameCost):
        refs = list(cost.ref_costs)
        prox = cost.prox_cost
        weights = (cost.REF_WEIGHT, cost.PROX_WEIGHT)
        if isinstance(prox, ProximityCost):
            has_prox, radius, n_dims = True, prox.radius, list(prox.n_dims)
        else:
            has_prox, radius, n_dims = False, 0.0, [2] * len(refs)
    elif isinstance(cost, ReferenceCost):
        refs, has_prox, radius, n_dims, weights = [cost], False, 0.0, [2], (1.0, 0.0)
    else:
        raise TypeError(f"unsupported cost object {type(cost).__name__}; only GameCost / ReferenceCost run on the GPU")
    if len(refs) != len(submodels):
        raise ValueError("dynamics and cost describe different numbers of agents")
    if len(n_dims) != len(refs):
        raise ValueError("ProximityCost.n_dims must have one entry per agent")
    for rc in refs:
        if rc.Q.shape != (s, s) or rc.R.shape != (c, c) or rc.Qf.shape != (s, s) or rc.xf.size != s:
            raise ValueError("ReferenceCost shapes do not match the per-agent model sizes")
    ids = [sm.id for sm in submodels]
    return ProblemSpec(
        [sm.model.value for sm in submodels], dts.pop(), s, c, n_dims,
        [np.asarray(rc.Q, dtype=np.float64) for rc in refs], [np.asarray(rc.R, dtype=np.float64) for rc in refs],
        [np.asarray(rc.Qf, dtype=np.float64) for rc in refs], np.concatenate([np.asarray(rc.xf, dtype=np.float64).reshape(-1) for rc in refs]),
        radius, weights, has_prox, ids,
    )


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class CompiledBatch:
    """Device-resident descriptor for B problems that share (a, s, c, N, dt)."""

    def __init__(self, specs, N, device=None):
        if not specs:
            raise ValueError("empty batch")
        key = specs[0].key
        if any(sp.key != key for sp in specs):
            raise ValueError("a CompiledBatch needs uniform (agents, s, c, dt); bin the problems first")
        _native.require_device()
        self.device = torch.device(device) if device is not None else default_device()
        self.B, self.N = len(specs), int(N)
        self.a, self.s, self.c, self.dt = key
        self.n, self.m = self.a * self.s, self.a * self.c
        # de-duplicated cost tables
        table, cost_idx = {}, np.empty((self.B, self.a), dtype=np.int32)
        Qs, Rs, Qfs = [], [], []
        last = (None, None, None, -1)
        for b, sp in enumerate(specs):
            for i in range(self.a):
                Q, R, Qf = sp.Q[i], sp.R[i], sp.Qf[i]
                if Q is last[0] and R is last[1] and Qf is last[2]:
                    cost_idx[b, i] = last[3]
                    continue
                k = (Q.tobytes(), R.tobytes(), Qf.tobytes())
                idx = table.get(k)
                if idx is None:
                    idx = table[k] = len(Qs)
                    Qs.append(Q), Rs.append(R), Qfs.append(Qf)
                cost_idx[b, i] = idx
                last = (Q, R, Qf, idx)
        dev = self.device
        f64 = dict(dtype=torch.float64, device=dev)
        models_np = np.array([sp.models for sp in specs], dtype=np.int32)
        self.model_hint = int(models_np.flat[0]) + 1 if np.all(models_np == models_np.flat[0]) else 0
        self.t_model = torch.as_tensor(models_np).to(dev)
        self.t_ndims = torch.as_tensor(np.array([sp.n_dims for sp in specs], dtype=np.int32)).to(dev)
        self.t_cidx = torch.as_tensor(cost_idx).to(dev)
        self.t_Q = torch.as_tensor(np.stack(Qs)).to(**f64).contiguous()
        self.t_R = torch.as_tensor(np.stack(Rs)).to(**f64).contiguous()
        self.t_Qf = torch.as_tensor(np.stack(Qfs)).to(**f64).contiguous()
        self.t_xf = torch.as_tensor(np.stack([sp.xf for sp in specs])).to(**f64).contiguous()
        self.t_radius = torch.as_tensor(np.array([sp.radius for sp in specs], dtype=np.float64)).to(dev)
        self.t_weights = torch.as_tensor(np.array([sp.weights for sp in specs], dtype=np.float64)).to(dev).contiguous()
        self.t_hasprox = torch.as_tensor(np.array([sp.has_prox for sp in specs], dtype=np.int32)).to(dev)
        # the bounded line search needs every stage cost >= 0: PSD reference-cost matrices, non-negative weights
        def _psd(M):
            return bool(np.all(np.linalg.eigvalsh(0.5 * (M + M.T)) >= -1e-12 * max(1.0, float(np.abs(M).max()))))

        self.costs_nonnegative = (all(_psd(Q) and _psd(R) and _psd(Qf) for Q, R, Qf in zip(Qs, Rs, Qfs))
                                  and all(sp.weights[0] >= 0.0 and sp.weights[1] >= 0.0 for sp in specs))
        self.struct = self._make_struct(self.N)
        self.stage_stride = int(_native.lib().dpilqr_stage_stride(self.a, self.s, self.c))

    @classmethod
    def from_tensors(cls, N, a, s, c, dt, t_model, t_ndims, t_cidx, t_Q, t_R, t_Qf, t_xf, t_radius, t_weights, t_hasprox,
                     model_hint, costs_nonnegative, device):
        """Descriptor built directly from device tensors (no per-problem Python objects): the batched DP-iLQR and
        receding-horizon drivers derive the descriptors of their sub-batches with gathers on the device."""
        self = cls.__new__(cls)
        _native.require_device()
        self.device = torch.device(device)
        self.B, self.N = int(t_model.shape[0]), int(N)
        self.a, self.s, self.c, self.dt = int(a), int(s), int(c), float(dt)
        self.n, self.m = self.a * self.s, self.a * self.c
        i32 = dict(dtype=torch.int32, device=self.device)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.t_model = t_model.to(**i32).contiguous()
        self.t_ndims = t_ndims.to(**i32).contiguous()
        self.t_cidx = t_cidx.to(**i32).contiguous()
        self.t_Q, self.t_R, self.t_Qf = t_Q, t_R, t_Qf
        self.t_xf = t_xf.to(**f64).contiguous()
        self.t_radius = t_radius.to(**f64).contiguous()
        self.t_weights = t_weights.to(**f64).contiguous()
        self.t_hasprox = t_hasprox.to(**i32).contiguous()
        self.model_hint = int(model_hint)
        self.costs_nonnegative = bool(costs_nonnegative)
        self.struct = self._make_struct(self.N)
        self.stage_stride = int(_native.lib().dpilqr_stage_stride(self.a, self.s, self.c))
        return self

    def select(self, index, N=None):
        """Sub-batch of the problems at ``index`` (a device int64 tensor), optionally with another horizon."""
        return CompiledBatch.from_tensors(
            self.N if N is None else N, self.a, self.s, self.c, self.dt, self.t_model[index], self.t_ndims[index], self.t_cidx[index],
            self.t_Q, self.t_R, self.t_Qf, self.t_xf[index], self.t_radius[index], self.t_weights[index], self.t_hasprox[index],
            self.model_hint, self.costs_nonnegative, self.device)

    def _make_struct(self, horizon):
        return BatchStruct(
            self.B, self.a, self.s, self.c, int(horizon), int(self.t_Q.shape[0]), self.dt,
            self.t_model.data_ptr(), self.t_ndims.data_ptr(), self.t_cidx.data_ptr(), self.t_Q.data_ptr(),
            self.t_R.data_ptr(), self.t_Qf.data_ptr(), self.t_xf.data_ptr(), self.t_radius.data_ptr(),
            self.t_weights.data_ptr(), self.t_hasprox.data_ptr(), self.model_hint, 0,
        )

    # ---------------------------------------------------------------- helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, arr, shape):
        if not isinstance(arr, torch.Tensor):
            arr = np.ascontiguousarray(arr, dtype=np.float64)
            if not arr.flags.writeable:  # (read-only views, e.g. arrays of an .npz: torch wants a writable buffer)
                arr = arr.copy()
            arr = torch.from_numpy(arr)
        t = arr
        t = t.to(device=self.device, dtype=torch.float64, non_blocking=True).contiguous()
        if tuple(t.shape) != tuple(shape):
            t = t.reshape(shape)
        return t

    def _empty(self, *shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # ---------------------------------------------------------------- kernels
    def rollout(self, x0, U):
        """_rollout for every problem (reference control.py:80-93): returns X [B,N+1,n], J [B]."""
        x0 = self._dev(x0, (self.B, self.n))
        U = self._dev(U, (self.B, self.N, self.m))
        X, Uc, J = self._empty(self.B, 1, self.N + 1, self.n), self._empty(self.B, 1, self.N, self.m), self._empty(self.B, 1)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().dpilqr_rollout_linesearch(
                ctypes.byref(self.struct), _ptr(x0), _ptr(U), None, None, None, 1, _ptr(X), _ptr(Uc), _ptr(J), self._stream()))
        return X[:, 0], J[:, 0]

    def forward_pass(self, X, U, K, d, alphas=None):
        """_forward_pass for every problem and every alpha at once (reference control.py:95-114)."""
        X = self._dev(X, (self.B, self.N + 1, self.n))
        U = self._dev(U, (self.B, self.N, self.m))
        K = self._dev(K, (self.B, self.N, self.m, self.n))
        d = self._dev(d, (self.B, self.N, self.m))
        if alphas is None:
            n_alpha, a_arr = N_LS_ITER, None
        else:
            a_np = np.ascontiguousarray(np.atleast_1d(np.asarray(alphas, dtype=np.float64)))
            n_alpha, a_arr = a_np.size, a_np.ctypes.data_as(ctypes.c_void_p)
        Xc = self._empty(self.B, n_alpha, self.N + 1, self.n)
        Uc = self._empty(self.B, n_alpha, self.N, self.m)
        Jc = self._empty(self.B, n_alpha)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().dpilqr_rollout_linesearch(
                ctypes.byref(self.struct), _ptr(X), _ptr(U), _ptr(K), _ptr(d), a_arr, n_alpha, _ptr(Xc), _ptr(Uc), _ptr(Jc),
                self._stream()))
        return Xc, Uc, Jc

    def linearize_quadraticize(self, X, U):
        """Fused linearise+quadraticise of whole trajectories -> structured stage records."""
        X = self._dev(X, (self.B, self.N + 1, self.n))
        U = self._dev(U, (self.B, self.N, self.m))
        stage = self._empty(self.B, self.N + 1, self.stage_stride)
        status = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().dpilqr_linearize_quadraticize(
                ctypes.byref(self.struct), _ptr(X), _ptr(U), _ptr(stage), _ptr(status), self._stream()))
        return stage, status

    def stage_to_dense(self, stage):
        """Dense (A, B, L_x, L_u, L_xx, L_uu) views of stage records, for hooks and tests."""
        R = self.N + 1
        A, Bm = self._empty(self.B, R, self.n, self.n), self._empty(self.B, R, self.n, self.m)
        Lx, Lu = self._empty(self.B, R, self.n), self._empty(self.B, R, self.m)
        Lxx, Luu = self._empty(self.B, R, self.n, self.n), self._empty(self.B, R, self.m, self.m)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().dpilqr_stage_to_dense(
                ctypes.byref(self.struct), _ptr(stage), _ptr(A), _ptr(Bm), _ptr(Lx), _ptr(Lu), _ptr(Lxx), _ptr(Luu), self._stream()))
        return A, Bm, Lx, Lu, Lxx, Luu

    def backward(self, stage, mu):
        """_backward_pass for every problem (reference control.py:116-148): K [B,N,m,n], d [B,N,m]."""
        mu_t = self._dev(np.broadcast_to(np.asarray(mu, dtype=np.float64), (self.B,)) if not isinstance(mu, torch.Tensor) else mu, (self.B,))
        K, d = self._empty(self.B, self.N, self.m, self.n), self._empty(self.B, self.N, self.m)
        status = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().dpilqr_backward(
                ctypes.byref(self.struct), _ptr(stage), _ptr(mu_t), _ptr(K), _ptr(d), _ptr(status), self._stream()))
        return K, d, status

    def cost(self, X, U=None, terminal=False):
        """GameCost value at ``rows`` points per problem: X [B,rows,n], U [B,rows,m] -> [B,rows]."""
        X = X if isinstance(X, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64))
        rows = X.shape[1]
        X = self._dev(X, (self.B, rows, self.n))
        U = None if (U is None or terminal) else self._dev(U, (self.B, rows, self.m))
        if U is None and not terminal:
            raise ValueError("U is required for a running cost")
        L = self._empty(self.B, rows)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().dpilqr_game_cost(
                ctypes.byref(self.struct), rows, _ptr(X), _ptr(U), int(bool(terminal)), _ptr(L), self._stream()))
        return L

    # ---------------------------------------------------------------- full solve
    def solve(self, x0, U0, n_lqr_iter=50, tol=1e-3, t_kill=None, n_alpha=N_LS_ITER, trace=False, profile=False, bounded_search=True):
        """ilqrSolver.solve for the whole batch (reference control.py:150-225).

        ``x0`` [B,n] and ``U0`` [B,N,m] may be NumPy arrays, host (ideally pinned) or CUDA
        tensors.  Returns a dict of CUDA tensors: X, U, J (last tried cost), J_star, iters,
        status (+ trace_alpha / trace_mu / trace_J when ``trace``) and ``total_iters``.

        ``bounded_search`` (default on, used only when every stage cost is provably >= 0): a candidate of the line
        search stops rolling out once its accumulated cost exceeds the problem's best cost -- it is rejected already;
        its entry of ``trace_J`` then reads ``_native.J_ABORTED``.  Accepted steps, iteration counts and all returned
        trajectories and costs are those of the full search."""
        if n_lqr_iter < 0:
            raise ValueError("n_lqr_iter must be >= 0")
        x0 = self._dev(x0, (self.B, self.n))
        U0 = self._dev(U0, (self.B, self.N, self.m))
        lib = _native.lib()
        ws_bytes = int(lib.dpilqr_workspace_bytes(self.B, self.a, self.s, self.c, self.N, n_alpha))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        X, U = self._empty(self.B, self.N + 1, self.n), self._empty(self.B, self.N, self.m)
        J, Js = self._empty(self.B), self._empty(self.B)
        iters = self._empty(self.B, dtype=torch.int32)
        status = self._empty(self.B, dtype=torch.int32)
        ta = tm = tj = None
        if trace:
            ta = self._empty(self.B, max(n_lqr_iter, 1), dtype=torch.int32)
            tm = self._empty(self.B, max(n_lqr_iter, 1))
            tj = self._empty(self.B, max(n_lqr_iter, 1), n_alpha)
        opts = SolveOpts(int(n_lqr_iter), int(n_alpha), float(tol), float(t_kill) if t_kill else 0.0, int(bool(trace)), int(bool(profile)),
                         int(bool(bounded_search) and self.costs_nonnegative), 0)
        with torch.cuda.device(self.device):
            total = _native.check(lib.dpilqr_solve_batch(
                ctypes.byref(self.struct), ctypes.byref(opts), _ptr(x0), _ptr(U0), _ptr(X), _ptr(U), _ptr(J), _ptr(Js),
                _ptr(iters), _ptr(status), _ptr(ta), _ptr(tm), _ptr(tj), _ptr(ws), ws_bytes, self._stream()))
        out = dict(X=X, U=U, J=J, J_star=Js, iters=iters, status=status, total_iters=int(total))
        if trace:
            out.update(trace_alpha=ta, trace_mu=tm, trace_J=tj)
        return out


def bin_specs(specs):
    """Group problem indices by (agents, s, c, dt) so each bin is one kernel configuration."""
    bins = {}
    for k, sp in enumerate(specs):
        bins.setdefault(sp.key, []).append(k)
    return bins


def raise_for_status(status):
    """The exceptions the reference raises for one problem, from its status word: the ``Point.ndim`` assertion of
    quadraticize_distance (reference cost.py:279) and np.linalg.solve's LinAlgError for a singular Q_uu
    (reference control.py:141-142)."""
    status = int(status)
    if status & _native.ST_POINT_NDIM:
        raise AssertionError
    if status & _native.ST_SINGULAR:
        raise np.linalg.LinAlgError("Singular matrix")


def solve_specs(specs, x0s, U0s, N, device=None, on_error="raise", **kw):
    """Solve a ragged list of problems: bin, solve each bin in one batch, scatter back.

    Returns a list of per-problem dicts of NumPy arrays (X, U, J, J_star, iters, status[, trace_*]).
    ``on_error="raise"`` (default): the first problem, in list order, whose solve the reference would have aborted
    with an exception raises that exception here too (every sub-problem of the reference goes through
    ilqrSolver.solve, so solve_distributed / solve_rhc / selfish_warmstart abort the same way);
    ``on_error="status"`` leaves the decision to the caller (per-problem ``status`` words)."""
    if on_error not in ("raise", "status"):
        raise ValueError("on_error must be 'raise' or 'status'")
    results = [None] * len(specs)
    for key, idxs in bin_specs(specs).items():
        batch = CompiledBatch([specs[k] for k in idxs], N, device)
        x0 = np.stack([np.asarray(x0s[k], dtype=np.float64).reshape(-1) for k in idxs])
        U0 = np.stack([np.asarray(U0s[k], dtype=np.float64) for k in idxs])
        out = batch.solve(x0, U0, **kw)
        host = {k: v.cpu().numpy() for k, v in out.items() if isinstance(v, torch.Tensor)}
        for j, k in enumerate(idxs):
            results[k] = {name: arr[j] for name, arr in host.items()}
    if on_error == "raise":
        for res in results:
            raise_for_status(res["status"])
    return results


class SolvePipeline:
    """Keeps ``depth`` batched solves in flight on one device (not in the reference, whose batches are a process pool).

    One worker thread and one CUDA stream per solve in flight.  The launches of concurrent solves interleave -- the
    line-search launches are bound by the latency of their 51-step chains, not by the machine -- and the library runs
    the tail of a solve (the few problems that need many more iterations than the rest: a tenth of the time of a
    4096-scenario batch for 3 % of its work) on a high-priority stream beside the other solves (csrc/solver.cu).
    Throughput of a stream of batches then follows their work rather than the latency of their stragglers.  Each solve
    in flight holds its own workspace (18 GB for 4096 ten-drone scenarios).

        pipe = SolvePipeline(device, depth=2)
        for out in pipe.map(lambda args: batch.solve(*args), jobs): ...
    """

    def __init__(self, device=None, depth=2):
        from concurrent.futures import ThreadPoolExecutor

        _native.require_device()
        self.device = torch.device(device) if device is not None else default_device()
        self.depth = max(1, int(depth))
        self._pool = ThreadPoolExecutor(max_workers=self.depth, thread_name_prefix="dpilqr-solve")
        self._streams = {}

    def _run(self, fn, item):
        import threading

        torch.cuda.set_device(self.device)
        key = threading.get_ident()
        stream = self._streams.get(key)
        if stream is None:
            stream = self._streams[key] = torch.cuda.Stream(self.device)
        with torch.cuda.stream(stream):
            out = fn(item)
            stream.synchronize()
        return out

    def submit(self, fn, item):
        return self._pool.submit(self._run, fn, item)

    def map(self, fn, items):
        """Results in submission order; at most ``depth`` items are in flight."""
        from collections import deque

        pending = deque()
        for item in items:
            if len(pending) >= self.depth:
                yield pending.popleft().result()
            pending.append(self.submit(fn, item))
        while pending:
            yield pending.popleft().result()

    def close(self):
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
