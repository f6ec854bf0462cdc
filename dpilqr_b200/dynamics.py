"""Dynamical models of the drop-in API, evaluated by the CUDA model library.

Mirrors the public surface of reference dpilqr/dynamics.py and dpilqr/bbdynamicswrap.pyx
(``Model``, ``f``, ``integrate``, ``linearize``, the model classes, ``MultiDynamicalModel``).
Every evaluation -- even of a single (x, u) -- is a launch of the device library through
the C ABI; there is no host implementation of the ODEs in this package.
"""

import abc
import ctypes
from enum import Enum

import numpy as np
import torch

from . import _native
from .util import split_agents_gen, uniform_block_diag


class Model(Enum):
    """Native model ids (reference bbdynamicswrap.pyx:8-16; Bike5D appended)."""

    DoubleInt4D = 0
    DoubleInt6D = 1
    Car3D = 2
    Unicycle4D = 3
    Quadcopter6D = 4
    Human6D = 5
    HumanLin6D = 6
    Quadcopter12D = 7
    Bike5D = 8


def _device_eval(mode, model, dt, x, u):
    """Run the batched device entry point on host arrays x [count,nx], u [count,nu]."""
    if not isinstance(model, Model):
        raise ValueError()  # reference bbdynamicswrap.pyx:53-54,127-128
    lib = _native.lib()
    _native.require_device()
    nx, nu = lib.dpilqr_model_nx(model.value), lib.dpilqr_model_nu(model.value)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, nx)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, nu)
    if x.shape[0] != u.shape[0]:
        raise ValueError("x and u must hold the same number of samples")
    count = x.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    xd, ud = torch.as_tensor(x).to(dev), torch.as_tensor(u).to(dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    if mode == "f":
        out = torch.empty((count, nx), dtype=torch.float64, device=dev)
        _native.check(lib.dpilqr_f(model.value, count, p(xd), p(ud), p(out), stream))
        return out.cpu().numpy()
    if mode == "integrate":
        out = torch.empty((count, nx), dtype=torch.float64, device=dev)
        _native.check(lib.dpilqr_integrate(model.value, float(dt), count, p(xd), p(ud), p(out), stream))
        return out.cpu().numpy()
    A = torch.empty((count, nx, nx), dtype=torch.float64, device=dev)
    B = torch.empty((count, nx, nu), dtype=torch.float64, device=dev)
    _native.check(lib.dpilqr_linearize(model.value, float(dt), count, p(xd), p(ud), p(A), p(B), stream))
    return A.cpu().numpy(), B.cpu().numpy()


def _check_contiguous(x, u):
    # reference bbdynamicswrap.pyx:57-58
    if not x.flags["C_CONTIGUOUS"] or not u.flags["C_CONTIGUOUS"]:
        raise ValueError("Contiguous")


def f(x, u, model):
    """Continuous derivative of one agent (reference bbdynamicswrap.pyx:61-90)."""
    if not isinstance(model, Model):
        raise ValueError()
    _check_contiguous(x, u)
    return _device_eval("f", model, 0.0, x, u)[0]


def integrate(x, u, dt, model):
    """One zero-order-hold RK4 step of one agent (reference bbdynamicswrap.pyx:93-123)."""
    if not isinstance(model, Model):
        raise ValueError()
    _check_contiguous(x, u)
    return _device_eval("integrate", model, dt, x, u)[0]


def linearize(x, u, dt, model):
    """Euler-discretised Jacobians (A, B) of one agent (reference bbdynamicswrap.pyx:125-164)."""
    A, B = _device_eval("linearize", model, dt, np.ascontiguousarray(x), np.ascontiguousarray(u))
    return A[0], B[0]


class DynamicalModel(abc.ABC):
    """Base class with the reference's id counter semantics (reference dynamics.py:54-92)."""

    _id = 0
    model = None  # native model id, set by the concrete classes

    def __init__(self, n_x, n_u, dt, id=None):
        if not id:
            id = DynamicalModel._id
            DynamicalModel._id += 1
        self.n_x = n_x
        self.n_u = n_u
        self.dt = dt
        self.id = id
        self.NX_EYE = np.eye(self.n_x, dtype=np.float32)

    def __call__(self, x, u):
        return _device_eval("integrate", self.model, self.dt, np.asarray(x, dtype=np.float64), np.asarray(u, dtype=np.float64))[0]

    def f(self, x, u):
        return _device_eval("f", self.model, 0.0, np.asarray(x, dtype=np.float64), np.asarray(u, dtype=np.float64))[0]

    def linearize(self, x, u):
        A, B = _device_eval("linearize", self.model, self.dt, np.asarray(x, dtype=np.float64), np.asarray(u, dtype=np.float64))
        return A[0], B[0]

    @classmethod
    def _reset_ids(cls):
        DynamicalModel._id = 0

    def __repr__(self):
        return f"{type(self).__name__}(n_x: {self.n_x}, n_u: {self.n_u}, id: {self.id})"


class CppModel(DynamicalModel):
    """Models the reference implements in its C++ library (reference dynamics.py:117-130)."""

    def __init__(self, dt, *args, **kwargs):
        super().__init__(dt, *args, **kwargs)


class SymbolicModel(DynamicalModel):
    """Kept for API compatibility (reference dynamics.py:95-114).  The one symbolic model the
    reference ships, Bike5D, has a native device implementation here, so no sympy is involved."""


class MultiDynamicalModel(DynamicalModel):
    """Concatenation of per-agent models with uniform strides (reference dynamics.py:133-202)."""

    def __init__(self, submodels):
        self.submodels = submodels
        self.n_players = len(submodels)
        self.x_dims = [sm.n_x for sm in submodels]
        self.u_dims = [sm.n_u for sm in submodels]
        self.ids = [sm.id for sm in submodels]
        super().__init__(sum(self.x_dims), sum(self.u_dims), submodels[0].dt, -1)

    def _per_model(self, mode, x, u):
        """Evaluate all agents, one device call per distinct model class."""
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        u = np.asarray(u, dtype=np.float64).reshape(-1)
        nx, nu = self.x_dims[0], self.u_dims[0]
        groups = {}
        for i, sm in enumerate(self.submodels):
            groups.setdefault(sm.model, []).append(i)
        out = [None] * self.n_players
        for model, idxs in groups.items():
            xs = np.stack([x[i * nx:(i + 1) * nx] for i in idxs])
            us = np.stack([u[i * nu:(i + 1) * nu] for i in idxs])
            res = _device_eval(mode, model, self.submodels[idxs[0]].dt, xs, us)
            for k, i in enumerate(idxs):
                out[i] = (res[0][k], res[1][k]) if mode == "linearize" else res[k]
        return out

    def f(self, x, u):
        return np.concatenate(self._per_model("f", x, u)).reshape(np.shape(x))

    def __call__(self, x, u):
        return np.concatenate(self._per_model("integrate", x, u)).reshape(np.shape(x))

    def linearize(self, x, u):
        subs = self._per_model("linearize", x, u)
        return uniform_block_diag(*[ab[0] for ab in subs]), uniform_block_diag(*[ab[1] for ab in subs])

    def split(self, graph):
        """One MultiDynamicalModel per graph entry, members kept in this model's order (reference dynamics.py:188-198)."""
        return [
            MultiDynamicalModel([sm for sm in self.submodels if sm.id in graph[problem]])
            for problem in graph
        ]

    def __repr__(self):
        inner = ",\n\t".join(repr(sm) for sm in self.submodels)
        return f"MultiDynamicalModel(\n\t{inner}\n)"


def _native_model(name, base, n_x, n_u, model, doc):
    def __init__(self, dt, *args, **kwargs):
        base.__init__(self, n_x, n_u, dt, *args, **kwargs)

    return type(name, (base,), {"__init__": __init__, "model": model, "__doc__": doc})


# reference dynamics.py:205-256
DoubleIntDynamics4D = _native_model("DoubleIntDynamics4D", CppModel, 4, 2, Model.DoubleInt4D, "x=[px,py,vx,vy], u=[ax,ay]")
DoubleIntDynamics6D = _native_model("DoubleIntDynamics6D", CppModel, 6, 3, Model.DoubleInt6D, "x=[p(3),v(3)], u=[a(3)]")
CarDynamics3D = _native_model("CarDynamics3D", CppModel, 3, 2, Model.Car3D, "x=[px,py,theta], u=[v,omega]")
UnicycleDynamics4D = _native_model("UnicycleDynamics4D", CppModel, 4, 2, Model.Unicycle4D, "x=[px,py,v,theta], u=[a,omega]")
QuadcopterDynamics6D = _native_model("QuadcopterDynamics6D", CppModel, 6, 3, Model.Quadcopter6D, "x=[p(3),v(3)], u=[tau,phi,theta]")
QuadcopterDynamics12D = _native_model("QuadcopterDynamics12D", CppModel, 12, 4, Model.Quadcopter12D, "x=[p,psi,theta,phi,v_body,w_body], u=[tau(3),f_z]")
HumanDynamics6D = _native_model("HumanDynamics6D", CppModel, 6, 3, Model.Human6D, "planar unicycle at constant height, heading as control")
HumanDynamicsLin6D = _native_model("HumanDynamicsLin6D", CppModel, 6, 3, Model.HumanLin6D, "planar double integrator at constant height")
# reference dynamics.py:254-277 (sympy there; single-step RK4 + Euler Jacobians on the device here)
BikeDynamics5D = _native_model("BikeDynamics5D", SymbolicModel, 5, 2, Model.Bike5D, "x=[px,py,v,theta,phi], u=[a,rho]")


def linearize_finite_difference(f, x, u):
    """Finite-difference Jacobians of a discrete step function (reference dynamics.py:281-290).
    Test helper only; delegates to SciPy."""
    from scipy.optimize import approx_fprime

    n_x = x.size
    eps = np.sqrt(np.finfo(float).eps)
    A = np.vstack([approx_fprime(x, lambda x: f(x, u)[i], eps) for i in range(n_x)])
    B = np.vstack([approx_fprime(u, lambda u: f(x, u)[i], eps) for i in range(n_x)])
    return A, B
