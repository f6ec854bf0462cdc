"""Synthetic benchmark scenarios (the Monte-Carlo shape of reference scripts/analysis.py:35-79,
with the inputs fixed in SURVEY.md section 8d / BASELINE.md section 3).

Scenario ``k`` with ``a`` Quadcopter12D agents: ``np.random.seed(k); random.seed(k)``,
``random_setup(a, 12, rel_dist=a, var=a/2, n_d=3, random=True, energy=3a)``, ids 100+i,
dt 0.1, N 50, Q = I, R = I, Qf = 1000 I, radius 0.5, hover warm start.
"""

import ctypes
import random

import numpy as np

from .engine import ProblemSpec
from .util import random_setup

G = 9.80665
HOVER_THRUST = G * 63 / 2000
QUAD12 = 7


def quad12_inputs(k, a, N=50):
    """(x0 [n], xf [n], U0 [N, m]) of scenario k -- host NumPy, seeded like the reference harness."""
    np.random.seed(k)
    random.seed(k)
    x0, xf = random_setup(a, 12, is_rotation=False, rel_dist=a, var=a / 2, n_d=3, random=True, energy=3.0 * a)
    U0 = np.tile([0.0, 0.0, 0.0, HOVER_THRUST], (N, a))
    return x0.reshape(-1), xf.reshape(-1), U0


_Q, _R, _QF = np.eye(12), np.eye(4), 1000.0 * np.eye(12)


def quad12_spec(xf, a, radius=0.5, dt=0.1):
    return ProblemSpec([QUAD12] * a, dt, 12, 4, [3] * a, [_Q] * a, [_R] * a, [_QF] * a, xf, radius, (1.0, 200.0), True,
                       [100 + i for i in range(a)])


def quad12_batch(first, count, a, N=50):
    """specs, x0 [count, n], U0 [count, N, m] for scenarios first .. first+count-1."""
    specs, x0s, U0s = [], [], []
    for k in range(first, first + count):
        x0, xf, U0 = quad12_inputs(k, a, N)
        specs.append(quad12_spec(xf, a))
        x0s.append(x0)
        U0s.append(U0)
    return specs, np.stack(x0s), np.stack(U0s)


def random_setup_batch(first, count, n_agents, n_states, n_d=2, var=3.0, energy=None, device=None):
    """``random_setup(n_agents, n_states, n_d=n_d, random=True, var=var, energy=energy)`` (reference util.py:165-195) for
    the seeds ``first .. first+count-1`` in one kernel launch, each scenario as if preceded by ``np.random.seed(seed)``:
    (x0, xf) as CUDA tensors [count, n_agents*n_states], bit-identical to the host path (the kernel runs NumPy's legacy
    MT19937 stream and NumPy's reduction orders).  ``random=False`` (mutual repulsion), ``is_rotation`` and ``do_face``
    stay on the host (``util.random_setup``)."""
    import torch

    from . import _native
    from .engine import default_device

    _native.require_device()
    dev = torch.device(device) if device is not None else default_device()
    x0 = torch.empty((count, n_agents * n_states), dtype=torch.float64, device=dev)
    xf = torch.empty_like(x0)
    with torch.cuda.device(dev):
        _native.check(_native.lib().dpilqr_random_setup(
            int(first), int(count), int(n_agents), int(n_states), int(n_d), float(var), float(energy or 0.0),
            ctypes.c_void_p(x0.data_ptr()), ctypes.c_void_p(xf.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return x0, xf


def quad12_batch_device(first, count, a, N=50, device=None, radius=0.5, dt=0.1):
    """The metric batch built on the device: (CompiledBatch, x0 [count, n], U0 [count, N, m]) for scenarios
    first .. first+count-1, without a per-scenario host loop -- same inputs, bit for bit, as :func:`quad12_batch`."""
    import torch

    from .engine import CompiledBatch, default_device

    dev = torch.device(device) if device is not None else default_device()
    x0, xf = random_setup_batch(first, count, a, 12, n_d=3, var=a / 2, energy=3.0 * a, device=dev)
    i32, f64 = dict(dtype=torch.int32, device=dev), dict(dtype=torch.float64, device=dev)
    eye12, eye4 = torch.eye(12, **f64), torch.eye(4, **f64)
    batch = CompiledBatch.from_tensors(
        N, a, 12, 4, dt, torch.full((count, a), QUAD12, **i32), torch.full((count, a), 3, **i32), torch.zeros((count, a), **i32),
        eye12[None].contiguous(), eye4[None].contiguous(), (1000.0 * eye12)[None].contiguous(), xf,
        torch.full((count,), radius, **f64), torch.tensor([1.0, 200.0], **f64).repeat(count, 1), torch.ones(count, **i32),
        model_hint=QUAD12 + 1, costs_nonnegative=True, device=dev)
    U0 = torch.tensor([0.0, 0.0, 0.0, HOVER_THRUST], **f64).repeat(count, N, a)
    return batch, x0, U0
