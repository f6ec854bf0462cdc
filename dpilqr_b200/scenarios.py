"""Synthetic benchmark scenarios (the Monte-Carlo shape of reference scripts/analysis.py:35-79,
with the inputs fixed in SURVEY.md section 8d / BASELINE.md section 3).

Scenario ``k`` with ``a`` Quadcopter12D agents: ``np.random.seed(k); random.seed(k)``,
``random_setup(a, 12, rel_dist=a, var=a/2, n_d=3, random=True, energy=3a)``, ids 100+i,
dt 0.1, N 50, Q = I, R = I, Qf = 1000 I, radius 0.5, hover warm start.
"""

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
