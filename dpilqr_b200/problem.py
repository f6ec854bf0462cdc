"""ilqrProblem: dynamics + cost in one object (reference dpilqr/problem.py)."""

from time import perf_counter as pc

import numpy as np

from .cost import GameCost, ReferenceCost
from .dynamics import DynamicalModel, MultiDynamicalModel
from .util import split_agents_gen


class ilqrProblem:
    """Centralized optimal control problem (reference problem.py:15-94)."""

    def __init__(self, dynamics, cost):
        self.dynamics = dynamics
        self.game_cost = cost
        self.n_agents = 1
        if isinstance(cost, GameCost):
            self.n_agents = len(cost.ref_costs)

    @property
    def ids(self):
        if not isinstance(self.dynamics, MultiDynamicalModel):
            raise NotImplementedError("Only MultiDynamicalModel's have an 'ids' attribute")
        if not self.dynamics.ids == self.game_cost.ids:
            raise ValueError(f"Dynamics and cost have inconsistent ID's: {self}")
        return self.dynamics.ids.copy()

    def split(self, graph):
        """Sub-problems dictated by the interaction graph (reference problem.py:36-47)."""
        return [ilqrProblem(d, c) for d, c in zip(self.dynamics.split(graph), self.game_cost.split(graph))]

    def extract(self, X, U, id_):
        """Columns of agent ``id_`` in this problem's joint trajectory (reference problem.py:49-64)."""
        if id_ not in self.ids:
            raise IndexError(f"Index {id_} not in ids: {self.ids}.")
        k = self.ids.index(id_)
        s, c = self.game_cost.x_dims[0], self.game_cost.u_dims[0]
        return X[:, k * s:(k + 1) * s], U[:, k * c:(k + 1) * c]

    def selfish_warmstart(self, x0, N):
        """Warm start that ignores the other agents (reference problem.py:66-91): all the
        single-agent solves run as ONE batch on the GPU."""
        from .engine import solve_specs, spec_from_problem

        t0 = pc()
        x0 = np.asarray(x0, dtype=np.float64).reshape(-1, 1)
        subproblems = self.split({id_: [id_] for id_ in self.ids})
        specs = [spec_from_problem(sub) for sub in subproblems]
        x0s = [xi for xi in split_agents_gen(x0, self.game_cost.x_dims)]
        U0s = [np.zeros((N, sub.dynamics.n_u)) for sub in subproblems]
        results = solve_specs(specs, x0s, U0s, N)
        U_warm = np.zeros((N, self.dynamics.n_u))
        for sub, id_, res in zip(subproblems, self.ids, results):
            nu_i = sub.dynamics.n_u
            k = self.ids.index(id_)
            U_warm[:, nu_i * k:nu_i * (k + 1)] = res["U"]
        print(f"All: {self.ids}\nTook {pc() - t0} seconds\n" + "=" * 80)
        return U_warm

    def __repr__(self):
        return f"ilqrProblem(\n\t{self.dynamics},\n\t{self.game_cost}\n)"


def solve_subproblem(args, **kwargs):
    """Solve one sub-problem and keep only agent ``id_``'s columns (reference problem.py:97-105)."""
    from .control import ilqrSolver

    subproblem, x0, U, id_, verbose = args
    subsolver = ilqrSolver(subproblem, U.shape[0])
    Xi, Ui, _ = subsolver.solve(x0, U, verbose=verbose, **kwargs)
    return *subproblem.extract(Xi, Ui, id_), id_


def solve_subproblem_starmap(subproblem, x0, U, id_):
    return solve_subproblem((subproblem, x0, U, id_))


def _reset_ids():
    """Reset the model / cost id counters (reference problem.py:113-116)."""
    DynamicalModel._reset_ids()
    ReferenceCost._reset_ids()
