"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the API surface matches the reference's flat namespace, host logic behaves, and
the product fails loudly (no CPU fallback) when no GPU is present."""

import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# names re-exported by reference dpilqr/__init__.py:1-58
REFERENCE_API = """Model f integrate linearize RecedingHorizonController ilqrSolver Cost GameCost ProximityCost ReferenceCost
quadraticize_distance quadraticize_finite_difference define_inter_graph_threshold solve_centralized solve_distributed solve_rhc
BikeDynamics5D CarDynamics3D DoubleIntDynamics4D DoubleIntDynamics6D DynamicalModel HumanDynamics6D HumanDynamicsLin6D
MultiDynamicalModel QuadcopterDynamics6D QuadcopterDynamics12D SymbolicModel UnicycleDynamics4D linearize_finite_difference
eyeball_scenario make_trajectory_gif plot_interaction_graph plot_pairwise_distances plot_solve set_bounds _reset_ids ilqrProblem
Point compute_energy compute_pairwise_distance compute_pairwise_distance_nd distance_to_goal normalize_energy perturb_state
pos_mask random_setup randomize_locs repopath split_agents split_agents_gen split_graph uniform_block_diag π""".split()


@pytest.fixture(scope="module")
def built_lib():
    from dpilqr_b200 import build

    return build.build()


def test_header_symbols_are_exported(built_lib):
    import ctypes

    header = open(os.path.join(ROOT, "include", "dpilqr_b200.h")).read()
    declared = set(re.findall(r"\b(dpilqr_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 21
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    from dpilqr_b200 import _native

    assert set(_native.EXPORTS) == declared
    assert _native.lib().dpilqr_version() >= 100
    assert _native.lib().dpilqr_model_nx(7) == 12 and _native.lib().dpilqr_model_nu(7) == 4
    assert _native.lib().dpilqr_stage_stride(10, 12, 4) == 2616
    assert _native.lib().dpilqr_workspace_bytes(4096, 10, 12, 4, 50, 10) > 0


def test_struct_layout_matches_header(built_lib):
    import ctypes

    from dpilqr_b200 import _native

    assert ctypes.sizeof(_native.BatchStruct) == 6 * 4 + 8 + 10 * 8 + 2 * 4
    assert ctypes.sizeof(_native.SolveOpts) == 40


def test_api_surface_matches_reference_namespace():
    import dpilqr
    import dpilqr_b200 as dp

    for name in REFERENCE_API:
        assert hasattr(dp, name), name
        assert hasattr(dpilqr, name), name
    assert [m.name for m in dp.Model][:8] == ["DoubleInt4D", "DoubleInt6D", "Car3D", "Unicycle4D", "Quadcopter6D", "Human6D",
                                              "HumanLin6D", "Quadcopter12D"]
    assert dp.ilqrSolver.DELTA_0 == 2.0 and dp.ilqrSolver.MU_MIN == 1e-6 and dp.ilqrSolver.N_LS_ITER == 10


def test_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import dpilqr_b200 as dp
    from dpilqr_b200._native import NativeError

    with pytest.raises(NativeError, match="no CPU fallback"):
        dp.integrate(np.zeros(4), np.zeros(2), 0.1, dp.Model.DoubleInt4D)
    with pytest.raises(NativeError):
        dp.define_inter_graph_threshold(np.zeros((1, 8)), 0.5, [4, 4], [0, 1])


def _toy_problem(dp, a=3):
    dp._reset_ids()
    ids = [100 + i for i in range(a)]
    dyn = dp.MultiDynamicalModel([dp.UnicycleDynamics4D(0.1, id_) for id_ in ids])
    xf = np.arange(4.0 * a)
    costs = [dp.ReferenceCost(xf[4 * i:4 * i + 4], np.eye(4), np.eye(2), 10 * np.eye(4), id_) for i, id_ in enumerate(ids)]
    return dp.ilqrProblem(dyn, dp.GameCost(costs, dp.ProximityCost([4] * a, 0.5, [2] * a)))


def test_problem_compiler_and_split_semantics():
    import dpilqr_b200 as dp

    prob = _toy_problem(dp)
    spec = dp.spec_from_problem(prob)
    assert (spec.a, spec.s, spec.c, spec.dt) == (3, 4, 2, 0.1)
    assert spec.models == [3, 3, 3] and spec.has_prox and spec.weights == (1.0, 200.0)
    assert spec.ids == [100, 101, 102] == prob.ids
    graph = {100: [100, 102], 101: [101], 102: [100, 102]}
    subs = prob.split(graph)
    assert [p.ids for p in subs] == [[100, 102], [101], [100, 102]]
    assert subs[0].game_cost.prox_cost.n_dims == [2, 2]
    sub = spec.subset([0, 2])
    assert sub.ids == [100, 102] and np.array_equal(sub.xf, np.r_[spec.xf[:4], spec.xf[8:]])
    X, U = np.arange(22.0).reshape(1, 22)[:, :12].repeat(3, 0), np.zeros((2, 6))
    Xi, Ui = prob.extract(X, U, 101)
    assert np.array_equal(Xi, X[:, 4:8])
    with pytest.raises(IndexError):
        prob.extract(X, U, 7)
    # ids counters: falsy ids draw from the class counters (reference dynamics.py:57-62, cost.py:42-51)
    dp._reset_ids()
    assert dp.UnicycleDynamics4D(0.1).id == 0 and dp.UnicycleDynamics4D(0.1).id == 1
    assert dp.ReferenceCost(np.zeros(4), np.eye(4), np.eye(2)).id == 0
    bins = dp.bin_specs([spec, sub, spec])
    assert sorted(len(v) for v in bins.values()) == [1, 2]


def test_compiler_rejects_what_the_kernels_cannot_run():
    import dpilqr_b200 as dp

    prob = _toy_problem(dp)

    class Custom(dp.Cost):
        def __call__(self, *a):
            return 0.0

        def quadraticize(self):
            pass

    with pytest.raises(TypeError):
        dp.spec_from_problem(dp.ilqrProblem(prob.dynamics, Custom()))
    mixed = dp.MultiDynamicalModel([dp.UnicycleDynamics4D(0.1, 5), dp.QuadcopterDynamics6D(0.1, 6)])
    with pytest.raises(ValueError):
        dp.spec_from_problem(dp.ilqrProblem(mixed, prob.game_cost))


def test_solve_rhc_argument_check():
    import dpilqr_b200 as dp

    prob = _toy_problem(dp)
    with pytest.raises(ValueError):
        dp.solve_rhc(prob, np.zeros(12), 10)
    with pytest.raises(ValueError):
        dp.solve_rhc(prob, np.zeros(12), 10, J_converge=1.0, dist_converge=0.1)


def test_scenario_generation_matches_reference_golden():
    """random_setup consumes the global RNG exactly like the reference: x0/xf of the golden cases are reproduced."""
    import random

    import dpilqr_b200 as dp
    from helpers import golden

    case = golden("solve_quad12_a10_s1.npz")
    np.random.seed(1)
    random.seed(1)
    x0, xf = dp.random_setup(10, 12, is_rotation=False, rel_dist=10, var=5.0, n_d=3, random=True, energy=30.0)
    assert np.array_equal(x0.flatten(), case["x0"]) and np.array_equal(xf.flatten(), case["xf"])
