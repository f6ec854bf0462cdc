"""dpilqr_b200 -- B200-native batched iLQR hot path behind the dpilqr API.

The flat namespace mirrors reference dpilqr/__init__.py:1-58 (same names), plus the batched
front door (``CompiledBatch``, ``solve_specs``, ``solve_distributed_batch``, ...).
"""

from .dynamics import Model, f, integrate, linearize
from .control import RecedingHorizonController, ilqrSolver
from .cost import (
    Cost,
    GameCost,
    ProximityCost,
    ReferenceCost,
    quadraticize_distance,
    quadraticize_finite_difference,
)
from .distributed import (
    define_inter_graph_threshold,
    inter_graph_batch,
    solve_centralized,
    solve_distributed,
    solve_distributed_batch,
    solve_rhc,
)
from .dynamics import (
    BikeDynamics5D,
    CarDynamics3D,
    CppModel,
    DoubleIntDynamics4D,
    DoubleIntDynamics6D,
    DynamicalModel,
    HumanDynamics6D,
    HumanDynamicsLin6D,
    MultiDynamicalModel,
    QuadcopterDynamics6D,
    QuadcopterDynamics12D,
    SymbolicModel,
    UnicycleDynamics4D,
    linearize_finite_difference,
)
from .batched import LOG_HEADER, solve_distributed_round, solve_rhc_batch, trajectory_metrics
from .engine import CompiledBatch, ProblemSpec, SolvePipeline, bin_specs, raise_for_status, solve_specs, spec_from_problem
from .graphics import (
    eyeball_scenario,
    make_trajectory_gif,
    plot_interaction_graph,
    plot_pairwise_distances,
    plot_solve,
    set_bounds,
)
from .problem import _reset_ids, ilqrProblem, solve_subproblem
from .util import (
    Point,
    compute_energy,
    compute_pairwise_distance,
    compute_pairwise_distance_nd,
    distance_to_goal,
    normalize_energy,
    perturb_state,
    pos_mask,
    random_setup,
    randomize_locs,
    repopath,
    split_agents,
    split_agents_gen,
    split_graph,
    uniform_block_diag,
    π,
)

__version__ = "0.1.0"
