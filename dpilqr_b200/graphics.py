"""Plotting helpers are outside the GPU hot path (reference dpilqr/graphics.py); the names are
kept importable so ``from dpilqr import *`` style code loads, and they need matplotlib."""


def _needs_matplotlib(name):
    def stub(*args, **kwargs):
        raise NotImplementedError(f"{name} is a plotting helper of the reference and is not part of dpilqr_b200")

    stub.__name__ = name
    return stub


eyeball_scenario = _needs_matplotlib("eyeball_scenario")
make_trajectory_gif = _needs_matplotlib("make_trajectory_gif")
plot_interaction_graph = _needs_matplotlib("plot_interaction_graph")
plot_pairwise_distances = _needs_matplotlib("plot_pairwise_distances")
plot_solve = _needs_matplotlib("plot_solve")
set_bounds = _needs_matplotlib("set_bounds")
