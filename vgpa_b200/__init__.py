"""
vgpa_b200 -- B200 (sm_100a) implementation of VGPA's variational free-energy and
gradient evaluation, behind the reference's own Python interface.

    from vgpa_b200 import VarGP, SCG, Simulation, BatchEvaluator

The first attribute access loads libvgpa_b200.so and fails loudly if the library
has not been built (python -m vgpa_b200.build).  There is no CPU fallback.
"""
import importlib

_EXPORTS = {
    "BatchEvaluator": "engine", "model_energy": "engine", "obs_energy": "engine",
    "solve_bwd": "engine", "solve_fwd": "engine",
    "make_trajectories": "engine", "collect_observations": "engine",
    "StochasticProcess": "dynamics", "DoubleWell": "dynamics", "OrnsteinUhlenbeck": "dynamics",
    "Lorenz63": "dynamics", "Lorenz96": "dynamics", "dynamical_systems": "dynamics",
    "FwdOde": "ode", "BwdOde": "ode", "GaussianLikelihood": "likelihood", "PriorKL0": "prior",
    "VarGP": "variational", "SCG": "scg", "Simulation": "simulation",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    # lazy so that `python -m vgpa_b200.build` can run before the library exists
    if name in _EXPORTS:
        return getattr(importlib.import_module(f".{_EXPORTS[name]}", __name__), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
