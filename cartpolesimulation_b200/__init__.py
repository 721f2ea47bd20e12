"""cartpolesimulation_b200 -- B200-native (sm_100a) implementation of CartPoleSimulation's MPPI rollout hot path.

Host side: Python mirrors of the reference's plugin interfaces (optimizer / predictor / cost function) over a
C ABI (include/cps.h, libcps_b200.so) into hand-written CUDA kernels.  No CPU fallback, no multi-backend
dispatch: every entry point raises if the CUDA library or a CUDA device is missing.
"""
from . import _lib
from ._lib import CpsError, build, library_path
from .config import DEFAULT_COST_CONFIG, DEFAULT_MPPI_CONFIG, DEFAULT_PHYSICS, cost_vector, physics_vector

__all__ = ["Engine", "optimizer_mppi_b200", "optimizer_cem_b200", "optimizer_cem_gmm_b200", "optimizer_rpgd_b200", "optimizer_random_action_b200", "PredictorWrapper", "predictor_ODE", "predictor_ODE_v0",
           "CostFunctionWrapper", "CpsError", "build", "library_path", "VariableParameters"]


class VariableParameters:
    """Plain attribute bag standing in for SI_Toolkit/General/variable_parameters.py:6-29
    (target_position, target_equilibrium, L, m_pole)."""

    def __init__(self, **attrs):
        self.set_attributes(attrs)

    def set_attributes(self, attrs, **kwargs):
        for k, v in attrs.items():
            setattr(self, k, v)

    def update_attributes(self, attrs):
        self.set_attributes(attrs)


def __getattr__(name):  # lazy: importing the package must not need torch.cuda or the built library
    if name == "Engine":
        from .core import Engine
        return Engine
    if name == "optimizer_mppi_b200":
        from .optimizer_mppi_b200 import optimizer_mppi_b200
        return optimizer_mppi_b200
    if name in ("optimizer_cem_b200", "optimizer_cem_gmm_b200", "optimizer_rpgd_b200", "optimizer_random_action_b200"):
        from . import optimizer_forward_b200
        return getattr(optimizer_forward_b200, name)
    if name in ("PredictorWrapper", "predictor_ODE", "predictor_ODE_v0"):
        from . import predictors
        return getattr(predictors, name)
    if name == "CostFunctionWrapper":
        from .cost_functions import CostFunctionWrapper
        return CostFunctionWrapper
    raise AttributeError(name)
