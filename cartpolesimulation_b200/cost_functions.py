"""CartPole cost-function plugins backed by CUDA kernels, with the reference's plugin surface:

  cost_function_base      Control_Toolkit/Cost_Functions/__init__.py:9-104
  CostFunctionWrapper     Control_Toolkit/Cost_Functions/cost_function_wrapper.py:16-115
  default                 Control_Toolkit_ASF/Cost_Functions/CartPole/default.py:19-88
  quadratic_boundary      .../quadratic_boundary.py:22-87
  quadratic_boundary_grad_minimal   .../quadratic_boundary_grad_minimal.py:17-140
  quadratic_boundary_grad           .../quadratic_boundary_grad.py:17-268

Inside `optimizer_mppi_b200` the cost is fused into the rollout kernel and these objects only carry the
configuration; the methods below serve every OTHER caller of the plugin API (the forward-only optimizers, logging).
"""
from __future__ import annotations

import numpy as np
import torch

from . import config as cfgmod
from .core import Engine
from .predictors import read_variable


class cost_function_base:
    supported_computation_libraries = ("Numpy", "TF", "Pytorch")
    MIN_COST = -1.0
    MAX_COST = 0.0
    COST_RANGE = MAX_COST - MIN_COST
    name: str = None

    def __init__(self, variable_parameters, ComputationLib=None, device=None, config: dict | None = None) -> None:
        self.variable_parameters = variable_parameters
        self.lib = ComputationLib
        self.batch_size = None
        self.horizon = None
        self.reload_cost_parameters_from_config_flag = False
        self.logged_attributes = {}
        self.config = cfgmod.cost_config(self.name) if config is None else dict(config)
        self._device = device
        self.engine = None
        self._var = None
        if self.name in ("default", "quadratic_boundary"):
            self.MAX_COST = float(cfgmod.max_cost(self.name, self.config))
            self.COST_RANGE = self.MAX_COST - self.MIN_COST

    def configure(self, batch_size: int, horizon: int):
        self.batch_size = batch_size
        self.horizon = horizon
        self.engine = Engine(num_rollouts=max(int(batch_size), 1), horizon=int(horizon), integrator="ODE",
                             cost=self.name, device=self._device)
        self.engine.set_cost_params(cfgmod.cost_vector(self.name, self.config))
        self._var = None

    # hot reload (cost_function_wrapper.py:71-74 polls the flag; the plugin re-reads its YAML block)
    def reload_cost_parameters_from_config(self):
        self.config = cfgmod.cost_config(self.name)
        if self.engine is not None:
            self.engine.set_cost_params(cfgmod.cost_vector(self.name, self.config))

    def set_computation_library(self, ComputationLib):
        self.lib = ComputationLib

    def set_logged_attributes(self, logged_attributes_dict):
        self.logged_attributes = dict(logged_attributes_dict)

    def _refresh(self):
        if self.engine is None:
            raise RuntimeError("cost function used before configure(batch_size, horizon)")
        var = (read_variable(self.variable_parameters, "target_position", 0.0),
               read_variable(self.variable_parameters, "target_equilibrium", 1.0))
        if var != self._var:
            self.engine.set_variable_parameters(target_position=var[0], target_equilibrium=var[1])
            self._var = var

    def _dev(self, x):
        dev = self.engine.device
        if isinstance(x, torch.Tensor):
            return x.detach().to(device=dev, dtype=torch.float32).contiguous(), x.device
        if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
            x = x.numpy()
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device=dev), None

    @staticmethod
    def _back(t, where):
        if where is None:
            return t.cpu().numpy()
        return t if where == t.device else t.to(where)

    @staticmethod
    def _u_prev(previous_input):
        if previous_input is None:
            return 0.0
        return float(np.asarray(previous_input.detach().cpu() if isinstance(previous_input, torch.Tensor)
                                else previous_input).reshape(-1)[0])

    def get_terminal_cost(self, terminal_states):
        self._refresh()
        s, where = self._dev(terminal_states)
        return self._back(self.engine.terminal_cost(s).reshape(-1, 1), where)

    def get_stage_cost(self, states, inputs, previous_input):
        """states [K,T,6] (the reference passes state_horizon[:, :-1, :]), inputs [K,T,1] -> [K,T]."""
        self._refresh()
        s, where = self._dev(states)
        q, _ = self._dev(inputs)
        return self._back(self.engine.stage_cost(s, q.reshape(q.shape[0], -1), self._u_prev(previous_input)), where)

    def _get_stage_cost(self, states, inputs, previous_input):
        self._refresh()
        s, where = self._dev(states)
        q, _ = self._dev(inputs)
        return self._back(self.engine.stage_cost(s, q.reshape(q.shape[0], -1), self._u_prev(previous_input),
                                                 unshifted=True), where)

    def get_summed_stage_cost(self, states, inputs, previous_input):
        st = self.get_stage_cost(states[:, :-1, :], inputs, previous_input)
        return st.sum(1)

    def get_trajectory_cost(self, state_horizon, inputs, previous_input=None):
        """[K,T+1,6], [K,T,1] -> [K]: mean over the T+1 entries (Cost_Functions/__init__.py:74-93)."""
        self._refresh()
        s, where = self._dev(state_horizon)
        q, _ = self._dev(inputs)
        return self._back(self.engine.trajectory_cost(s, q.reshape(q.shape[0], -1), self._u_prev(previous_input)),
                          where)


class default(cost_function_base):
    name = "default"


class quadratic_boundary(cost_function_base):
    name = "quadratic_boundary"


class quadratic_boundary_grad_minimal(cost_function_base):
    name = "quadratic_boundary_grad_minimal"


class quadratic_boundary_grad(cost_function_base):
    name = "quadratic_boundary_grad"


PLUGINS = {c.name: c for c in (default, quadratic_boundary, quadratic_boundary_grad_minimal, quadratic_boundary_grad)}


class CostFunctionWrapper:
    def __init__(self, cost_function_name_default: str = "default"):
        self.cost_function = None
        self.cost_function_name_default = cost_function_name_default
        self.cost_function_name = None

    def configure(self, batch_size: int, horizon: int, variable_parameters, environment_name: str = "CartPole",
                  computation_library=None, cost_function_specification: str = None, device=None, config=None):
        self.batch_size = batch_size
        self.horizon = horizon
        self.variable_parameters = variable_parameters
        self.environment_name = environment_name
        self.computation_library = computation_library
        self.cost_function_specification = cost_function_specification
        self.update_cost_function_name_from_specification(cost_function_specification)
        if environment_name != "CartPole":
            raise NotImplementedError(f"only the CartPole cost plugins are implemented, not {environment_name}")
        if self.cost_function_name not in PLUGINS:
            raise ValueError(f"cost function {self.cost_function_name} is not available "
                             f"(supported: {sorted(PLUGINS)})")
        self.cost_function = PLUGINS[self.cost_function_name](variable_parameters, computation_library, device=device,
                                                              config=config)
        self.cost_function.configure(batch_size=batch_size, horizon=horizon)

    def update_cost_parameters_from_config(self):
        if self.cost_function.reload_cost_parameters_from_config_flag:
            self.cost_function.reload_cost_parameters_from_config()
            self.cost_function.reload_cost_parameters_from_config_flag = False

    def update_cost_function_name_from_specification(self, cost_function_specification: str = None):
        if cost_function_specification is None:
            self.cost_function_name = self.cost_function_name_default.replace("-", "_")
        elif isinstance(cost_function_specification, str):
            self.cost_function_name = cost_function_specification.replace("-", "_")
        else:
            raise ValueError(f"Cannot interpret cost function specification {cost_function_specification}.")

    def get_terminal_cost(self, terminal_states):
        return self.cost_function.get_terminal_cost(terminal_states)

    def get_stage_cost(self, states, inputs, previous_input):
        return self.cost_function.get_stage_cost(states, inputs, previous_input)

    def get_trajectory_cost(self, state_horizon, inputs, previous_input=None):
        return self.cost_function.get_trajectory_cost(state_horizon, inputs, previous_input)

    def get_summed_stage_cost(self, state_horizon, inputs, previous_input=None):
        return self.cost_function.get_summed_stage_cost(state_horizon, inputs, previous_input)

    def copy(self):
        c = CostFunctionWrapper(self.cost_function_name_default)
        c.cost_function_name = self.cost_function_name
        return c
