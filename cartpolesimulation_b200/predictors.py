"""Predictor plugins backed by the CUDA rollout kernel -- same names, arguments and error behaviour as the
reference's predictor interface:

  PredictorWrapper        SI_Toolkit/src/SI_Toolkit/Predictors/predictor_wrapper.py:14-191
  predictor_ODE_v0        SI_Toolkit/src/SI_Toolkit/Predictors/predictor_ODE_v0.py:18-77   (explicit Euler + bounce)
  predictor_ODE           SI_Toolkit/src/SI_Toolkit/Predictors/predictor_ODE.py:23-101     (Euler-Cromer + atan2)
  template_predictor      SI_Toolkit/src/SI_Toolkit/Predictors/__init__.py:6-39

`predict_core(s[B,6], Q[B,T,1]) -> [B,T+1,6]` is what all 15 reference optimizers call
(e.g. Control_Toolkit/Optimizers/optimizer_mppi.py:187).  Tensors may be torch CUDA tensors (zero copy), torch CPU
tensors or numpy arrays (copied up and back); the arithmetic always runs on the GPU -- there is no CPU path.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from . import _lib as L
from . import config as cfgmod
from .core import Engine

STATE_VARIABLES = np.sort(["angle", "angleD", "angle_cos", "angle_sin", "position", "positionD"])
STATE_INDICES = {x: int(np.where(STATE_VARIABLES == x)[0][0]) for x in STATE_VARIABLES}
CONTROL_INPUTS = np.sort(["Q"])


def read_variable(vp, name, default):
    """variable_parameters attributes are 0-d tensors / numpy scalars / floats (General/variable_parameters.py:6-29)."""
    if vp is None or not hasattr(vp, name):
        return float(default)
    v = getattr(vp, name)
    try:
        return float(v)
    except Exception:
        return float(np.asarray(v).reshape(-1)[0])


class template_predictor:
    supported_computation_libraries = ("Numpy", "TF", "Pytorch")

    def __init__(self, horizon: int, batch_size: int) -> None:
        self.horizon = horizon
        self.batch_size = batch_size
        self.predictor_initial_input_features = STATE_VARIABLES
        self.predictor_external_input_features = CONTROL_INPUTS
        self.predictor_output_features = STATE_VARIABLES
        self.num_states = len(STATE_VARIABLES)
        self.num_control_inputs = len(CONTROL_INPUTS)

    def predict_core(self, s, Q):
        raise NotImplementedError()

    def predict(self, s, Q):
        raise NotImplementedError()


class _ode_predictor_base(template_predictor):
    integrator = None

    def __init__(self, horizon: int, dt: float, intermediate_steps: int = 10, batch_size: int = 1,
                 variable_parameters=None, device=None, fast_sincos=False, exact_atan2=False, **kwargs):
        super().__init__(horizon=horizon, batch_size=batch_size)
        self.dt = dt
        self.intermediate_steps = int(intermediate_steps)
        self.variable_parameters = variable_parameters
        self.engine = Engine(num_rollouts=max(int(batch_size), 1), horizon=int(horizon), dt=float(dt),
                             substeps=self.intermediate_steps, integrator=self.integrator, cost=None, device=device,
                             fast_sincos=fast_sincos, exact_atan2=exact_atan2)
        self.device = self.engine.device
        self._var = None
        self.output = None
        self.initial_state = None

    def _refresh_variable_parameters(self):
        var = (read_variable(self.variable_parameters, "L", cfgmod.DEFAULT_PHYSICS["L"]),
               read_variable(self.variable_parameters, "m_pole", cfgmod.DEFAULT_PHYSICS["m_pole"]))
        if var != self._var:
            self.engine.set_variable_parameters(L=var[0], m_pole=var[1])
            self._var = var

    @staticmethod
    def _to_dev(x, device):
        if isinstance(x, torch.Tensor):
            return x.detach().to(device=device, dtype=torch.float32).contiguous(), ("torch", x.device)
        if isinstance(x, np.ndarray) or np.isscalar(x) or isinstance(x, (list, tuple)):
            return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device=device), ("numpy", None)
        if hasattr(x, "numpy"):  # e.g. an eager TF tensor
            return torch.as_tensor(np.ascontiguousarray(x.numpy(), dtype=np.float32), device=device), ("numpy", None)
        raise ValueError(f"unsupported tensor type {type(x)}")

    def predict_core(self, s, Q):
        """s: [B,6] (or [1,6] / [6]: tiled), Q: [B,T,1] -> [B,T+1,6], row 0 = s."""
        self._refresh_variable_parameters()
        s_d, kind = self._to_dev(s, self.device)
        Q_d, _ = self._to_dev(Q, self.device)
        if Q_d.ndim != 3 or Q_d.shape[2] != 1:
            raise ValueError(f"Q must have shape [batch_size, horizon, 1], got {tuple(Q_d.shape)}")
        traj, _ = self.engine.rollout(s_d, Q_d[:, :, 0], q_layout=L.ROLLOUT_MAJOR, traj_layout=L.ROLLOUT_MAJOR)
        self.output = traj
        if kind[0] == "torch":
            return traj if kind[1] == self.device else traj.to(kind[1])
        return traj.cpu().numpy()

    def update_internal_state(self, Q0=None, s=None):
        pass  # ODE predictors are stateless (predictor_ODE.py:99-100, predictor_ODE_v0.py:76-77)


class predictor_ODE_v0(_ode_predictor_base):
    """numpy in / numpy out `predict`, tolerant of Q.ndim 1/2/3 and a single initial state; returns a squeezed
    array when batch_size == 1 (predictor_ODE_v0.py:42-74)."""
    integrator = "ODE_v0"
    supported_computation_libraries = ("Numpy",)

    def predict(self, initial_state: np.ndarray, Q: np.ndarray, params=None) -> np.ndarray:
        initial_state = np.asarray(initial_state, dtype=np.float32)
        Q = np.asarray(Q, dtype=np.float32)
        if Q.ndim == 3:
            self.batch_size = Q.shape[0]
        elif Q.ndim == 2:
            self.batch_size = 1
            Q = Q[np.newaxis, :, :]
        elif Q.ndim == 1:
            self.batch_size = 1
            Q = Q[np.newaxis, np.newaxis, :]
        else:
            raise ValueError()
        if initial_state.ndim == 1:
            initial_state = initial_state[np.newaxis, :]
        if initial_state.shape[0] == 1 and Q.shape[0] != 1:
            pass  # tiled on the device
        elif initial_state.shape[0] == Q.shape[0]:
            pass
        else:
            raise ValueError('Batch size of control input contradict batch size of initial state')
        self.initial_state = initial_state
        out = self.predict_core(initial_state, Q[:, :self.horizon, :])
        self.output = out
        return out if (self.batch_size > 1) else np.squeeze(out)


class predictor_ODE(_ode_predictor_base):
    integrator = "ODE"

    def predict(self, initial_state, Q):
        """predictor_ODE.predict (predictor_ODE.py:71-83): check_dimensions then predict_core, numpy out."""
        initial_state = np.asarray(initial_state, dtype=np.float32)
        Q = np.asarray(Q, dtype=np.float32)
        if initial_state.ndim == 1:
            initial_state = initial_state[np.newaxis, :]
        if Q.ndim == 2:
            Q = Q[np.newaxis, :, :]
        elif Q.ndim == 1:
            Q = Q[np.newaxis, np.newaxis, :]
        self.initial_state = initial_state
        self.batch_size = Q.shape[0]
        return self.predict_core(initial_state, Q)


NETWORK_NAMES = ['Dense', 'RNN', 'GRU', 'DeltaGRU', 'LSTM', 'Custom']


class PredictorWrapper:
    """Deferred-configuration wrapper with the reference's surface (predictor_wrapper.py:14-191)."""

    def __init__(self, predictors_config: dict | None = None, predictor_name_default: str = "ODE_default"):
        self.horizon = None
        self.batch_size = None
        self.num_states = None
        self.num_control_inputs = None
        self.predictor = None
        self.predictors_config = copy.deepcopy(predictors_config or cfgmod.DEFAULT_PREDICTOR_CONFIG)
        self.predictor_name_default = predictor_name_default
        self.predictor_name = predictor_name_default
        self.predictor_config = copy.deepcopy(self.predictors_config[self.predictor_name])
        self.predictor_type = self.predictor_config['predictor_type']
        self.model_name = self.predictor_config.get('model_name')

    def configure(self, batch_size: int, horizon: int, dt: float, computation_library=None, variable_parameters=None,
                  predictor_specification=None, compile_standalone=False, mode=None, hls=False, **kwargs):
        self.update_predictor_config_from_specification(predictor_specification)
        self.batch_size = batch_size
        self.horizon = horizon
        cfg = {k: v for k, v in self.predictor_config.items() if k not in ("predictor_type", "model_name",
                                                                         "computation_library_name")}
        if self.predictor_type == 'ODE_v0':
            self.predictor = predictor_ODE_v0(horizon=horizon, dt=dt, batch_size=batch_size,
                                              variable_parameters=variable_parameters, **cfg, **kwargs)
        elif self.predictor_type == 'ODE':
            self.predictor = predictor_ODE(horizon=horizon, dt=dt, batch_size=batch_size,
                                           variable_parameters=variable_parameters, **cfg, **kwargs)
        elif self.predictor_type == 'neural':
            from .neural import predictor_autoregressive_neural
            self.predictor = predictor_autoregressive_neural(horizon=horizon, dt=dt, batch_size=batch_size,
                                                             variable_parameters=variable_parameters,
                                                             model_name=self.model_name, **cfg, **kwargs)
        elif self.predictor_type == 'GP':
            raise NotImplementedError('GP predictors are outside the B200 hot path (they need gpflow/TF)')
        else:
            raise NotImplementedError('Type of the predictor not recognised.')
        self.num_states = self.predictor.num_states
        self.num_control_inputs = self.predictor.num_control_inputs

    def configure_with_compilation(self, batch_size, horizon, dt, predictor_specification=None, mode=None, hls=False):
        self.configure(batch_size, horizon, dt, predictor_specification=predictor_specification)

    def update_predictor_config_from_specification(self, predictor_specification: str = None):
        if predictor_specification is None:
            return
        comps = predictor_specification.split(":")
        predictor_name = {"ODE": "ODE_default", "ODE_v0": "ODE_v0_default", "neural": "neural_default",
                          "GP": "GP_default"}.get(comps[0])
        model_name = None
        if predictor_name == "GP_default" and predictor_name not in self.predictors_config:
            raise NotImplementedError('GP predictors are outside the B200 hot path (they need gpflow/TF)')
        if predictor_name is None and comps[0] in self.predictors_config:
            predictor_name = comps[0]
        if predictor_name is None:
            if any(n in predictor_specification for n in NETWORK_NAMES):
                predictor_name = 'neural_default'
                model_name = comps[0]
            elif 'SGP' in predictor_specification:
                raise NotImplementedError('GP predictors are outside the B200 hot path')
        if predictor_name is None or predictor_name not in self.predictors_config:
            raise ValueError('{} is an invalid predictor specification'.format(predictor_specification))
        if len(comps) > 1 and model_name is None:
            model_name = comps[1]
        self.predictor_name = predictor_name
        self.predictor_config = copy.deepcopy(self.predictors_config[predictor_name])
        self.predictor_type = self.predictor_config['predictor_type']
        if model_name is not None:
            self.predictor_config['model_name'] = model_name
        self.model_name = self.predictor_config.get('model_name')

    def predict(self, s, Q):
        return self.predictor.predict(s, Q)

    def predict_core(self, s, Q):
        return self.predictor.predict_core(s, Q)

    def update(self, Q0, s):
        if self.predictor_type == 'neural':
            self.predictor.update_internal_state_tf(s=s, Q0=Q0)

    def copy(self):
        c = PredictorWrapper(self.predictors_config, self.predictor_name_default)
        c.predictor_name = self.predictor_name
        c.predictor_config = copy.deepcopy(self.predictor_config)
        c.predictor_type = self.predictor_type
        c.model_name = self.model_name
        return c
