"""controller_mppi_cartpole_b200 -- the legacy MPPI controller on the B200 rollout path.

Mirror of Control_Toolkit_ASF/Controllers/controller_mppi_cartpole.py:338-569 (`configure`, `step`,
`initialize_perturbations`, `update_control_vector`, `controller_reset`; attributes `u`, `u_prev`, `delta_u`,
`S_tilde_k`, `rng_mppi`, `rng_mppi_rnn`, `iteration`).  The reference keeps its parameters in module globals that the
GUI options window rewrites while the controller runs (GUI/_ControllerGUI_MPPIOptionsWindow.py:29-46); here they are
instance attributes with the same names (`mpc_horizon`, `num_rollouts`, `dd_weight`, ..., `LBD`, `SAMPLING_TYPE`), read
again at every step: weight changes are re-folded into the kernel parameters, a horizon change rebuilds the handle and
carries `u` over like update_control_vector (:543-553).

What runs where: the perturbations are drawn on the HOST with the reference's own generator (numpy SFC64) and sampler
code paths, so a seeded run sees bit-identical delta_u; trajectory_rollouts + update_inputs + the shift
(:164-215, :301-335, :531-536) are ONE kernel launch (cps_legacy_step_host -> legacy_mppi_kernel).  No CPU fallback.

Not mirrored: the matplotlib report (controller_report, :555-861) and the auxiliary-controller warm-up branch, which
the reference itself disables (:381-391 sets auxiliary_controller_available = False).
"""
from __future__ import annotations

from datetime import datetime

import numpy as np
from numpy.random import SFC64, Generator

from . import _lib as L
from .core import Engine

# Control_Toolkit_ASF/config_controllers.yml:9-30 ("mppi-cartpole")
DEFAULT_CONFIG = dict(seed=None, mpc_horizon=35, num_rollouts=3500, update_every=1, predictor_specification="ODE",
                      dd_weight=120.0, ep_weight=50000.0, ekp_weight=0.01, ekc_weight=5.0, cc_weight=1.0,
                      ccrc_weight=1.0, cost_noise=0.0, R=1.0, LBD=100.0, NU=1000.0, SQRTRHOINV=0.02,
                      SAMPLING_TYPE="interpolated", controller_logging=False, WASH_OUT_LEN=100)

_WEIGHTS = ("dd_weight", "ep_weight", "ekp_weight", "ekc_weight", "cc_weight", "ccrc_weight", "R", "LBD", "NU")


class controller_mppi_cartpole_b200:
    controller_name = "mppi-cartpole-b200"

    def __init__(self, config: dict | None = None, dt: float = 0.02, intermediate_steps: int = 10,
                 actuator_noise: float = 0.0, target_position: float = 0.0, device: int | None = None, **kwargs):
        """config: the "mppi-cartpole" block of config_controllers.yml (missing keys take the shipped values);
        dt: config_data_gen.yml dt.control (:48); actuator_noise: cartpole_physical_parameters.yml actuator_noise (p_Q, :81);
        predictor_specification: "ODE" (Euler-Cromer) or "ODE_v0" (explicit Euler + bounce), optionally "ODE:n"."""
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(config or {})
        cfg.update(kwargs)
        self.config = cfg
        for k in ("mpc_horizon", "num_rollouts", "update_every", "SAMPLING_TYPE", "cost_noise", "SQRTRHOINV") + _WEIGHTS:
            setattr(self, k, cfg[k])
        self.dt = float(dt)
        self.p_Q = float(actuator_noise)
        self.LOGGING = bool(cfg["controller_logging"])
        spec = str(cfg["predictor_specification"]).split(":")
        if spec[0] not in ("ODE", "ODE_v0"):
            raise NotImplementedError(f"controller_mppi_cartpole_b200: predictor {spec[0]!r} (ODE and ODE_v0 are built)")
        self.predictor_type = spec[0]
        self.intermediate_steps = int(spec[1]) if len(spec) > 1 else int(intermediate_steps)
        self.device = device
        self.target_position = np.float32(target_position)
        self.SQRTRHODTINV = self.SQRTRHOINV * (1 / np.sqrt(self.dt))  # numpy float64 scalar, as in the reference (:91)
        self.engine = None
        self._folded = None
        self.logs = {"cost_to_go": [], "inputs": [], "trajectory": [], "target_trajectory": []}
        self.configure()

    # -- reference interface ----------------------------------------------------------------------
    def configure(self):
        seed = self.config["seed"]
        if seed is None:
            seed = int((datetime.now() - datetime(1970, 1, 1)).total_seconds() * 1000.0)
        self.rng_mppi = Generator(SFC64(seed))
        self.rng_mppi_rnn = Generator(SFC64(seed * 2))
        # cost-weight jitter (:353-358): five draws even when cost_noise = 0, which positions the stream
        n = self.cost_noise
        self.dd_weight = self.dd_weight * (1 + n * self.rng_mppi.uniform(-1.0, 1.0))
        self.ep_weight = self.ep_weight * (1 + n * self.rng_mppi.uniform(-1.0, 1.0))
        self.ekp_weight = self.ekp_weight * (1 + n * self.rng_mppi.uniform(-1.0, 1.0))
        self.ekc_weight = self.ekc_weight * (1 + n * self.rng_mppi.uniform(-1.0, 1.0))
        self.cc_weight = self.cc_weight * (1 + n * self.rng_mppi.uniform(-1.0, 1.0))
        self.iteration = -1
        self.wash_out_len = self.config["WASH_OUT_LEN"]
        self.warm_up_countdown = self.wash_out_len
        self.u = np.zeros(self.mpc_horizon, dtype=np.float32)
        self.u_prev = np.zeros_like(self.u)
        self.delta_u = np.zeros((self.num_rollouts, self.mpc_horizon), dtype=np.float32)
        self.S_tilde_k = np.zeros(self.num_rollouts, dtype=np.float32)
        self._build_engine()

    def _build_engine(self):
        if self.engine is not None:
            self.engine.close()
        self.engine = Engine(self.num_rollouts, self.mpc_horizon, dt=self.dt, substeps=self.intermediate_steps,
                             integrator=self.predictor_type, cost="legacy_mppi", noise_mode="direct", device=self.device)
        self._engine_shape = (self.num_rollouts, self.mpc_horizon)
        self._folded = None
        self._du32 = np.zeros((self.num_rollouts, self.mpc_horizon), dtype=np.float32)

    def _sync_parameters(self):
        key = tuple(float(getattr(self, k)) for k in _WEIGHTS) + (float(self.target_position),)
        if key == self._folded:
            return
        e = self.engine
        e.set_cost_params([self.dd_weight, self.ep_weight, self.ekp_weight, self.ekc_weight, self.ccrc_weight])
        e.set_mppi_params(cc_weight=self.cc_weight, R=self.R, LBD=self.LBD, NU=self.NU, SQRTRHOINV=self.SQRTRHOINV)
        e.set_variable_parameters(target_position=float(self.target_position))
        self._folded = key

    INTERPOLATION_STEP = 10   # `step` of the interpolated sampler (:431)

    @property
    def delta_u(self) -> np.ndarray:
        """The perturbations of the last solve [K, T]; after a device-interpolated solve they are read back on first use."""
        if self._delta_u is None:
            self._delta_u = self.engine.legacy_get_perturbations()
        return self._delta_u

    @delta_u.setter
    def delta_u(self, value):
        self._delta_u = value

    def _interpolated_knots(self, stdev: float = 1.0) -> np.ndarray:
        """The draws of the "interpolated" sampler (:430-441) at the knots 0, 10, 20, ...: the reference assigns
        `stdev * rng.standard_normal(size=(K, n_knots), dtype=float32)` into its float32 array, i.e. rounds to float32."""
        K, T, step = self.num_rollouts, self.mpc_horizon, self.INTERPOLATION_STEP
        n_knots = int(np.ceil(T / step)) + 1
        return np.ascontiguousarray((stdev * self.rng_mppi.standard_normal(size=(K, n_knots), dtype=np.float32)).astype(np.float32))

    def initialize_perturbations(self, stdev: float = 1.0, sampling_type: str = None) -> np.ndarray:
        """The reference's five samplers (:392-457), same generator calls in the same order.  `stdev` is a numpy float64
        scalar in the reference; under NEP 50 `stdev * float32 array` is then float64, which the iid / repeated types
        return as is -- kept, and rounded to float32 only at the upload."""
        K, T, rng = self.num_rollouts, self.mpc_horizon, self.rng_mppi
        if sampling_type == "random_walk":
            delta_u = np.empty((K, T), dtype=np.float32)
            delta_u[:, 0] = stdev * rng.standard_normal(size=(K,), dtype=np.float32)
            for i in range(1, T):
                delta_u[:, i] = delta_u[:, i - 1] + stdev * rng.standard_normal(size=(K,), dtype=np.float32)
        elif sampling_type == "uniform":
            delta_u = np.empty((K, T), dtype=np.float32)
            for i in range(T):
                delta_u[:, i] = rng.uniform(low=-1.0, high=1.0, size=(K,)).astype(np.float32)
        elif sampling_type == "repeated":
            delta_u = np.tile(stdev * rng.standard_normal(size=(K, 1), dtype=np.float32), (1, T))
        elif sampling_type == "interpolated":
            from scipy.interpolate import interp1d
            step = 10
            range_stop = int(np.ceil(T / step) * step) + 1
            knots = np.arange(start=0, stop=range_stop, step=step)
            between = np.delete(np.arange(start=0, stop=range_stop, step=1), knots)
            delta_u = np.zeros(shape=(K, range_stop), dtype=np.float32)
            delta_u[:, knots] = stdev * rng.standard_normal(size=(K, knots.size), dtype=np.float32)
            delta_u[:, between] = interp1d(knots, delta_u[:, knots])(between)
            delta_u = delta_u[:, :T]
        else:
            delta_u = stdev * rng.standard_normal(size=(K, T), dtype=np.float32)
        return delta_u

    def update_attributes(self, updated_attributes: dict):
        for k, v in updated_attributes.items():
            if k == "target_position":
                self.target_position = np.float32(v)
            else:
                setattr(self, k, v)

    def step(self, s: np.ndarray, time=None, updated_attributes: dict = {}):
        self.update_attributes(updated_attributes)
        self.s = s
        self.iteration += 1
        if (self.num_rollouts, self.mpc_horizon) != self._engine_shape:  # changed in the GUI while running (:474-479)
            self.update_control_vector()
        if self.iteration % self.update_every == 0:
            if self.SAMPLING_TYPE == "interpolated":
                # the shipped sampler: only the knot draws are made here (the same generator call as the reference's); the
                # linear interpolation between them -- scipy's interp1d in the reference, three quarters of this controller's
                # host time -- runs on the device, bit-identical (cps_legacy_step_host_knots)
                knots = self._interpolated_knots(stdev=self.SQRTRHODTINV)
                self._sync_parameters()
                Q = self.engine.legacy_step_host_knots(np.asarray(s, dtype=np.float32), knots, self.INTERPOLATION_STEP)
                self._delta_u = None   # on the device; `delta_u` fetches it on demand
            else:
                self.delta_u = self.initialize_perturbations(stdev=self.SQRTRHODTINV, sampling_type=self.SAMPLING_TYPE)
                self._sync_parameters()
                np.copyto(self._du32, self.delta_u, casting="same_kind")
                Q = self.engine.legacy_step_host(np.asarray(s, dtype=np.float32), self._du32, L.ROLLOUT_MAJOR)
        else:
            Q = self.engine.legacy_advance()
        if self.LOGGING:
            u_next, u_upd = self.engine.legacy_get_inputs()
            self.logs["inputs"].append(u_upd)
            self.logs["trajectory"].append(np.copy(s))
            self.logs["target_trajectory"].append(np.copy(self.target_position))
        # actuator noise and clipping (:526-528); the uniform is drawn even when p_Q = 0
        Q = np.float32(Q * (1 + self.p_Q * self.rng_mppi.uniform(-1.0, 1.0)))
        Q = np.clip(Q, -1.0, 1.0, dtype=np.float32)
        return Q

    def _pull_inputs(self):
        """`u` (already shifted for the next iteration) and `u_prev`, as the reference holds them after step()."""
        self.u, self.u_prev = self.engine.legacy_get_inputs()
        return self.u, self.u_prev

    def update_control_vector(self):
        """Horizon / rollout-count change: new handle, `u` zero-padded or sliced, u_prev = u (:543-553)."""
        u_old, _ = self.engine.legacy_get_inputs()
        n = min(self.mpc_horizon, u_old.size)
        u_new = np.zeros(self.mpc_horizon, dtype=np.float32)
        u_new[:n] = u_old[:n]
        self._build_engine()
        self.engine.legacy_set_inputs(u_new, u_new)
        self.u, self.u_prev = u_new, u_new.copy()

    def controller_reset(self):
        self.logs = {k: [] for k in self.logs}
        self.warm_up_countdown = self.wash_out_len

    def controller_report(self):
        raise NotImplementedError("the matplotlib report of controller_mppi_cartpole (:555-861) is out of scope; "
                                  "read .logs / engine outputs instead")
