"""Fleets: many independent closed-loop experiments advanced together on the device (include/cps.h "fleet").

Host-side mirror of what the reference does for ONE cartpole per process when it generates data
(CartPole/data_generator.py:94-236 random_experiment_setter + generate_random_initial_state,
CartPole/random_target_generator.py:9-87, CartPole/__init__.py:578-739 setup/run_cartpole_random_experiment):
random initial states, a random target-position trace per experiment, the up/down target-equilibrium schedule, then
controller period after controller period of [MPPI solve -> plant].  Here the solve and the plant of ALL experiments
run in one kernel launch per controller period (cps_fleet_step); the host only prepares the target tables and, at the
end, reads the record rows back.  No CPU fallback: everything below the table preparation happens in libcps_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib as L
from .core import Engine, _check_dev, _ptr

TRACK_HALF_LENGTH = float(np.float32((44.0e-2 - 4.4e-2) / 2.0))


@dataclass
class DataGenConfig:
    """config_data_gen.yml:6-31 (defaults as shipped)."""
    length_of_experiment: float = 360.0
    angle_init_limits: tuple = (0.0, 180.0)          # degrees; mirrored to the left half plane with p = 0.5
    angleD_init_limit: float = 1200.0                 # degrees / s
    position_init_limit: float = 0.8                  # fraction of TrackHalfLength
    positionD_init_limit: float = 0.5
    start_at_target: bool = True
    track_fraction_usable_for_target_position: float = 1.0
    target_position_end: float | None = None
    initial_target_equilibrium: int = 1
    keep_target_equilibrium_x_seconds_up: float = 10.0
    keep_target_equilibrium_x_seconds_down: float = 2.5
    dt_simulation: float = 0.002
    dt_control: float = 0.02
    track_relative_complexity: float = 1.0
    interpolation_type: tuple = ("previous", "0-derivative-smooth")   # cycled over the experiments (:205-209)
    turning_points_period: str = "regular"
    position: float | None = None
    positionD: float | None = None
    angle: float | None = None
    angleD: float | None = None
    extra: dict = field(default_factory=dict)


def random_initial_state(rng, cfg: DataGenConfig) -> np.ndarray:
    """generate_random_initial_state (CartPole/data_generator.py:238-275); draws in the same order."""
    s = np.zeros(6, dtype=np.float32)
    thl, f32 = np.float32(TRACK_HALF_LENGTH), np.float32
    # python float * float32 0-d array * python float: float32 arithmetic under NEP 50, as in the reference
    s[4] = f32(rng.uniform(-1.0, 1.0)) * thl * f32(cfg.position_init_limit) if cfg.position is None else cfg.position
    s[5] = f32(rng.uniform(-1.0, 1.0)) * thl * f32(cfg.positionD_init_limit) if cfg.positionD is None else cfg.positionD
    if cfg.angle is None:
        lo, hi = cfg.angle_init_limits
        if rng.uniform() > 0.5:
            s[0] = rng.uniform(lo, hi) * (np.pi / 180.0)
        else:
            s[0] = rng.uniform(-hi, -lo) * (np.pi / 180.0)
    else:
        s[0] = cfg.angle
    s[1] = rng.uniform(-1.0, 1.0) * cfg.angleD_init_limit * (np.pi / 180.0) if cfg.angleD is None else cfg.angleD
    s[2], s[3] = np.cos(s[0]), np.sin(s[0])
    return s


def random_trace_function(rng, cfg: DataGenConfig, interpolation_type: str, start_at, end_at):
    """Generate_Random_Trace_Function (CartPole/random_target_generator.py:9-87)."""
    from scipy.interpolate import BPoly, interp1d
    frac = cfg.track_fraction_usable_for_target_position
    n_tp = int(np.floor(cfg.length_of_experiment * cfg.track_relative_complexity))
    y = rng.uniform(-1.0, 1.0, n_tp) * frac * TRACK_HALF_LENGTH
    if n_tp == 0:
        y = np.array([0.0, 0.0])
    elif n_tp == 1:
        if start_at is not None:
            y[0] = start_at
        elif end_at is not None:
            y[0] = end_at
        y = np.append(y, y[0])
    else:
        if start_at is not None:
            y[0] = start_at
        if end_at is not None:
            y[-1] = end_at
    random_samples = max(n_tp - 2, 0)
    if cfg.turning_points_period == "random":
        t_init = np.concatenate([[0.0], np.sort(rng.uniform(0.0, 1.0, random_samples)), [1.0]])
    elif cfg.turning_points_period == "regular":
        t_init = np.linspace(0, 1.0, num=random_samples + 2, endpoint=True)
    else:
        raise NotImplementedError("There is no mode corresponding to this value of turning_points_period variable")
    t_init = t_init * cfg.length_of_experiment
    if interpolation_type == "0-derivative-smooth":
        f = BPoly.from_derivatives(t_init, [[v, 0] for v in y], extrapolate="periodic")
    elif interpolation_type == "linear":
        f = interp1d(t_init, y, kind="linear", fill_value="extrapolate")
    elif interpolation_type == "previous":
        f = interp1d(t_init, y, kind="previous", fill_value="extrapolate")
    else:
        raise ValueError("Unknown interpolation type.")
    lim = frac * TRACK_HALF_LENGTH
    return lambda t: np.clip(f(t), -lim, lim)


def control_times(n_periods: int, cfg: DataGenConfig) -> np.ndarray:
    """Plant time at which the controller is called in period j: the reference accumulates self.time + dt_simulation
    tick by tick (CartPole/__init__.py:326-327); period 0 is solved at time 0 (set_cartpole_state_at_t0, :866-881)."""
    n_sim = int(round(cfg.dt_control / cfg.dt_simulation))
    t = np.concatenate([[0.0], np.cumsum(np.full(n_periods * n_sim, cfg.dt_simulation))])
    return t[::n_sim][:n_periods]


def target_equilibrium_schedule(n_periods: int, cfg: DataGenConfig, te0: int | None = None) -> np.ndarray:
    """update_target_equilibrium (CartPole/__init__.py:380-388) evaluated on every plant tick; returns the value the
    controller sees in each period, float32 [n_periods]."""
    n_sim = int(round(cfg.dt_control / cfg.dt_simulation))
    cur = float(cfg.initial_target_equilibrium if te0 is None else te0)
    out = np.empty(n_periods, dtype=np.float32)
    out[0] = cur
    t, t_last = 0.0, None
    for i in range(1, (n_periods - 1) * n_sim + 1):
        t = t + cfg.dt_simulation
        if t_last is None:
            t_last = t
        elif cur == -1 and (t - t_last) > cfg.keep_target_equilibrium_x_seconds_down:
            t_last, cur = t, -cur
        elif cur == 1 and (t - t_last) > cfg.keep_target_equilibrium_x_seconds_up:
            t_last, cur = t, -cur
        if i % n_sim == 0:
            out[i // n_sim] = cur
    return out


def make_experiments(n_experiments: int, n_periods: int, cfg: DataGenConfig | None = None, seed: int = 0,
                     experiment_offset: int = 0):
    """Initial states [E,6], target-position table [n_periods,E] and target-equilibrium table [n_periods,E] for
    experiments experiment_offset .. experiment_offset+E-1.  Each experiment has its own numpy Generator seeded with
    (seed, global index), so the tables do not depend on how the experiments are sharded over GPUs."""
    cfg = cfg or DataGenConfig()
    times = control_times(n_periods, cfg)
    s0 = np.zeros((n_experiments, 6), dtype=np.float32)
    tp = np.zeros((n_periods, n_experiments), dtype=np.float32)
    te_one = target_equilibrium_schedule(n_periods, cfg)
    te = np.repeat(te_one[:, None], n_experiments, axis=1)
    itypes = cfg.interpolation_type if isinstance(cfg.interpolation_type, (list, tuple)) else (cfg.interpolation_type,)
    for e in range(n_experiments):
        g = experiment_offset + e
        rng = np.random.default_rng([seed, g])
        s0[e] = random_initial_state(rng, cfg)
        start_at = float(s0[e, 4]) if cfg.start_at_target else \
            cfg.track_fraction_usable_for_target_position * TRACK_HALF_LENGTH * rng.uniform(-1.0, 1.0)
        end_at = cfg.track_fraction_usable_for_target_position * TRACK_HALF_LENGTH * rng.uniform(-1.0, 1.0) \
            if cfg.target_position_end is None else cfg.target_position_end
        f = random_trace_function(rng, cfg, itypes[g % len(itypes)], start_at, end_at)
        tp[:, e] = np.asarray(f(np.minimum(times, cfg.length_of_experiment)), dtype=np.float64).astype(np.float32)
    return s0, tp, te


class Fleet:
    """E closed-loop experiments on one device = one cps_handle with a fleet attached."""

    def __init__(self, n_experiments: int, num_rollouts: int = 2000, horizon: int = 50, dt: float = 0.02,
                 substeps: int = 10, integrator: str = "ODE", cost: str = "quadratic_boundary_grad_minimal",
                 interp_period: int = 10, device: int | None = None, noise: str = "philox", seed: int = 0,
                 experiment_offset: int = 0, dt_simulation: float = 0.002, no_pairs: bool = False):
        if noise not in ("philox", "supplied"):
            raise ValueError("noise must be 'philox' or 'supplied'")
        self.engine = Engine(num_rollouts, horizon, dt=dt, substeps=substeps, integrator=integrator, cost=cost,
                             interp_period=interp_period, device=device, no_pairs=no_pairs)
        self.E, self.K, self.T = int(n_experiments), int(num_rollouts), int(horizon)
        self.n_ind, self.device = self.engine.n_ind, self.engine.device
        n_sim = int(round(dt / dt_simulation))
        if abs(n_sim * dt_simulation - dt) > 1e-9:
            raise ValueError("dt (control) must be a multiple of dt_simulation")
        c = L.cps_fleet_config(C.sizeof(L.cps_fleet_config), self.E, n_sim,
                               L.FLEET_NOISE_PHILOX if noise == "philox" else L.FLEET_NOISE_SUPPLIED,
                               float(dt_simulation), int(seed), int(experiment_offset))
        self.noise_source = noise
        self.engine._chk(self.engine.lib.cps_fleet_create(self.engine._h, C.byref(c)))

    def close(self):
        self.engine.close()

    def reset(self, states, period: int = 0):
        s = np.ascontiguousarray(states, dtype=np.float32)
        if s.shape != (self.E, 6):
            raise ValueError(f"states must have shape ({self.E}, 6)")
        self.engine.use_current_stream()
        self.engine._chk(self.engine.lib.cps_fleet_set_states(self.engine._h, s.ctypes.data_as(L._FP), int(period)))

    def states(self, with_u_nom: bool = False):
        s = np.zeros((self.E, 6), dtype=np.float32)
        u_nom = np.zeros((self.E, self.T), dtype=np.float32) if with_u_nom else None
        u_prev = np.zeros(self.E, dtype=np.float32) if with_u_nom else None
        self.engine.use_current_stream()
        self.engine._chk(self.engine.lib.cps_fleet_get_states(
            self.engine._h, s.ctypes.data_as(L._FP), None if u_nom is None else u_nom.ctypes.data_as(L._FP),
            None if u_prev is None else u_prev.ctypes.data_as(L._FP)))
        return (s, u_nom, u_prev) if with_u_nom else s

    @property
    def period(self) -> int:
        return int(self.engine.lib.cps_fleet_period(self.engine._h))

    def noise(self, period: int) -> torch.Tensor:
        """The draws the Philox source uses in `period`: cuda tensor [E, n_ind, K]."""
        out = torch.empty((self.E, self.n_ind, self.K), device=self.device, dtype=torch.float32)
        self.engine.use_current_stream()
        self.engine._chk(self.engine.lib.cps_fleet_noise(self.engine._h, int(period), _ptr(out)))
        return out

    _CONTROL_NOISE = {"OFF": 0, "additive": 1, "truncnorm": 2}

    def set_plant_models(self, control_noise_mode: str = "OFF", control_noise: float = 0.0, control_bias: float = 0.0,
                         noise_mode: str = "OFF", sigma_angle: float = 0.0, sigma_position: float = 0.0,
                         sigma_angleD: float = 0.0, sigma_positionD: float = 0.0, latency: float = 0.0):
        """cps_fleet_set_plant_models with the reference's configuration keys (cartpole_physical_parameters.yml:13-24:
        controlDisturbance_mode / controlDisturbance / controlBias, latency, noise.noise_mode / sigma_*).  Call before
        reset(); the first solve after reset() sees the true state (the reference's controller call at t = 0)."""
        if control_noise_mode not in self._CONTROL_NOISE:
            raise ValueError(f"controlDisturbance_mode with value {control_noise_mode} not valid")
        m = L.cps_fleet_plant_models(C.sizeof(L.cps_fleet_plant_models), self._CONTROL_NOISE[control_noise_mode],
                                     float(control_noise), float(control_bias), 0 if noise_mode == "OFF" else 1,
                                     float(sigma_angle), float(sigma_position), float(sigma_angleD), float(sigma_positionD),
                                     float(latency))
        self.engine.use_current_stream()
        self.engine._chk(self.engine.lib.cps_fleet_set_plant_models(self.engine._h, C.byref(m)))
        self.plant_models = bool(m.control_noise_mode or m.measurement_noise or m.latency > 0.0)

    def observed(self) -> np.ndarray:
        """[E, 6]: the states as the controller will see them at the next solve (delayed, noisy)."""
        o = np.zeros((self.E, 6), dtype=np.float32)
        self.engine.use_current_stream()
        self.engine._chk(self.engine.lib.cps_fleet_get_observed(self.engine._h, o.ctypes.data_as(L._FP)))
        return o

    def run(self, n_periods: int, target_position=None, target_equilibrium=None, noise=None, record=None, J_out=None,
            ctrl_draws=None, meas_draws=None):
        """cps_fleet_step: n_periods launches, no synchronisation.  target_position / target_equilibrium: cuda tensors
        [n_periods, E] or None; noise: cuda tensor [n_periods, E, n_ind, K] for a 'supplied' fleet;
        record: cuda tensor [n_periods, E, 16] or None; J_out: [n_periods, E, K] or None.  With plant-side models
        (set_plant_models): ctrl_draws [n_periods, E] and meas_draws [n_periods, sim_substeps, E, 4] supply their draws
        (cps_fleet_step_noisy); a 'philox' fleet draws them itself when they are None."""
        eng = self.engine
        eng.use_current_stream()
        for t, name, numel in ((target_position, "target_position", n_periods * self.E),
                               (target_equilibrium, "target_equilibrium", n_periods * self.E),
                               (noise, "noise", n_periods * self.E * self.n_ind * self.K),
                               (record, "record", n_periods * self.E * L.FLEET_RECORD),
                               (J_out, "J_out", n_periods * self.E * self.K)):
            if t is not None:
                _check_dev(t, name, self.device)
                if t.numel() != numel:
                    raise ValueError(f"{name} has {t.numel()} elements, expected {numel}")
        if getattr(self, "plant_models", False):
            for t, name in ((ctrl_draws, "ctrl_draws"), (meas_draws, "meas_draws")):
                if t is not None:
                    _check_dev(t, name, self.device)
            eng._chk(eng.lib.cps_fleet_step_noisy(eng._h, int(n_periods), _ptr(target_position), _ptr(target_equilibrium),
                                                  _ptr(noise), _ptr(record), _ptr(J_out), _ptr(ctrl_draws), _ptr(meas_draws)))
            return record
        eng._chk(eng.lib.cps_fleet_step(eng._h, int(n_periods), _ptr(target_position), _ptr(target_equilibrium),
                                        _ptr(noise), _ptr(record), _ptr(J_out)))
        return record
