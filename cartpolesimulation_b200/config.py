"""Default parameters of the path (values of the reference's YAML files) and the packers that turn them
into the flat vectors the C ABI takes.  If the reference's YAML files are present in the working directory
(drop-in use inside a CartPoleSimulation checkout) they are read instead, with the same keys.

Sources: cartpole_physical_parameters.yml:6-17,33-42; Control_Toolkit_ASF/config_cost_function.yml:5-58;
Control_Toolkit_ASF/config_optimizers.yml:87-97; SI_Toolkit_ASF/config_predictors.yml:18-25.
"""
from __future__ import annotations

import copy
import os

import numpy as np

PHYS_ORDER = ["k", "m_cart", "m_pole", "g", "J_fric", "M_fric", "L", "u_max", "TrackHalfLength"]

DEFAULT_PHYSICS = dict(k=1.0 / 3.0, m_cart=0.230, m_pole=0.087, g=9.81, J_fric=5.0e-5, M_fric=3.22, L=0.395,
                       u_max=1.77, TrackHalfLength=(44.0e-2 - 4.4e-2) / 2.0)

DEFAULT_COST_CONFIG = {
    "default": dict(dd_weight=600.0, ep_weight=20000.0, cc_weight=1.0, ccrc_weight=1.0, R=1.0),
    "quadratic_boundary": dict(dd_weight=600.0, ep_weight=20000.0, cc_weight=1.0, ccrc_weight=1.0, R=1.0),
    "quadratic_boundary_grad_minimal": dict(dd_quadratic_weight_up=10.0, ep_weight_up=40.0, ekp_weight_up=1.0,
                                            db_weight_up=10000.0, cc_weight_up=5.0, R=1.0,
                                            permissible_track_fraction=0.85),
    "quadratic_boundary_grad": dict(
        dd_quadratic_weight_up=500.0, dd_linear_weight_up=0.0, ep_weight_up=6000.0,
        target_angular_speed_sqr_max_correction_up=0.0, ekp_weight_up=30.0, db_weight_up=10000.0,
        cc_weight_up=5.0, ccrc_weight_up=0.0,
        dd_quadratic_weight_down=500.0, dd_linear_weight_down=0.0, ep_weight_down=6000.0,
        target_angular_speed_sqr_max_correction_down=100.0, ekp_weight_down=30.0, db_weight_down=10000.0,
        cc_weight_down=5.0, ccrc_weight_down=0.0,
        permissible_track_fraction=0.85, admissible_angle=0.0, R=1.0),
}

DEFAULT_MPPI_CONFIG = dict(seed=None, mpc_horizon=35, mpc_timestep=0.02, num_rollouts=3500, cc_weight=1.0, R=1.0,
                           LBD=100.0, NU=1000.0, SQRTRHOINV=0.03, period_interpolation_inducing_points=10)

DEFAULT_PREDICTOR_CONFIG = {
    "ODE_v0_default": dict(predictor_type="ODE_v0", model_name=None, intermediate_steps=10),
    "ODE_default": dict(predictor_type="ODE", model_name=None, intermediate_steps=10),
    "neural_default": dict(predictor_type="neural", model_name="GRU-6IN-32H1-32H2-5OUT-0",
                           path_to_model="./SI_Toolkit_ASF/Experiments/Experiment-2/Models/",
                           update_before_predicting=True),
}


def _load_yaml(path):
    try:
        import yaml
        with open(path, "r") as f:
            return yaml.safe_load(f)
    except Exception:
        return None


def cost_config(name: str) -> dict:
    """Weights of cost plugin `name`: the YAML in the cwd if there is one (the reference reads it cwd-relative at
    import time, e.g. quadratic_boundary.py:13-19), else the shipped defaults."""
    y = _load_yaml(os.path.join("Control_Toolkit_ASF", "config_cost_function.yml"))
    if y and "CartPole" in y and name in y["CartPole"]:
        return dict(y["CartPole"][name])
    if name not in DEFAULT_COST_CONFIG:
        raise ValueError(f"unknown cost function {name!r}")
    return copy.deepcopy(DEFAULT_COST_CONFIG[name])


def mppi_config() -> dict:
    y = _load_yaml(os.path.join("Control_Toolkit_ASF", "config_optimizers.yml"))
    if y and "mppi" in y:
        return dict(y["mppi"])
    return dict(DEFAULT_MPPI_CONFIG)


def physics_vector(**overrides) -> np.ndarray:
    """fp32-rounded physical constants in CPS_PH_* order (CartPole/cartpole_parameters.py:27-29 casts to fp32)."""
    d = dict(DEFAULT_PHYSICS)
    unknown = set(overrides) - set(d)
    if unknown:
        raise ValueError(f"unknown physical parameters {sorted(unknown)}")
    d.update(overrides)
    return np.array([np.float32(d[k]) for k in PHYS_ORDER], dtype=np.float32)


def max_cost(name: str, cfg: dict, u_max=1.77) -> np.float32:
    """MAX_COST of the shifted plugins (default.py:20, quadratic_boundary.py:23).  u_max is a 0-d float32 array
    there and python floats are weakly typed, so the expression evaluates in float32."""
    u2 = np.float32(u_max) ** 2
    base = np.float32(cfg["dd_weight"] * 1.0e7 + cfg["ep_weight"] + np.float32(cfg["cc_weight"] * cfg["R"]) * u2)
    if name == "quadratic_boundary":
        base = np.float32(base + np.float32(cfg["ccrc_weight"] * 4) * u2)
    return base


def cost_vector(name: str | None, cfg: dict | None = None, u_max=1.77) -> np.ndarray:
    """Flat weight vector in the order cps_set_cost_params documents (DESIGN.md "cost parameter vectors")."""
    if name in (None, "none"):
        return np.zeros(0, dtype=np.float32)
    cfg = cost_config(name) if cfg is None else cfg
    if name in ("default", "quadratic_boundary"):
        v = [cfg["dd_weight"], cfg["ep_weight"], cfg["cc_weight"], cfg["ccrc_weight"], cfg["R"],
             max_cost(name, cfg, u_max)]
    elif name == "quadratic_boundary_grad_minimal":
        v = [cfg["dd_quadratic_weight_up"], cfg["db_weight_up"], cfg["ep_weight_up"], cfg["ekp_weight_up"],
             cfg["cc_weight_up"], cfg["R"], cfg["permissible_track_fraction"]]
    elif name == "quadratic_boundary_grad":
        v = []
        for sfx in ("_up", "_down"):
            v += [cfg["dd_quadratic_weight" + sfx], cfg["dd_linear_weight" + sfx], cfg["db_weight" + sfx],
                  cfg["ep_weight" + sfx], cfg["ekp_weight" + sfx], cfg["cc_weight" + sfx], cfg["ccrc_weight" + sfx],
                  cfg["target_angular_speed_sqr_max_correction" + sfx]]
        v += [cfg["R"], cfg["permissible_track_fraction"],
              np.float32(np.pi) * np.float32(cfg["admissible_angle"]) / np.float32(180.0)]
    else:
        raise ValueError(f"unknown cost function {name!r}")
    return np.array([np.float32(x) for x in v], dtype=np.float32)
