"""Generate tests/golden/legacy_*.npz by running the UNMODIFIED legacy controller of the reference
(Control_Toolkit_ASF/Controllers/controller_mppi_cartpole.py) in this container.

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/gen_golden_legacy.py
The controller's parameters are module globals (the GUI mutates them at run time,
GUI/_ControllerGUI_MPPIOptionsWindow.py:29-46); the harness sets them the same way, swaps the module-level `predictor`
for the reference's own predictor_ODE_v0 / torch predictor_ODE, and builds the controller object without
template_controller.__init__ (which only reads config files): configure() and step() run unmodified.
The plant of the closed loop is the model itself (next state = predictor's one-step prediction with the returned Q).
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")

from oracle import ref_loader as R  # noqa: E402


class _PredShim:
    """What the module-level PredictorWrapper offers to this controller: predict / update / horizon / predictor_type."""

    def __init__(self, kind, T, K):
        self.kind, self.horizon, self.predictor_type = kind, T, kind
        if kind == "ODE_v0":
            self.p = R.predictor_ODE_v0(T, 0.02, 10, K)
        else:
            self.p = R.predictor_ODE(T, 0.02, 10, K)

    def predict(self, s, Q):
        if self.kind == "ODE_v0":
            return self.p.predict(np.asarray(s, np.float32), np.asarray(Q, np.float32))
        import torch
        out = self.p.predict_core(torch.from_numpy(np.asarray(s, np.float32)), torch.from_numpy(np.asarray(Q, np.float32)))
        return out.numpy()

    def update(self, Q0, s):
        pass


def run(name, kind, K, T, sampling, seed, n_steps, s0, target_position=0.0, update_every=1, weights=None):
    R.load()
    m = importlib.import_module("Control_Toolkit_ASF.Controllers.controller_mppi_cartpole")
    m.num_rollouts, m.mpc_horizon, m.update_every, m.SAMPLING_TYPE = K, T, update_every, sampling
    m.LOGGING = False
    m.config_mppi_cartpole["seed"] = seed
    base = dict(dd_weight=120.0, ep_weight=50000.0, ekp_weight=0.01, ekc_weight=5.0, cc_weight=1.0, ccrc_weight=1.0)
    base.update(weights or {})
    for k, v in base.items():
        setattr(m, k, v)
    m.predictor = _PredShim(kind, T, K)
    one_step = _PredShim(kind, 1, 1)
    ctrl = object.__new__(m.controller_mppi_cartpole)
    ctrl.variable_parameters = types.SimpleNamespace(target_position=np.float32(target_position))
    ctrl.update_attributes = lambda d: None
    ctrl.configure()
    rec = dict(s=[], delta_u=[], S=[], u_updated=[], Q=[], u_in=[], u_prev_in=[])
    s = np.asarray(s0, np.float32)
    for it in range(n_steps):
        u_in, u_prev_in = ctrl.u.copy(), ctrl.u_prev.copy()
        Q = ctrl.step(s.copy())
        rec["s"].append(s.copy())
        rec["u_in"].append(u_in)
        rec["u_prev_in"].append(u_prev_in)
        rec["delta_u"].append(np.array(ctrl.delta_u, np.float32))
        rec["S"].append(np.array(ctrl.S_tilde_k, np.float32))
        rec["u_updated"].append(ctrl.u_prev.copy())  # u_prev = copy of u after the update, before the shift (:531)
        rec["Q"].append(np.float32(Q))
        nxt = one_step.predict(s[None, :], np.full((1, 1, 1), Q, np.float32))
        s = np.asarray(nxt, np.float32).reshape(2, 6)[1]
    out = {k: np.stack(v) for k, v in rec.items()}
    meta = dict(reference="Control_Toolkit_ASF/Controllers/controller_mppi_cartpole.py: configure() + step(), unmodified",
                predictor=kind, K=K, T=T, sampling=sampling, seed=seed, update_every=update_every,
                target_position=target_position, weights=base, p_Q=float(m.p_Q), R=m.R, LBD=m.LBD, NU=m.NU,
                SQRTRHODTINV=float(m.SQRTRHODTINV), dt=0.02, n=10)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), **out)
    print(name, "Q:", out["Q"][:5], "S range", out["S"].min(), out["S"].max())


def hanging(eps=1e-3):
    a = np.pi - eps
    return np.array([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], dtype=np.float32)


if __name__ == "__main__":
    up = np.array([0.15, -0.4, np.cos(0.15), np.sin(0.15), 0.05, 0.1], dtype=np.float32)
    run("legacy_v0_interpolated", "ODE_v0", 512, 35, "interpolated", 7, 12, hanging())
    run("legacy_v0_iid_upright", "ODE_v0", 256, 20, "iid", 11, 8, up, target_position=0.05)
    run("legacy_ode_random_walk", "ODE", 256, 35, "random_walk", 3, 6, hanging(), update_every=2)
    run("legacy_v0_uniform", "ODE_v0", 128, 15, "uniform", 5, 4, up)
    run("legacy_v0_repeated", "ODE_v0", 128, 15, "repeated", 9, 4, hanging(),
        weights=dict(dd_weight=60.0, ekc_weight=1.0, ccrc_weight=3.0))
