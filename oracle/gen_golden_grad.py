"""Golden vectors for the gradient of the trajectory cost with respect to the input plan (the quantity the reference's
gradient-based optimizers -- RPGD: Control_Toolkit/Optimizers/optimizer_rpgd_tf.py:167-180 -- obtain from a GradientTape
around predict_and_cost): tests/golden/grad_*.npz.

TEST INFRASTRUCTURE ONLY.  TensorFlow is absent here; the same composition -- the UNMODIFIED predictor_ODE
(SI_Toolkit/Predictors/predictor_ODE.py, torch library) followed by the UNMODIFIED cost plugin's get_trajectory_cost --
is differentiated with torch autograd instead, in float32 as the reference runs it and once more in float64 (same
modules, inputs cast up) to tell truncation from rounding when the CUDA adjoint is compared.

    python oracle/gen_golden_grad.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import zlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import ref_loader as R  # noqa: E402
from oracle.gen_golden import hanging_state, make_states, save  # noqa: E402


def grad_of_cost(pred, cost, s, Q, u_prev, tp, te, overrides=None):
    import torch
    K, T = Q.shape
    lib = R.torch_lib()
    if cost == "quadratic_boundary_grad":
        # the plugin is TF-only in the reference: PyTorchLibrary has no stop_gradient / cond (computation_library.py:501).
        # Their torch counterparts: detach() is what tf.stop_gradient does to the tape; cond on a Python bool.
        lib.stop_gradient = lambda x: x.detach() if hasattr(x, "detach") else x
        lib.cond = lambda c, true_fn, false_fn: true_fn() if bool(c) else false_fn()
    vp = R.variable_parameters(lib, tp, te)
    cw = R.cost_function(cost, lib, vp, K, T)
    for key, value in (overrides or {}).items():   # config values of the plugin (config_cost_function.yml), not its code
        assert hasattr(cw.cost_function, key), key
        setattr(cw.cost_function, key, lib.to_variable(value, lib.float32))
    predictor = R.ODECoreAdapter(T, K, 0.02, 10, vp)
    s_t = torch.from_numpy(np.tile(s, (K, 1)).astype(np.float32))
    Qv = torch.from_numpy(Q[:, :, None].copy()).requires_grad_(True)
    traj = predictor.predict_core(s_t, Qv)
    J = cw.get_trajectory_cost(traj, Qv, np.float32(u_prev))
    J.sum().backward()
    return J.detach().numpy().reshape(-1).astype(np.float32), Qv.grad.numpy()[:, :, 0].astype(np.float32), traj.detach().numpy()


def main():
    if not R.available():
        raise SystemExit("reference tree not available; fixtures can only be regenerated in the build container")
    R.load()
    runs = [  # name, cost, K, T, state kind, plan scale, target position, target equilibrium, config overrides
        ("gradmin_K16_T35", "quadratic_boundary_grad_minimal", 16, 35, "hanging", 0.5, 0.0, 1.0, None),
        ("gradmin_K64_T20", "quadratic_boundary_grad_minimal", 64, 20, "random", 0.8, 0.05, 1.0, None),
        ("gradmin_down_K32_T50", "quadratic_boundary_grad_minimal", 32, 50, "hanging", 0.3, -0.03, -1.0, None),
        # quadratic_boundary_grad: the shipped weights; then the linear-distance and control-change terms switched on
        # (zero as shipped), equilibrium down with its angular-speed correction
        ("qbgrad_K16_T35", "quadratic_boundary_grad", 16, 35, "hanging", 0.5, 0.0, 1.0, None),
        ("qbgrad_terms_K64_T20", "quadratic_boundary_grad", 64, 20, "random", 0.8, 0.05, 1.0,
         dict(dd_linear_weight_up=40.0, ccrc_weight_up=2.0)),
        ("qbgrad_down_K32_T50", "quadratic_boundary_grad", 32, 50, "hanging", 0.3, -0.03, -1.0,
         dict(dd_linear_weight_down=40.0, ccrc_weight_down=2.0)),
    ]
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        try:
            for name, cost, K, T, kind, scale, tp, te, over in runs:
                rng = np.random.default_rng(zlib.crc32(name.encode()))
                s = hanging_state() if kind == "hanging" else make_states(rng, 1, "random")[0]
                if kind == "random":
                    s[4] = 0.17   # near the track boundary: the boundary-approach term and its derivative are active
                Q = np.clip(rng.normal(0.0, scale, (K, T)), -1, 1).astype(np.float32)
                J, G, traj = grad_of_cost("ODE", cost, s, Q, 0.1, tp, te, over)
                save("grad_" + name, dict(ref="torch autograd through SI_Toolkit/Predictors/predictor_ODE.py (torch library) + "
                                              "Control_Toolkit_ASF/Cost_Functions/CartPole/%s.py get_trajectory_cost; the "
                                              "reference takes the same derivative with tf.GradientTape "
                                              "(optimizer_rpgd_tf.py:167-180)" % cost,
                                          predictor="ODE", cost=cost, K=K, T=T, u_prev=0.1, target_position=tp, target_equilibrium=te,
                                          cost_config_overrides=over or {}),
                     s=s.astype(np.float32), Q=Q, J=J, G=G, traj_last=traj[:, -1].astype(np.float32))
        finally:
            txt = buf.getvalue()
    print("\n".join(l for l in txt.splitlines() if l.startswith("wrote")))


if __name__ == "__main__":
    main()
