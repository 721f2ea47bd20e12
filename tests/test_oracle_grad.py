"""Pins the hand-derived adjoint of predict_and_cost (oracle.plan_cost_grad: reverse sweep of the Euler-Cromer substep and the
quadratic_boundary_grad_minimal / quadratic_boundary_grad plugins) against torch autograd through the UNMODIFIED reference modules
(tests/golden/grad_*.npz, oracle/gen_golden_grad.py) -- the derivative the reference's RPGD takes with a GradientTape
(Control_Toolkit/Optimizers/optimizer_rpgd_tf.py:167-175)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.parity import load_golden

GRAD = ["grad_gradmin_K16_T35", "grad_gradmin_K64_T20", "grad_gradmin_down_K32_T50",
        "grad_qbgrad_K16_T35", "grad_qbgrad_terms_K64_T20", "grad_qbgrad_down_K32_T50"]


def cost_cfg(m):
    cfg = dict(O.DEFAULT_COST_CONFIG[m["cost"]])
    cfg.update(m.get("cost_config_overrides", {}))
    return cfg


@pytest.mark.parametrize("name", GRAD)
def test_adjoint_matches_autograd_through_the_reference(name):
    z, m = load_golden(name)
    J, G = O.plan_cost_grad(m["cost"], z["s"], z["Q"], m["u_prev"], m["target_position"], m["target_equilibrium"],
                            cost_cfg=cost_cfg(m))
    # float64 restatement against the reference's float32 autograd: its rounding noise is ~1e-5 of the largest entry; the
    # 500-substep quadratic_boundary_grad case measures 1.1e-5 (J) / 2.2e-5 (G)
    tol_J, tol_G = (2e-5, 4e-5) if name == "grad_qbgrad_down_K32_T50" else (1e-5, 2e-5)
    assert np.abs(J - z["J"]).max() <= tol_J * np.abs(z["J"]).max()
    assert np.abs(G - z["G"]).max() <= tol_G * np.abs(z["G"]).max()


def test_adjoint_matches_finite_differences():
    """Central differences of the same restatement's cost (float64): the derivation itself, independently of autograd."""
    z, m = load_golden("grad_gradmin_K16_T35")
    Q = z["Q"][:3].astype(np.float64)
    J, G = O.plan_cost_grad(m["cost"], z["s"], Q, 0.0, 0.02, 1.0)
    eps = 1e-6
    for t in (0, 7, 20, 34):
        Qp, Qm = Q.copy(), Q.copy()
        Qp[:, t] += eps
        Qm[:, t] -= eps
        fd = (O.plan_cost_grad(m["cost"], z["s"], Qp, 0.0, 0.02, 1.0)[0] - O.plan_cost_grad(m["cost"], z["s"], Qm, 0.0, 0.02, 1.0)[0]) / (2 * eps)
        np.testing.assert_allclose(G[:, t], fd, rtol=1e-5, atol=1e-7)


def test_adjoint_of_quadratic_boundary_grad_matches_finite_differences():
    """The same for quadratic_boundary_grad with its linear-distance and control-change terms on.  The plugin blocks the
    gradient through the kinetic term's angle-dependent target (stop_gradient, quadratic_boundary_grad.py:133-139); with the
    angular-speed correction chosen so that the target vanishes, the restatement must agree with plain differences."""
    z, m = load_golden("grad_qbgrad_terms_K64_T20")
    cfg = cost_cfg(m)
    cfg["target_angular_speed_sqr_max_correction_up"] = -120.0
    Q = z["Q"][:3].astype(np.float64)
    args = (m["cost"], z["s"])
    J, G = O.plan_cost_grad(*args, Q, 0.1, 0.05, 1.0, cost_cfg=cfg)
    eps = 1e-6
    for t in (0, 7, 19):
        Qp, Qm = Q.copy(), Q.copy()
        Qp[:, t] += eps
        Qm[:, t] -= eps
        fd = (O.plan_cost_grad(*args, Qp, 0.1, 0.05, 1.0, cost_cfg=cfg)[0] - O.plan_cost_grad(*args, Qm, 0.1, 0.05, 1.0, cost_cfg=cfg)[0]) / (2 * eps)
        np.testing.assert_allclose(G[:, t], fd, rtol=2e-5, atol=1e-6)
