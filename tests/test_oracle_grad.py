"""Pins the hand-derived adjoint of predict_and_cost (oracle.plan_cost_grad: reverse sweep of the Euler-Cromer substep and the
quadratic_boundary_grad_minimal plugin) against torch autograd through the UNMODIFIED reference modules
(tests/golden/grad_*.npz, oracle/gen_golden_grad.py) -- the derivative the reference's RPGD takes with a GradientTape
(Control_Toolkit/Optimizers/optimizer_rpgd_tf.py:167-175)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.parity import load_golden

GRAD = ["grad_gradmin_K16_T35", "grad_gradmin_K64_T20", "grad_gradmin_down_K32_T50"]


@pytest.mark.parametrize("name", GRAD)
def test_adjoint_matches_autograd_through_the_reference(name):
    z, m = load_golden(name)
    J, G = O.plan_cost_grad(m["cost"], z["s"], z["Q"], m["u_prev"], m["target_position"], m["target_equilibrium"])
    assert np.abs(J - z["J"]).max() <= 1e-5 * np.abs(z["J"]).max()
    # float64 restatement against the reference's float32 autograd: its rounding noise is ~1e-5 of the largest entry
    assert np.abs(G - z["G"]).max() <= 2e-5 * np.abs(z["G"]).max()


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
