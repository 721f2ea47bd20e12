"""GPU parity of the adjoint kernel (cps_plan_cost_grad: d predict_and_cost / dQ) and of the RPGD gradient step
(cps_rpgd_grad_step) against torch autograd through the unmodified reference modules (tests/golden/grad_*.npz) and the
numpy restatement (oracle.plan_cost_grad)."""
import numpy as np
import pytest

from tests.parity import load_golden, record

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GRAD = ["grad_gradmin_K16_T35", "grad_gradmin_K64_T20", "grad_gradmin_down_K32_T50"]


def _engine(K, T, tp=0.0, te=1.0, integ="ODE", cost="quadratic_boundary_grad_minimal"):
    from cartpolesimulation_b200.core import Engine
    eng = Engine(K, T, integrator=integ, cost=cost, device=0)
    eng.set_variable_parameters(tp, te)
    return eng


@pytest.mark.parametrize("name", GRAD)
@pytest.mark.parametrize("layout", ["rollout_major", "time_major"])
def test_gradient_matches_autograd_through_the_reference(name, layout):
    from cartpolesimulation_b200 import _lib as L
    z, m = load_golden(name)
    K, T = m["K"], m["T"]
    eng = _engine(K, T, m["target_position"], m["target_equilibrium"])
    Q = z["Q"] if layout == "rollout_major" else np.ascontiguousarray(z["Q"].T)
    J, G = eng.plan_cost_grad(torch.from_numpy(z["s"]).cuda(), torch.from_numpy(Q.copy()).cuda(),
                              L.ROLLOUT_MAJOR if layout == "rollout_major" else L.TIME_MAJOR, m["u_prev"])
    G = G.cpu().numpy() if layout == "rollout_major" else G.cpu().numpy().T
    eJ = float(np.abs(J.cpu().numpy() - z["J"]).max() / np.abs(z["J"]).max())
    eG = float(np.abs(G - z["G"]).max() / np.abs(z["G"]).max())
    record("gradient_vs_reference_autograd", f"{name}/{layout}", J=eJ, G=eG)
    assert eJ < 1e-5
    assert eG < 5e-5      # measured <= 1.5e-5 of the largest entry: two float32 realisations of a 350..500-substep adjoint
    assert eng.nonfinite_costs() == 0


@pytest.mark.parametrize("K,T", [(1, 1), (33, 7), (2000, 50), (5000, 20)])
def test_gradient_vs_oracle_sizes(K, T):
    from cartpolesimulation_b200 import _lib as L
    from oracle import oracle as O
    rng = np.random.default_rng(K + T)
    a = np.pi - 0.4
    s = np.array([a, 0.3, np.cos(a), np.sin(a), 0.05, -0.1], dtype=np.float32)
    Q = np.clip(rng.normal(0, 0.5, (K, T)), -1, 1).astype(np.float32)
    eng = _engine(K, T, 0.03, 1.0)
    J, G = eng.plan_cost_grad(torch.from_numpy(s).cuda(), torch.from_numpy(Q).cuda(), L.ROLLOUT_MAJOR, 0.0)
    Jr, Gr = O.plan_cost_grad("quadratic_boundary_grad_minimal", s, Q, 0.0, 0.03, 1.0)
    assert np.abs(J.cpu().numpy() - Jr).max() <= 1e-5 * np.abs(Jr).max()
    assert np.abs(G.cpu().numpy() - Gr).max() <= 5e-5 * max(np.abs(Gr).max(), 1e-3)
    # the costs are those of the forward-only planner kernel (rotation substeps) to float32 noise
    Jp = eng.plan_cost(torch.from_numpy(s).cuda(), torch.from_numpy(Q).cuda(), L.ROLLOUT_MAJOR, 0.0)[0].cpu().numpy()
    assert np.abs(J.cpu().numpy() - Jp).max() <= 1e-5 * np.abs(Jp).max()


def test_rpgd_grad_step_is_clip_adam_clip():
    """grad_step (optimizer_rpgd_tf.py:166-180): gradient -> clip_by_norm -> Adam (Keras legacy) -> clip to the limits,
    three consecutive steps against a numpy restatement fed with the oracle's gradients."""
    from oracle import oracle as O
    z, m = load_golden("grad_gradmin_K64_T20")
    K, T = m["K"], m["T"]
    eng = _engine(K, T, m["target_position"], m["target_equilibrium"])
    eng.rpgd_reset()
    lr, b1, b2, eps, clip = 0.05, 0.9, 0.999, 1e-8, 5.0
    Q = torch.from_numpy(z["Q"].copy()).cuda()
    s = torch.from_numpy(z["s"]).cuda()
    Qr = z["Q"].astype(np.float64)
    mm, vv = np.zeros_like(Qr), np.zeros_like(Qr)
    for it in range(1, 4):
        Jd = torch.empty(K, device="cuda")
        eng.rpgd_grad_step(s, Q, 0.0, lr, b1, b2, eps, clip, J_out=Jd)
        Jr, G = O.plan_cost_grad(m["cost"], z["s"], Qr, 0.0, m["target_position"], m["target_equilibrium"])
        nrm = np.sqrt((G ** 2).sum(axis=1, keepdims=True))
        G = G * clip / np.maximum(nrm, clip)
        mm = b1 * mm + (1 - b1) * G
        vv = b2 * vv + (1 - b2) * G * G
        lr_t = lr * np.sqrt(1 - b2 ** it) / (1 - b1 ** it)
        Qr = np.clip(Qr - lr_t * mm / (np.sqrt(vv) + eps), -1.0, 1.0)
        np.testing.assert_allclose(Q.cpu().numpy(), Qr, rtol=0, atol=2e-5)
        assert np.abs(Jd.cpu().numpy() - Jr).max() <= 2e-5 * np.abs(Jr).max()
    m_dev, v_dev, iters = eng.rpgd_adam_state()
    assert iters == 3
    np.testing.assert_allclose(m_dev.cpu().numpy(), mm, rtol=0, atol=1e-4 * np.abs(mm).max())
    eng.rpgd_reset()
    assert eng.rpgd_adam_state()[2] == 0 and float(eng.rpgd_adam_state()[0].abs().max()) == 0.0


def test_gradient_rejects_other_configurations():
    from cartpolesimulation_b200 import _lib as L
    for integ, cost in (("ODE_v0", "quadratic_boundary_grad_minimal"), ("ODE", "quadratic_boundary")):
        eng = _engine(8, 5, integ=integ, cost=cost)
        with pytest.raises(NotImplementedError):
            eng.plan_cost_grad(torch.zeros(6, device="cuda"), torch.zeros((8, 5), device="cuda"), L.ROLLOUT_MAJOR, 0.0)
