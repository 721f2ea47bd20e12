"""GPU parity of the adjoint kernel (cps_plan_cost_grad: d predict_and_cost / dQ) and of the RPGD gradient step
(cps_rpgd_grad_step) against torch autograd through the unmodified reference modules (tests/golden/grad_*.npz) and the
numpy restatement (oracle.plan_cost_grad)."""
import numpy as np
import pytest

from tests.parity import close_except_few, load_golden, record

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GRAD = ["grad_gradmin_K16_T35", "grad_gradmin_K64_T20", "grad_gradmin_down_K32_T50",
        "grad_qbgrad_K16_T35", "grad_qbgrad_terms_K64_T20", "grad_qbgrad_down_K32_T50"]


def _engine(K, T, tp=0.0, te=1.0, integ="ODE", cost="quadratic_boundary_grad_minimal", cost_config=None):
    from cartpolesimulation_b200 import config as cfgmod
    from cartpolesimulation_b200.core import Engine
    eng = Engine(K, T, integrator=integ, cost=cost, device=0)
    eng.set_variable_parameters(tp, te)
    if cost_config:   # non-default plugin weights (config_cost_function.yml values), folded as the plugin folds them
        cfg = dict(cfgmod.DEFAULT_COST_CONFIG[cost])
        cfg.update(cost_config)
        eng.set_cost_params(cfgmod.cost_vector(cost, cfg))
    return eng


@pytest.mark.parametrize("name", GRAD)
@pytest.mark.parametrize("layout", ["rollout_major", "time_major"])
def test_gradient_matches_autograd_through_the_reference(name, layout):
    from cartpolesimulation_b200 import _lib as L
    z, m = load_golden(name)
    K, T = m["K"], m["T"]
    eng = _engine(K, T, m["target_position"], m["target_equilibrium"], cost=m["cost"], cost_config=m.get("cost_config_overrides"))
    Q = z["Q"] if layout == "rollout_major" else np.ascontiguousarray(z["Q"].T)
    J, G = eng.plan_cost_grad(torch.from_numpy(z["s"]).cuda(), torch.from_numpy(Q.copy()).cuda(),
                              L.ROLLOUT_MAJOR if layout == "rollout_major" else L.TIME_MAJOR, m["u_prev"])
    G = G.cpu().numpy() if layout == "rollout_major" else G.cpu().numpy().T
    eJ = float(np.abs(J.cpu().numpy() - z["J"]).max() / np.abs(z["J"]).max())
    eG = float(np.abs(G - z["G"]).max() / np.abs(z["G"]).max())
    record("gradient_vs_reference_autograd", f"{name}/{layout}", J=eJ, G=eG)
    assert eJ < (2e-5 if name == "grad_qbgrad_down_K32_T50" else 1e-5)   # that golden's own float32 noise is 1.1e-5 (test_oracle_grad.py)
    assert eG < 5e-5      # measured <= 2.5e-5 of the largest entry: two float32 realisations of a 350..500-substep adjoint
    assert eng.nonfinite_costs() == 0


# (8192, 20): more (step, plan) pairs than the fused Jacobian + reverse kernel takes -> records in global memory; (8, 600): a
# horizon beyond its 512 steps
@pytest.mark.parametrize("K,T", [(1, 1), (33, 7), (2000, 50), (5000, 20), (8192, 20), (8, 600)])
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


@pytest.mark.parametrize("te", [1.0, -1.0])
def test_gradient_quadratic_boundary_grad_vs_oracle(te):
    """quadratic_boundary_grad with its linear-distance and control-change terms on, both target equilibria, u_prev != 0."""
    from cartpolesimulation_b200 import _lib as L
    from oracle import oracle as O
    K, T = 300, 25
    rng = np.random.default_rng(11)
    a = 2.2
    s = np.array([a, -0.4, np.cos(a), np.sin(a), 0.16, 0.2], dtype=np.float32)
    Q = np.clip(rng.normal(0, 0.6, (K, T)), -1, 1).astype(np.float32)
    sfx = "_up" if te == 1.0 else "_down"
    over = {"dd_linear_weight" + sfx: 25.0, "ccrc_weight" + sfx: 1.5}
    eng = _engine(K, T, -0.02, te, cost="quadratic_boundary_grad", cost_config=over)
    J, G = eng.plan_cost_grad(torch.from_numpy(s).cuda(), torch.from_numpy(Q).cuda(), L.ROLLOUT_MAJOR, -0.3)
    cfg = dict(O.DEFAULT_COST_CONFIG["quadratic_boundary_grad"])
    cfg.update(over)
    Jr, Gr = O.plan_cost_grad("quadratic_boundary_grad", s, Q, -0.3, -0.02, te, cost_cfg=cfg)
    eJ = float(np.abs(J.cpu().numpy() - Jr).max() / np.abs(Jr).max())
    eG = float(np.abs(G.cpu().numpy() - Gr).max() / np.abs(Gr).max())
    record("gradient_qb_grad_vs_oracle", f"te{te:+.0f}", J=eJ, G=eG)
    assert eJ < 1e-5 and eG < 5e-5


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


RPGD = ["plan_rpgd_default", "plan_rpgd_resamp3", "plan_rpgd_qbgrad"]


class _Draws:
    def __init__(self, draws):
        self.draws, self.i = list(draws), 0

    def normal(self, shape, mean=0.0, stddev=1.0, dtype=None):
        d = self.draws[self.i]
        self.i += 1
        assert list(d.shape) == list(shape)
        return d * stddev + mean


def _make_rpgd(m, draws, logging=False):
    import cartpolesimulation_b200 as cps
    from cartpolesimulation_b200.optimizer_forward_b200 import optimizer_rpgd_b200
    vp = cps.VariableParameters(target_position=m["target_position"], target_equilibrium=m["target_equilibrium"], L=0.395,
                                m_pole=0.087)
    cost, predictor = cps.CostFunctionWrapper(), cps.PredictorWrapper()
    opt = optimizer_rpgd_b200(predictor=predictor, cost_function=cost,
                              control_limits=(np.array([-1.0], np.float32), np.array([1.0], np.float32)), seed=1,
                              mpc_horizon=m["T"], num_rollouts=m["K"], outer_its=m["outer_its"], resamp_per=m["resamp_per"],
                              opt_keep_k_ratio=m["opt_keep_k_ratio"], optimizer_logging=logging,
                              period_interpolation_inducing_points=m["period_interpolation_inducing_points"],
                              learning_rate=m["learning_rate"], gradmax_clip=m["gradmax_clip"], sample_stdev=m["sample_stdev"])
    predictor.configure(batch_size=m["K"], horizon=m["T"], dt=0.02, variable_parameters=vp, predictor_specification="ODE")
    cost.configure(batch_size=m["K"], horizon=m["T"], variable_parameters=vp, environment_name="CartPole",
                   computation_library=None, cost_function_specification=m["cost"])
    opt.configure(num_states=6, num_control_inputs=1, default_configure=False, dt=0.02, predictor_specification="ODE")
    opt.rng = _Draws(draws)
    opt.optimizer_reset()
    return opt


@pytest.mark.parametrize("name", RPGD)
def test_optimizer_rpgd_b200_matches_reference(name):
    """optimizer_rpgd_b200.step against recordings of the UNMODIFIED optimizer_rpgd_tf (tests/golden/plan_rpgd_*.npz): every
    solve starts from the reference's recorded state (plans, Adam moments, counters), so a near-tie in the cost order cannot
    compound; compared: the plans after the gradient steps, the control, the warm start with its resampled rows, the
    re-sorted and shifted Adam moments."""
    z, m = load_golden(name)
    K, T = m["K"], m["T"]
    draws = [torch.from_numpy(z["draw0"][:, :, None].copy())] + [torch.from_numpy(d[:, :, None].copy()) for d in z["resamp_draws"]]
    opt = _make_rpgd(m, draws, logging=True)
    W = opt._W.cpu().numpy()
    np.testing.assert_allclose(opt.Q_tf.cpu().numpy(), np.clip(z["draw0"] * m["sample_stdev"], -1, 1) @ W, rtol=0, atol=1e-6)
    for i in range(m["steps"]):
        u = opt.step(z["s"][i].copy())
        assert isinstance(u, np.ndarray) and u.dtype == np.float32
        Q_opt = opt.logging_values["Q_logged"][:, :, 0]
        eQ, eu = float(np.abs(Q_opt - z["Q"][i]).max()), abs(float(u) - float(z["u"][i]))
        record("optimizer_rpgd", f"{name}/{i}", Q=eQ, u=eu, Q_entries_over_2e_4=int((np.abs(Q_opt - z["Q"][i]) > 2e-4).sum()))
        # quadratic_boundary_grad: a plan entry whose gradient is at rounding level takes an ill-conditioned Adam step
        # (g / (|g| + eps)); the float64 restatement shows the same single entry (tests/test_oracle_plan.py)
        few = dict(max_outliers=2, outlier_atol=2e-3) if name == "plan_rpgd_qbgrad" else {}
        assert close_except_few(Q_opt, z["Q"][i], 2e-4, **few), eQ
        assert eu < 1e-4                                              # north_star: selected control within 1e-4
        assert close_except_few(opt.Q_tf.cpu().numpy(), z["Q_next"][i], 2e-4, **few)
        mm, vv, it = opt.engine.rpgd_adam_state()
        assert it == int(z["adam_iterations"][i])
        np.testing.assert_allclose(mm.cpu().numpy(), z["adam_m"][i], rtol=0, atol=2e-4 * max(1.0, np.abs(z["adam_m"][i]).max()))
        np.testing.assert_allclose(vv.cpu().numpy(), z["adam_v"][i], rtol=0, atol=2e-4 * max(1.0, np.abs(z["adam_v"][i]).max()))
        # restart the next solve from the reference's own state
        opt.Q_tf.copy_(torch.from_numpy(z["Q_next"][i]).cuda())
        mm.copy_(torch.from_numpy(z["adam_m"][i]).cuda())
        vv.copy_(torch.from_numpy(z["adam_v"][i]).cuda())
    assert opt.optimizer_name == "rpgd-b200" and opt.count == m["steps"]


def test_optimizer_rpgd_b200_free_running_and_own_rng():
    """Free running (its own state from solve to solve) the control follows the reference over the first solves, and with
    its own seeded generator two instances agree; a solve lowers the cost of the plans it starts from."""
    z, m = load_golden("plan_rpgd_default")
    draws = [torch.from_numpy(z["draw0"][:, :, None].copy())] + [torch.from_numpy(d[:, :, None].copy()) for d in z["resamp_draws"]]
    opt = _make_rpgd(m, draws)
    for i in range(4):
        assert abs(float(opt.step(z["s"][i].copy())) - float(z["u"][i])) < 2e-3
    import cartpolesimulation_b200 as cps
    from cartpolesimulation_b200 import _lib as L
    a, b = _make_rpgd(m, draws), _make_rpgd(m, draws)
    for o in (a, b):
        o.rng = o._own_rng
        o.optimizer_reset()
    s = z["s"][0]
    J0 = a.engine.plan_cost(torch.from_numpy(s).cuda(), a.Q_tf, L.ROLLOUT_MAJOR, 0.0)[0].cpu().numpy()
    ua, ub = [float(a.step(s)) for _ in range(3)], [float(b.step(s)) for _ in range(3)]
    assert ua == ub and all(abs(v) <= 1.0 for v in ua)
    b2 = _make_rpgd(m, draws)
    b2.rng = b2._own_rng
    b2.optimizer_reset()
    Q_start = b2.Q_tf.clone()
    for _ in range(m["outer_its"]):
        b2.engine.rpgd_grad_step(torch.from_numpy(s).cuda(), b2.Q_tf)
    J1 = b2.engine.plan_cost(torch.from_numpy(s).cuda(), b2.Q_tf, L.ROLLOUT_MAJOR, 0.0)[0].cpu().numpy()
    J0 = b2.engine.plan_cost(torch.from_numpy(s).cuda(), Q_start, L.ROLLOUT_MAJOR, 0.0)[0].cpu().numpy()
    assert (J1 < J0).mean() > 0.8


@pytest.mark.parametrize("K,T,resample", [(16, 35, True), (16, 35, False), (37, 9, True), (2000, 50, True), (4096, 5, False)])
def test_rpgd_finish_matches_the_tensor_restatement(K, T, resample):
    """cps_rpgd_finish = get_action + the bookkeeping of optimizer_rpgd_tf.step (:182-224, :297-356), here against the same
    steps written with torch operations (stable sort, gather, shift, concatenation); costs with ties."""
    eng = _engine(K, T)
    eng.rpgd_reset()
    g = torch.Generator(device="cuda").manual_seed(K * 100 + T)
    J = torch.randint(0, max(2, K // 3), (K,), generator=g, device="cuda").float()   # many equal costs: the order must be stable
    Q = torch.rand((K, T), generator=g, device="cuda") * 2 - 1
    m, v, _ = eng.rpgd_adam_state()
    m.copy_(torch.rand((K, T), generator=g, device="cuda"))
    v.copy_(torch.rand((K, T), generator=g, device="cuda"))
    ages = torch.randint(0, 50, (K,), generator=g, device="cuda", dtype=torch.int32)
    keep, sp = max(1, (3 * K) // 4), 1
    fresh = (torch.rand((K - keep, T), generator=g, device="cuda") - 0.5) if resample else None
    # restatement
    best = torch.sort(J, stable=True).indices[:keep]
    u_nom_ref = Q[best[0]].clone()
    Qn = torch.cat([Q[:, sp:], Q[:, -1:].repeat(1, sp)], dim=1)
    zeros = torch.zeros((K, 1), device="cuda")
    if resample:
        Qn = torch.cat([fresh, Qn[best]], dim=0)
        ages_ref = torch.cat([torch.zeros(K - keep, dtype=torch.int32, device="cuda"), ages[best]]) + 1
        z = torch.zeros((K - keep, T), device="cuda")
        m_ref = torch.cat([z, torch.cat([m[best][:, 1:], zeros[:keep]], dim=1)], dim=0)
        v_ref = torch.cat([z, torch.cat([v[best][:, 1:], zeros[:keep]], dim=1)], dim=0)
    else:
        ages_ref = ages + 1
        m_ref = torch.cat([m[:, 1:], zeros], dim=1)
        v_ref = torch.cat([v[:, 1:], zeros], dim=1)
    u_nom = eng.rpgd_finish(J, Q, fresh, keep, sp, ages)
    m2, v2, _ = eng.rpgd_adam_state()
    assert np.array_equal(u_nom, u_nom_ref.cpu().numpy())
    assert torch.equal(Q, Qn) and torch.equal(m2, m_ref) and torch.equal(v2, v_ref) and torch.equal(ages, ages_ref)


def test_rpgd_finish_rejects_large_batches():
    eng = _engine(5000, 4)
    with pytest.raises(NotImplementedError):
        eng.rpgd_finish(torch.zeros(5000, device="cuda"), torch.zeros((5000, 4), device="cuda"), None, 3750, 1, None)
