"""GPU tests of the reference-facing plugin objects (optimizer / predictor / cost wrappers): same names, argument
meaning and error behaviour as the reference's own classes, results against the golden vectors frozen from it."""
import numpy as np
import pytest

from tests.parity import load_golden, traj_err, vec_err

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


class InjectedNormal:
    """Same hook as in the reference harness: replaces optimizer.rng (optimizer_mppi.py:172-174)."""

    def __init__(self, draws):
        self.draws, self.i = list(draws), 0

    def normal(self, shape, dtype=None):
        d = self.draws[self.i]
        self.i += 1
        assert list(d.shape) == list(shape)
        return d


def _make_optimizer(m, logging=False, **kw):
    import cartpolesimulation_b200 as cps
    from cartpolesimulation_b200.optimizer_mppi_b200 import optimizer_mppi_b200
    vp = cps.VariableParameters(target_position=m["target_position"], target_equilibrium=m["target_equilibrium"],
                                L=0.395, m_pole=0.087)
    cost = cps.CostFunctionWrapper()
    predictor = cps.PredictorWrapper()
    opt = optimizer_mppi_b200(predictor=predictor, cost_function=cost,
                              control_limits=(np.array([-1.0], np.float32), np.array([1.0], np.float32)),
                              computation_library=None, seed=1, cc_weight=m["cc_weight"], R=m["R"], LBD=m["LBD"],
                              mpc_horizon=m["T"], num_rollouts=m["K"], NU=m["NU"], SQRTRHOINV=m["SQRTRHOINV"],
                              period_interpolation_inducing_points=m["p"], optimizer_logging=logging,
                              calculate_optimal_trajectory=False, **kw)
    predictor.configure(batch_size=m["K"], horizon=m["T"], dt=m["dt"], variable_parameters=vp,
                        predictor_specification=m["predictor"])
    cost.configure(batch_size=m["K"], horizon=m["T"], variable_parameters=vp, environment_name="CartPole",
                   computation_library=None, cost_function_specification=m["cost"])
    opt.configure(dt=m["dt"], predictor_specification=m["predictor"], num_states=predictor.num_states,
                  num_control_inputs=predictor.num_control_inputs)
    return opt, vp


@pytest.mark.parametrize("run", ["ode_gradmin", "v0_gradmin", "ode_grad_down", "ode_gradmin_T100"])
@pytest.mark.parametrize("logging", [False, True])
def test_optimizer_step_sequence_matches_reference(run, logging):
    """controller_mpc-style use: consecutive optimizer.step(s) calls with the reference's injected draws; the
    optimizer carries u_nom and u_prev itself (warm-start shift, last returned control)."""
    z, m = load_golden("mppi_" + run)
    opt, _ = _make_optimizer(m, logging=logging)
    opt.rng = InjectedNormal([torch.from_numpy(e[:, :, None].copy()) for e in z["eps"]])
    for i in range(m["steps"]):
        u = opt.step(z["s"][i].copy())
        assert isinstance(u, np.ndarray) and u.dtype == np.float32 and u.shape == ()
        assert abs(float(u) - float(z["u"][i])) < 1e-4, (i, float(u), float(z["u"][i]))
        np.testing.assert_allclose(opt.u_nom.numpy().reshape(-1), z["u_nom"][i], rtol=0, atol=1e-4)
        assert opt.u_nom.shape == (1, m["T"], 1)
        if logging:
            lv = opt.logging_values
            assert lv["Q_logged"].shape == (m["K"], m["T"], 1)
            assert lv["rollout_trajectories_logged"].shape == (m["K"], m["T"] + 1, 6)
            assert vec_err(lv["J_logged"], z["J"][i]) < 3e-5
            np.testing.assert_array_equal(lv["s_logged"], z["s"][i])
            if i == 0:
                np.testing.assert_allclose(lv["Q_logged"][:, :, 0], z["u_run0"], rtol=0, atol=3e-7)
    assert opt.optimizer_name == "mppi-b200"
    opt.optimizer_reset()
    np.testing.assert_array_equal(opt.u_nom.numpy().reshape(-1), np.zeros(m["T"], np.float32))


def test_optimizer_own_rng_is_seeded_and_variable_parameters_are_reread():
    z, m = load_golden("mppi_ode_gradmin")
    m = dict(m, K=512)
    a, vp_a = _make_optimizer(m)
    b, vp_b = _make_optimizer(m)
    s = z["s"][0]
    ua = [float(a.step(s)) for _ in range(3)]
    ub = [float(b.step(s)) for _ in range(3)]
    assert ua == ub  # same seed -> same device draws -> bitwise same controls
    assert all(abs(x) <= 1.0 for x in ua)
    # target change is picked up on the next step without reconfiguring (CartPole/__init__.py:512-519)
    vp_b.update_attributes({"target_position": 0.15, "target_equilibrium": -1.0})
    vp_a.update_attributes({"target_position": 0.15, "target_equilibrium": -1.0})
    assert float(a.step(s)) == float(b.step(s))
    c, _ = _make_optimizer(m)
    for _ in range(3):
        c.step(s)
    assert float(c.step(s)) != float(a.u)  # c still tracks the old target


def test_optimal_trajectory_and_errors():
    z, m = load_golden("mppi_ode_gradmin")
    import cartpolesimulation_b200 as cps
    from cartpolesimulation_b200.optimizer_mppi_b200 import optimizer_mppi_b200
    opt, _ = _make_optimizer(dict(m, K=256))
    opt.calculate_optimal_trajectory = True
    opt.step(z["s"][0])
    assert opt.optimal_trajectory.shape == (1, m["T"] + 1, 6)
    np.testing.assert_array_equal(opt.optimal_trajectory[0, 0], z["s"][0])
    assert opt.optimal_control_sequence.shape == (1, m["T"], 1)
    with pytest.raises(ValueError):
        opt.step(np.zeros(5, np.float32))
    bad = optimizer_mppi_b200(predictor="GP", cost_function="default", control_limits=([-1.0], [1.0]),
                              mpc_horizon=10, num_rollouts=32)
    with pytest.raises(ValueError):
        bad.configure(num_states=6, num_control_inputs=1, dt=0.02, predictor_specification="GP")
    with pytest.raises(ValueError):
        bad.configure(num_states=4, num_control_inputs=2, dt=0.02, predictor_specification="ODE")
    with pytest.raises(RuntimeError):
        optimizer_mppi_b200(predictor="ODE", cost_function="default", control_limits=([-1.0], [1.0])).step(z["s"][0])


@pytest.mark.parametrize("integ,fname", [("ODE_v0", "rollout_ode_v0"), ("ODE", "rollout_ode")])
def test_predictor_wrapper_interface(integ, fname):
    import cartpolesimulation_b200 as cps
    z, meta = load_golden(fname)
    s0, Q, ref = z["tiled__s0"], z["tiled__Q"], z["tiled__traj"]
    B, T = Q.shape
    pw = cps.PredictorWrapper()
    pw.configure(batch_size=B, horizon=T, dt=meta["dt"], predictor_specification=integ)
    assert pw.num_states == 6 and pw.num_control_inputs == 1 and pw.predictor_type == integ
    # predict_core: torch cuda in -> torch cuda out, numpy in -> numpy out, torch cpu in -> torch cpu out
    out_np = pw.predict_core(np.tile(s0, (B, 1)), Q[:, :, None])
    assert isinstance(out_np, np.ndarray) and out_np.shape == (B, T + 1, 6)
    assert max(traj_err(out_np, ref).values()) < 1e-5
    out_t = pw.predict_core(torch.from_numpy(s0).cuda(), torch.from_numpy(Q[:, :, None]).cuda())
    assert out_t.is_cuda
    np.testing.assert_array_equal(out_t.cpu().numpy(), out_np)
    out_c = pw.predict_core(torch.from_numpy(np.tile(s0, (B, 1))), torch.from_numpy(Q[:, :, None]))
    assert not out_c.is_cuda
    # predict: numpy, tolerant shapes
    p = pw.predict(s0[0], Q[:, :, None])
    np.testing.assert_array_equal(p, out_np)
    one = pw.predict(s0[0], Q[0, :, None])
    if integ == "ODE_v0":
        assert one.shape == (T + 1, 6)  # squeezed for batch 1 (predictor_ODE_v0.py:74)
        with pytest.raises(ValueError):
            pw.predict(np.tile(s0, (3, 1)), Q[:, :, None])
    else:
        assert one.shape == (1, T + 1, 6)
    np.testing.assert_array_equal(np.squeeze(one), out_np[0])
    pw.update(Q0=Q[:, :1, None], s=s0)  # no-op for ODE predictors
    c = pw.copy()
    assert c.predictor_type == integ and c.predictor is None
    with pytest.raises(ValueError):
        cps.PredictorWrapper().configure(batch_size=1, horizon=5, dt=0.02, predictor_specification="nonsense")


@pytest.mark.parametrize("name", ["default", "quadratic_boundary", "quadratic_boundary_grad_minimal",
                                  "quadratic_boundary_grad"])
def test_cost_wrapper_interface(name):
    import cartpolesimulation_b200 as cps
    z, meta = load_golden("costs")
    traj, Q = z["traj"], z["Q"]
    K, T = Q.shape
    tp, te, up = z["settings"][1]
    vp = cps.VariableParameters(target_position=torch.tensor(tp, dtype=torch.float32),
                                target_equilibrium=torch.tensor(te, dtype=torch.float32))
    cw = cps.CostFunctionWrapper()
    cw.configure(batch_size=K, horizon=T, variable_parameters=vp, environment_name="CartPole",
                 computation_library=None, cost_function_specification=name.replace("_", "-"))
    assert cw.cost_function_name == name
    tt, qq = torch.from_numpy(traj).cuda(), torch.from_numpy(Q[:, :, None]).cuda()
    J = cw.get_trajectory_cost(tt, qq, np.float32(up))
    st = cw.get_stage_cost(tt[:, :-1, :], qq, np.float32(up))
    term = cw.get_terminal_cost(tt[:, -1, :])
    assert J.shape == (K,) and st.shape == (K, T) and term.shape == (K, 1) and J.is_cuda
    assert vec_err(J.cpu().numpy(), z[f"{name}__1__J"]) < 2e-6
    np.testing.assert_array_equal(term.cpu().numpy().reshape(-1), z[f"{name}__1__terminal"])
    if name in ("default", "quadratic_boundary"):
        assert cw.cost_function.MAX_COST == float(z[f"{name}__max_cost"])
    else:
        assert vec_err(st.cpu().numpy(), z[f"{name}__1__stage"]) < 2e-6
        summed = cw.get_summed_stage_cost(tt, qq, np.float32(up))
        np.testing.assert_allclose(summed.cpu().numpy(), z[f"{name}__1__stage"].sum(1), rtol=1e-5)
    # numpy in -> numpy out
    Jn = cw.get_trajectory_cost(traj, Q[:, :, None], up)
    assert isinstance(Jn, np.ndarray)
    np.testing.assert_array_equal(Jn, J.cpu().numpy())
    # hot reload contract (cost_function_wrapper.py:71-74)
    cw.cost_function.reload_cost_parameters_from_config_flag = True
    cw.update_cost_parameters_from_config()
    assert cw.cost_function.reload_cost_parameters_from_config_flag is False
    with pytest.raises(ValueError):
        cps.CostFunctionWrapper().configure(batch_size=1, horizon=1, variable_parameters=vp,
                                            cost_function_specification="quadratic_boundary_nonconvex")


# ------------------------------------------------------------------------------------------------------------
# neural predictor through the plugin objects
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["net_GRU_6IN_64H1_64H2_5OUT_0", "net_Dense_6IN_32H1_32H2_5OUT_0"])
def test_neural_predictor_wrapper_interface(name, tmp_path):
    """PredictorWrapper -> predictor_autoregressive_neural loaded from a model directory in the reference's format
    (net-info .txt, torch checkpoint, normalisation csv): predict_core / update / predict vs the reference's outputs."""
    import cartpolesimulation_b200 as cps
    from tests.netutil import write_model_dir
    z, m = load_golden(name)
    model = write_model_dir(str(tmp_path), z)
    K, T = z["Q"].shape
    pw = cps.PredictorWrapper()
    pw.configure(batch_size=K, horizon=T, dt=0.02, predictor_specification=model)
    assert pw.predictor_type == "neural" and pw.num_states == 6 and pw.num_control_inputs == 1
    pw.predictor.update_before_predicting = False
    s0 = np.tile(z["s0"], (K, 1))
    Q = z["Q"][:, :, None]
    out = pw.predict_core(torch.from_numpy(s0), torch.from_numpy(Q))     # torch CPU in -> torch CPU out
    assert isinstance(out, torch.Tensor) and out.device.type == "cpu" and tuple(out.shape) == (K, T + 1, 6)
    assert max(traj_err(out.numpy(), z["traj_zero_h"]).values()) < 1e-5
    s_cur = s0.copy()
    for i in range(3):  # PredictorWrapper.update -> update_internal_state_tf (predictor_wrapper.py:173-177)
        pw.update(Q0=np.full((K, 1, 1), z["upd_q"][i], np.float32), s=s_cur)
        s_cur = s_cur.copy()
        s_cur[:, 1] += 0.05
    if m["type"] == "GRU":
        assert np.abs(pw.predictor.engine.net_get_state() - z["h_after_updates"].reshape(-1)).max() < 2e-6
    out2 = pw.predict(s_cur, Q)                                            # numpy in -> numpy out
    assert isinstance(out2, np.ndarray)
    assert max(traj_err(out2, z["traj_after_updates"]).values()) < 1e-5
    dev = pw.predictor.device
    out3 = pw.predict_core(torch.from_numpy(z["s_rand"]).to(dev), torch.from_numpy(Q).to(dev))  # CUDA in -> CUDA out
    assert out3.device == dev
    assert max(traj_err(out3.cpu().numpy(), z["traj_rand"]).values()) < 1e-5
    with pytest.raises(ValueError):
        pw.predict_core(torch.from_numpy(s0), torch.from_numpy(z["Q"]))    # Q without the feature axis


def test_neural_predictor_errors(tmp_path):
    import cartpolesimulation_b200 as cps
    pw = cps.PredictorWrapper()
    with pytest.raises(FileNotFoundError):
        pw.configure(batch_size=4, horizon=5, dt=0.02, predictor_specification=str(tmp_path / "GRU-6IN-8H1-5OUT-0"))
    with pytest.raises(NotImplementedError):
        pw.configure(batch_size=4, horizon=5, dt=0.02, predictor_specification="GP")


@pytest.mark.parametrize("run", ["gru64_gradmin", "dense32_gradmin", "gru32_grad"])
def test_optimizer_neural_closed_loop_matches_reference(run, tmp_path):
    """optimizer_mppi_b200 + neural predictor, the same closed loop the golden was recorded on (injected draws); the
    solver carries its own u_nom, u_prev and hidden state from step to step."""
    from tests.netutil import write_model_dir
    z, m = load_golden("mppi_net_" + run)
    m = dict(m)
    m["predictor"] = write_model_dir(str(tmp_path), z)
    opt, vp = _make_optimizer(m, logging=(run == "gru64_gradmin"))
    K, n_ind = m["K"], z["eps"].shape[2]
    opt.rng = InjectedNormal([torch.from_numpy(z["eps"][i]).reshape(K, n_ind, 1) for i in range(m["steps"])])
    for i in range(m["steps"]):
        u = opt.step(z["s"][i].copy())
        assert isinstance(u, np.ndarray) and u.dtype == np.float32
        from tests.parity import record
        du, dn = abs(float(u) - float(z["u"][i])), float(np.abs(opt.u_nom.numpy().reshape(-1) - z["u_nom"][i]).max())
        record("optimizer_neural_closed_loop", f"{run}/{i}", u=du, u_nom=dn)
        # closed loop: the solver carries ITS u_nom / hidden state forward, so differences accumulate over the steps
        # (measured <= 3e-5 after 3-4 solves; north_star: 1e-4)
        assert du < 1e-4, (i, float(u), float(z["u"][i]))
        np.testing.assert_allclose(opt.u_nom.numpy().reshape(-1), z["u_nom"][i], rtol=0, atol=1e-4)
        if "h_after" in z.files:
            assert np.abs(opt.engine.net_get_state() - z["h_after"][i]).max() < 1e-5
        if opt.optimizer_logging:
            assert vec_err(opt.logging_values["J_logged"], z["J"][i]) < 2e-5
            if i == 0:
                assert max(traj_err(opt.logging_values["rollout_trajectories_logged"][:32], z["traj0"]).values()) < 1e-5


def test_optimizer_with_reference_shaped_wrappers():
    """The optimizer fed with objects shaped like the REFERENCE's wrappers (not this package's mirrors), real Engine on the
    GPU: the cost plugin is an instance of a class called `quadratic_boundary` whose weights are module-level constants
    (Control_Toolkit_ASF/Cost_Functions/CartPole/quadratic_boundary.py:10-21), the predictor wrapper carries
    predictor_type and predictor_config['intermediate_steps'] (SI_Toolkit/Predictors/predictor_wrapper.py), and
    variable_parameters holds 0-d tensors (Control_Toolkit/others/environment.py / General/variable_parameters.py).
    Checked against the oracle with the same (non-default) weights, substeps, pole length and mass."""
    import sys
    import types
    import torch
    from cartpolesimulation_b200.optimizer_mppi_b200 import optimizer_mppi_b200
    from oracle import oracle as O
    from tests.parity import shifted_cost_ok

    weights = dict(dd_weight=450.0, ep_weight=15000.0, cc_weight=1.5, ccrc_weight=0.5, R=1.0)
    mod = types.ModuleType("Control_Toolkit_ASF.Cost_Functions.CartPole.quadratic_boundary")
    for k, v in weights.items():
        setattr(mod, k, v)

    class quadratic_boundary:     # no .config dict: the weights live in the module, as in the reference
        pass
    quadratic_boundary.__module__ = mod.__name__
    mod.quadratic_boundary = quadratic_boundary
    sys.modules[mod.__name__] = mod
    try:
        vp = types.SimpleNamespace(target_position=torch.tensor(0.05), target_equilibrium=torch.tensor(1.0),
                                   L=torch.tensor(0.3), m_pole=torch.tensor(0.1))
        cost_wrapper = types.SimpleNamespace(cost_function=quadratic_boundary(), variable_parameters=vp)
        cost_wrapper.cost_function.variable_parameters = vp
        predictor_wrapper = types.SimpleNamespace(predictor_type="ODE", predictor_config={"intermediate_steps": 5},
                                                  predictor=types.SimpleNamespace())
        K, T = 1024, 30
        opt = optimizer_mppi_b200(predictor=predictor_wrapper, cost_function=cost_wrapper, control_limits=([-1.0], [1.0]),
                                  seed=3, mpc_horizon=T, num_rollouts=K, optimizer_logging=True)
        opt.configure(num_states=6, num_control_inputs=1, dt=0.02, predictor_specification="ODE")
        assert opt.cost_name == "quadratic_boundary" and opt.predictor_type == "ODE"
        rng = np.random.default_rng(11)
        eps = rng.standard_normal((K, opt.engine.n_ind)).astype(np.float32)

        class Injected:   # the reference's call shape: rng.normal([K, n_ind, 1], dtype)
            def normal(self, shape, dtype=None):
                assert list(shape) == [K, opt.engine.n_ind, 1]
                return torch.from_numpy(eps).reshape(shape)
        opt.rng = Injected()
        s = np.array([2.8, -0.4, np.cos(2.8), np.sin(2.8), 0.03, 0.1], dtype=np.float32)
        u = opt.step(s)
        ref = O.mppi_step("ODE", "quadratic_boundary", s, np.zeros(T, np.float32), eps=eps, u_prev=0.0, target_position=0.05,
                          target_equilibrium=1.0, n=5, phys=O.physics_vector(L=0.3, m_pole=0.1), cost_cfg=weights)
        assert shifted_cost_ok(opt.logging_values["J_logged"], ref["J"])
        assert abs(float(u) - float(ref["u"])) < 1e-5
        np.testing.assert_allclose(opt.u_nom.numpy().reshape(-1), ref["u_nom"], rtol=0, atol=1e-5)
        # the same call with default weights must differ: the module constants really reached the kernel
        ref_default = O.mppi_step("ODE", "quadratic_boundary", s, np.zeros(T, np.float32), eps=eps, u_prev=0.0,
                                  target_position=0.05, n=5, phys=O.physics_vector(L=0.3, m_pole=0.1))
        assert np.abs(ref_default["u_nom"] - ref["u_nom"]).max() > 1e-4
    finally:
        sys.modules.pop(mod.__name__, None)
