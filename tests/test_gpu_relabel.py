"""GPU parity of offline relabelling: cps_fleet_relabel and the add_control_along_trajectories mirror
against recordings of the UNMODIFIED reference function driving optimizer_mppi (tests/golden/relabel_*.npz) and
against the CPU oracle."""
import numpy as np
import pytest

from tests.parity import load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

STATE = ["angle", "angleD", "angle_cos", "angle_sin", "position", "positionD"]


def _stack(z, m, key, dtype=np.float32):
    return np.stack([z[f"f{f}__{key}"] for f in range(m["files"])], axis=1).astype(dtype)   # [calls, E, ...]


@pytest.mark.parametrize("name", ["relabel_plain_ode", "relabel_plain_v0", "relabel_integrate_ode",
                                  "relabel_differentiate_ode"])
def test_relabel_matches_reference(name):
    from cartpolesimulation_b200.relabel import Relabeller
    z, m = load_golden(name)
    E = m["files"]
    rl = Relabeller(E, m["K"], m["T"], integrator=m["predictor"], cost=m["cost"], noise="supplied", device=0)
    eps = np.ascontiguousarray(_stack(z, m, "eps").transpose(0, 1, 3, 2))   # [calls, E, K, n_ind] -> [calls, E, n_ind, K]
    rl.reset()
    Q = rl.relabel(_stack(z, m, "s"), _stack(z, m, "tp"), _stack(z, m, "te"), _stack(z, m, "L"),
                   noise=torch.from_numpy(eps).cuda(), chunk_rows=7)   # chunking must not matter
    ref = _stack(z, m, "u")
    np.testing.assert_allclose(Q, ref, rtol=0, atol=1e-4)   # north_star: selected control within 1e-4
    ev = max(m["evals"], 1)
    for f in range(E):
        if "differentiate" in name:   # central output of the five-point window; the derivative amplifies 1e-4 by 1/step
            np.testing.assert_allclose(Q[:, f].reshape(m["rows"], 5)[:, 2], z[f"f{f}___calculated_offline_dL"], rtol=0, atol=1e-4)
            continue
        label = Q[:, f].reshape(m["rows"], ev).astype(np.float64).mean(axis=1)
        np.testing.assert_allclose(label, z[f"f{f}__Q_calculated_offline"], rtol=0, atol=1e-4)
    # a second pass after reset() reproduces the first bit for bit (warm start and last control were cleared)
    rl.reset()
    Q2 = rl.relabel(_stack(z, m, "s"), _stack(z, m, "tp"), _stack(z, m, "te"), _stack(z, m, "L"),
                    noise=torch.from_numpy(eps).cuda())
    np.testing.assert_array_equal(Q, Q2)


def test_relabel_lockstep_equals_one_file_at_a_time():
    """Each file through optimizer_mppi_b200.step row by row (what the reference's loop would do with the drop-in
    optimizer) gives the lockstep result."""
    import cartpolesimulation_b200 as cps
    from cartpolesimulation_b200.relabel import Relabeller
    from tests.test_gpu_plugin import InjectedNormal, _make_optimizer
    z, m = load_golden("relabel_plain_ode")
    E = m["files"]
    rl = Relabeller(E, m["K"], m["T"], integrator="ODE", cost=m["cost"], noise="supplied", device=0)
    eps = np.ascontiguousarray(_stack(z, m, "eps").transpose(0, 1, 3, 2))
    Q = rl.relabel(_stack(z, m, "s"), _stack(z, m, "tp"), _stack(z, m, "te"), _stack(z, m, "L"),
                   noise=torch.from_numpy(eps).cuda())
    mm = dict(K=m["K"], T=m["T"], predictor="ODE", cost=m["cost"], target_position=0.0, target_equilibrium=1.0,
              cc_weight=1.0, R=1.0, LBD=100.0, NU=1000.0, SQRTRHOINV=0.03, p=10, dt=0.02)
    for f in range(E):
        opt, vp = _make_optimizer(mm)
        opt.rng = InjectedNormal([torch.from_numpy(e[:, :, None].copy()) for e in z[f"f{f}__eps"]])
        for r in range(m["rows"]):
            vp.update_attributes({"target_position": float(z[f"f{f}__tp"][r]), "target_equilibrium": float(z[f"f{f}__te"][r]),
                                  "L": float(z[f"f{f}__L"][r])})
            u = float(opt.step(z[f"f{f}__s"][r].copy()))
            assert abs(u - float(Q[r, f])) < 2e-6, (f, r, u, Q[r, f])


def test_relabel_philox_streams_do_not_depend_on_the_sharding():
    from cartpolesimulation_b200.relabel import Relabeller
    rng = np.random.default_rng(0)
    R, E, K, T = 6, 4, 512, 30
    ang = rng.uniform(-np.pi, np.pi, (R, E))
    s = np.stack([ang, rng.uniform(-3, 3, (R, E)), np.cos(ang), np.sin(ang), rng.uniform(-0.1, 0.1, (R, E)),
                  rng.uniform(-0.3, 0.3, (R, E))], axis=2).astype(np.float32)
    Lr = rng.uniform(0.25, 0.55, (R, E)).astype(np.float32)
    whole = Relabeller(E, K, T, noise="philox", seed=11, device=0).relabel(s, pole_length=Lr)
    a = Relabeller(2, K, T, noise="philox", seed=11, file_offset=0, device=0).relabel(s[:, :2], pole_length=Lr[:, :2])
    b = Relabeller(2, K, T, noise="philox", seed=11, file_offset=2, device=0).relabel(s[:, 2:], pole_length=Lr[:, 2:])
    np.testing.assert_array_equal(whole, np.concatenate([a, b], axis=1))
    assert np.isfinite(whole).all() and (np.abs(whole) <= 1.0).all()
    # the pole length reaches the controller's model: a different L gives different controls
    other = Relabeller(E, K, T, noise="philox", seed=11, device=0).relabel(s, pole_length=np.full((R, E), 0.395, np.float32))
    assert np.abs(other - whole).max() > 1e-4


def test_add_control_along_trajectories_dataframes():
    import pandas as pd
    from cartpolesimulation_b200.relabel import Relabeller, add_control_along_trajectories
    z, m = load_golden("relabel_plain_ode")
    E = m["files"]
    dfs = [pd.DataFrame(z[f"f{f}__table"], columns=m["columns"]) for f in range(E)]
    cfg = dict(state_components=STATE, environment_attributes_dict=m["environment_attributes_dict"])
    rl = Relabeller(E, m["K"], m["T"], integrator="ODE", cost=m["cost"], noise="supplied", device=0)
    eps = np.ascontiguousarray(_stack(z, m, "eps").transpose(0, 1, 3, 2))
    out = add_control_along_trajectories(dfs, cfg, controller_output_variable_name="Q_calculated_offline",
                                         relabeller=rl, noise=torch.from_numpy(eps).cuda())
    assert len(out) == E
    for f in range(E):
        assert list(out[f].columns) == m["columns"] + ["Q_calculated_offline"]
        np.testing.assert_allclose(out[f]["Q_calculated_offline"].to_numpy(), z[f"f{f}__Q_calculated_offline"], rtol=0, atol=1e-4)
    # files of different length, integration over L with in-kernel noise, labels only
    zi, mi = load_golden("relabel_integrate_ode")
    dfi = [pd.DataFrame(zi[f"f{f}__table"], columns=mi["columns"]) for f in range(2)]
    dfi[1] = dfi[1].iloc[:3].copy()
    cfg = dict(state_components=STATE, environment_attributes_dict=mi["environment_attributes_dict"],
               mppi=dict(num_rollouts=256, horizon=20, cost="quadratic_boundary_grad", seed=3, device=0))
    lab = add_control_along_trajectories(dfi, cfg, controller_output_variable_name="Q_calculated_offline",
                                         integration_num_evals=8, save_output_only=True, seed=5)
    assert [len(x) for x in lab] == [5, 3] and list(lab[0].columns) == ["Q_calculated_offline"]
    assert all(np.isfinite(x.to_numpy()).all() and (np.abs(x.to_numpy()) <= 1.0).all() for x in lab)
    lab2 = add_control_along_trajectories(dfi, cfg, controller_output_variable_name="Q_calculated_offline",
                                          integration_num_evals=8, save_output_only=True, seed=5)
    np.testing.assert_array_equal(lab[0].to_numpy(), lab2[0].to_numpy())   # seeded: reproducible


def test_relabel_errors():
    from cartpolesimulation_b200.relabel import Relabeller
    rl = Relabeller(2, 64, 10, noise="supplied", device=0)
    with pytest.raises(ValueError):
        rl.relabel(np.zeros((3, 2, 6), np.float32))            # a 'supplied' fleet needs its noise
    with pytest.raises(ValueError):
        rl.relabel_device(torch.zeros((3, 5, 6), device="cuda"))   # wrong number of files


def test_nquad_relabelling_matches_reference():
    """integration_method='nquad' (the reference's default, adaptive quadrature): scipy's nquad drives every file, the
    integrand evaluations of all files run as masked fleet launches.  Against a recording of the UNMODIFIED reference
    (integration(), preprocess_data_add_control_along_trajectories.py:346-362, around optimizer_mppi with injected draws):
    the same number of controller steps per file -- i.e. the same adaptive subdivisions -- and the same labels."""
    import pandas as pd
    from cartpolesimulation_b200.relabel import Relabeller, add_control_along_trajectories, nquad_rows
    z, m = load_golden("relabel_nquad_ode")
    E = m["files"]
    dfs = [pd.DataFrame(z[f"f{f}__table"], columns=m["columns"]) for f in range(E)]
    for f in range(E):
        dfs[f]["Q_applied_-1"] = 0.0
    noise = [torch.from_numpy(np.ascontiguousarray(z[f"f{f}__eps"].transpose(0, 2, 1))).cuda() for f in range(E)]   # [calls, n_ind, K]
    steps = []

    class Counting(Relabeller):   # records which files stepped in which launch and what they answered
        def relabel_device(self, *a, **k):
            q = super().relabel_device(*a, **k)
            steps.append((k["active"].cpu().numpy()[0].copy(), q.cpu().numpy()[0].copy()))
            return q

    rl = Counting(E, m["K"], m["T"], integrator=m["predictor"], cost=m["cost"], noise="supplied", device=0)
    cfg = dict(state_components=STATE, environment_attributes_dict=m["environment_attributes_dict"])
    out = add_control_along_trajectories(dfs, cfg, "Q_calculated_offline", integration_method="nquad",
                                         integration_num_evals=m["evals"], relabeller=rl, noise=noise)
    for f in range(E):
        calls = sum(int(a[f]) for a, _ in steps)
        assert calls == int(z[f"f{f}__n_calls"]), (f, calls)            # the same adaptive path as the reference took
        u = np.array([q[f] for a, q in steps if a[f]])
        np.testing.assert_allclose(u, z[f"f{f}__u"], rtol=0, atol=1e-4)  # every controller step along the way
        np.testing.assert_allclose(out[f]["Q_calculated_offline"].to_numpy(), z[f"f{f}__Q_calculated_offline"], rtol=0,
                                   atol=1e-4)
    assert any(a.sum() < E for a, _ in steps)   # the files did take different numbers of steps: some launches were masked
    rl.close()
