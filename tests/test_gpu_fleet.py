"""Fleet (closed-loop experiments on the device, cps_fleet_*) against the oracle closed loop (MPPI solve + plant,
oracle/cps_oracle.c) and against recordings of the reference's own closed loop (tests/golden/closed_loop_*.npz)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.parity import load_golden

pytestmark = pytest.mark.gpu

RANGES = np.array([np.pi, 18.38, 1.0, 1.0, 0.198, 1.125], dtype=np.float64)  # the reference's normalisation ranges


def state_err(a, b):
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    d[..., 0] = np.minimum(d[..., 0], 2 * np.pi - d[..., 0])
    return float((d / RANGES).max())


def oracle_closed_loop(integ, cost, s0, eps, tp, te, T, mp=None):
    """eps [P, K, n_ind]; returns states [P+1, 6], Q [P], u_nom [P, T]."""
    s = np.array(s0, np.float32)
    u_nom, u_prev = np.zeros(T, np.float32), 0.0
    states, Qs, unoms = [s.copy()], [], []
    for j in range(eps.shape[0]):
        r = O.mppi_step(integ, cost, s, u_nom, eps=eps[j], u_prev=u_prev, target_position=float(tp[j]),
                        target_equilibrium=float(te[j]))
        u_nom, u_prev = r["u_nom"], float(r["u"])
        s, _, _ = O.plant_period(s, u_prev)
        states.append(s.copy()); Qs.append(u_prev); unoms.append(u_nom.copy())
    return np.array(states), np.array(Qs, np.float32), np.array(unoms)


@pytest.mark.parametrize("case", ["gradmin", "gradmin_flip"])
def test_fleet_reproduces_reference_closed_loop(case):
    """The reference's optimizer_mppi + CartPole.update_state in closed loop, 25 controller periods, injected noise."""
    from cartpolesimulation_b200.fleet import Fleet
    z, meta = load_golden(f"closed_loop_{case}")
    K, T = meta["K"], meta["T"]
    P = len(z["Q"]) - 1
    fl = Fleet(1, K, T, integrator="ODE", cost=meta["cost"], noise="supplied")
    s0 = z["ctrl_s"][0:1].copy()
    fl.reset(s0)
    eps = torch.from_numpy(np.ascontiguousarray(z["eps"][:P].transpose(0, 2, 1))[:, None]).cuda()  # [P,1,n_ind,K]
    tp = torch.from_numpy(z["ctrl_tp"][:P].astype(np.float32)[:, None].copy()).cuda()
    te = torch.from_numpy(z["ctrl_te"][:P].astype(np.float32)[:, None].copy()).cuda()
    rec = torch.zeros((P, 1, 16), device="cuda")
    fl.run(P, tp, te, eps.contiguous(), rec)
    torch.cuda.synchronize()
    rec = rec.cpu().numpy()[:, 0]
    # record row j: the state the controller saw in period j and the control it chose
    got_s = rec[:, [1, 2, 4, 5, 6, 7]]
    assert state_err(got_s, z["ctrl_s"][:P]) < 1e-5
    np.testing.assert_allclose(rec[:, 9], z["Q"][:P], rtol=0, atol=1e-4)     # north_star: selected control within 1e-4
    np.testing.assert_allclose(rec[:, 12], z["ctrl_tp"][:P], atol=1e-7)
    np.testing.assert_array_equal(rec[:, 13], z["ctrl_te"][:P])
    s_end, u_nom, u_prev = fl.states(with_u_nom=True)
    assert state_err(s_end[0], z["ctrl_s"][P]) < 1e-5
    np.testing.assert_allclose(u_nom[0], z["u_nom"][P - 1], atol=1e-4)
    assert fl.period == P
    fl.close()


@pytest.mark.parametrize("integ", ["ODE", "ODE_v0"])
@pytest.mark.parametrize("cost", ["quadratic_boundary_grad_minimal", "quadratic_boundary_grad"])
def test_fleet_vs_oracle(integ, cost):
    """E = 5 experiments with different states / targets / equilibria, K not a multiple of the block size."""
    from cartpolesimulation_b200.fleet import Fleet
    E, K, T, P = 5, 300, 20, 12
    rng = np.random.default_rng(3)
    n_ind = O.num_inducing(T, 10)
    ang = np.array([np.pi - 1e-3, 0.1, -2.0, 1.0, 3.0])
    s0 = np.stack([ang, rng.uniform(-2, 2, E), np.cos(ang), np.sin(ang), rng.uniform(-0.1, 0.1, E), rng.uniform(-0.2, 0.2, E)], 1).astype(np.float32)
    eps = rng.standard_normal((P, E, K, n_ind)).astype(np.float32)
    tp = rng.uniform(-0.1, 0.1, (P, E)).astype(np.float32)
    te = np.where(rng.uniform(size=(P, E)) < 0.3, -1.0, 1.0).astype(np.float32)
    fl = Fleet(E, K, T, integrator=integ, cost=cost, noise="supplied")
    fl.reset(s0)
    rec = torch.zeros((P, E, 16), device="cuda")
    J = torch.zeros((P, E, K), device="cuda")
    fl.run(P, torch.from_numpy(tp).cuda(), torch.from_numpy(te).cuda(),
           torch.from_numpy(np.ascontiguousarray(eps.transpose(0, 1, 3, 2))).cuda(), rec, J)
    torch.cuda.synchronize()
    rec, J = rec.cpu().numpy(), J.cpu().numpy()
    s_end, u_nom, u_prev = fl.states(with_u_nom=True)
    for e in range(E):
        st, Qs, unoms = oracle_closed_loop(integ, cost, s0[e], eps[:, e], tp[:, e], te[:, e], T)
        assert state_err(rec[:, e][:, [1, 2, 4, 5, 6, 7]], st[:P]) < 1e-5, (integ, cost, e)
        np.testing.assert_allclose(rec[:, e, 9], Qs, rtol=0, atol=1e-4)
        assert state_err(s_end[e], st[P]) < 1e-5
        np.testing.assert_allclose(u_nom[e], unoms[-1], atol=1e-4)
        assert abs(u_prev[e] - Qs[-1]) < 1e-4
    # first-period costs against the oracle (no closed-loop amplification yet)
    r = O.mppi_step(integ, cost, s0[2], np.zeros(T, np.float32), eps=eps[0, 2], u_prev=0.0, target_position=float(tp[0, 2]),
                    target_equilibrium=float(te[0, 2]), want=("J",))
    np.testing.assert_allclose(J[0, 2], r["J"], rtol=3e-5, atol=3e-5 * np.abs(r["J"]).max())
    fl.close()


def test_fleet_plant_only():
    """With zero-width control limits the controller returns exactly the clip value, which leaves the plant alone
    under test.  It restates the oracle's plant (pinned bit-exact to the reference's) operation by operation; the only
    difference is CUDA's cosf/sinf vs libm's (<= 1 ulp), so one period restarted from the same state agrees to a few
    ulp and 40 free-running periods (bounces, wraps) to 1e-5 of the state ranges."""
    from cartpolesimulation_b200.fleet import Fleet
    z, _ = load_golden("plant_bounce")
    E, K, T, P = 3, 64, 10, 40
    fl = Fleet(E, K, T, integrator="ODE", cost="quadratic_boundary_grad_minimal", noise="philox", seed=1)
    s0 = np.stack([z["states"][0], z["states"][0], z["states"][0]]).copy()
    s0[1, 0], s0[1, 1] = 3.0, 12.0          # spinning pole: wraps
    s0[1, 2], s0[1, 3] = np.cos(s0[1, 0]), np.sin(s0[1, 0])
    s0[2, 4], s0[2, 5] = -0.15, -0.5
    for q in (1.0, -1.0):
        fl.engine.set_mppi_params(lo=q, hi=q)
        fl.reset(s0)
        rec = torch.zeros((P, E, 16), device="cuda")
        fl.run(P, record=rec)
        torch.cuda.synchronize()
        rec = rec.cpu().numpy()
        exact = total = 0
        for e in range(E):
            s_free = s0[e].copy()
            for j in range(P):
                got = rec[j, e][[1, 2, 4, 5, 6, 7]]
                assert state_err(got, s_free) < 1e-5, (q, e, j, got, s_free)
                assert rec[j, e, 9] == q and rec[j, e, 11] == np.float32(1.77) * np.float32(q)
                _, _, dd = O.plant_period(got, q)
                np.testing.assert_allclose(rec[j, e, [3, 8]], dd[0].astype(np.float32), rtol=1e-6)
                if j + 1 < P:   # one period restarted from the kernel's own state
                    nxt, _, _ = O.plant_period(got, q)
                    ref = rec[j + 1, e][[1, 2, 4, 5, 6, 7]]
                    assert state_err(ref, nxt) < 3e-7, (q, e, j, ref, nxt)
                    exact += int((ref.view(np.uint32) == nxt.view(np.uint32)).sum()); total += 6
                s_free, _, _ = O.plant_period(s_free, q)
        assert exact > 0.7 * total, (exact, total)   # most values are bit-identical
        if q == 1.0:
            assert np.abs(rec[:, 0, 6]).max() > 0.19   # the cart really reached the track end
    fl.close()


def test_philox_noise_statistics_and_determinism():
    from cartpolesimulation_b200.fleet import Fleet
    E, K, T = 4, 2000, 50
    fl = Fleet(E, K, T, noise="philox", seed=1234)
    a = fl.noise(0).cpu().numpy()
    b = fl.noise(1).cpu().numpy()
    assert a.shape == (E, 6, K)
    allv = np.concatenate([a.ravel(), b.ravel()])
    assert abs(allv.mean()) < 0.01 and abs(allv.std() - 1.0) < 0.01 and np.isfinite(allv).all()
    assert abs((allv ** 3).mean()) < 0.03 and abs((allv ** 4).mean() - 3.0) < 0.1
    assert np.abs(allv).max() < 6.0
    # distinct streams per experiment, period, rollout and draw; reproducible
    assert len(np.unique(allv)) > 0.99 * allv.size
    np.testing.assert_array_equal(a, fl.noise(0).cpu().numpy())
    assert abs(np.corrcoef(a[0].ravel(), a[1].ravel())[0, 1]) < 0.03
    assert abs(np.corrcoef(a[0].ravel(), b[0].ravel())[0, 1]) < 0.03
    fl2 = Fleet(2, K, T, noise="philox", seed=1234, experiment_offset=2)
    np.testing.assert_array_equal(fl2.noise(1).cpu().numpy(), b[2:])
    fl3 = Fleet(E, K, T, noise="philox", seed=1235)
    assert not np.array_equal(fl3.noise(0).cpu().numpy(), a)
    for f in (fl, fl2, fl3):
        f.close()


def test_philox_fleet_equals_supplied_fleet_and_sharding():
    """In-kernel generation == the same draws passed in; splitting the fleet over two handles changes nothing."""
    from cartpolesimulation_b200.fleet import Fleet, make_experiments, DataGenConfig
    E, K, T, P = 6, 500, 30, 8
    cfg = DataGenConfig(length_of_experiment=2.0, keep_target_equilibrium_x_seconds_up=0.06,
                        keep_target_equilibrium_x_seconds_down=0.04)
    s0, tp, te = make_experiments(E, P, cfg, seed=2)
    a = Fleet(E, K, T, noise="philox", seed=9)
    a.reset(s0)
    rec_a = torch.zeros((P, E, 16), device="cuda")
    a.run(P, torch.from_numpy(tp).cuda(), torch.from_numpy(te).cuda(), None, rec_a)
    noise = torch.stack([a.noise(j) for j in range(P)])
    b = Fleet(E, K, T, noise="supplied")
    b.reset(s0)
    rec_b = torch.zeros((P, E, 16), device="cuda")
    b.run(P, torch.from_numpy(tp).cuda(), torch.from_numpy(te).cuda(), noise.contiguous(), rec_b)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(rec_a.cpu().numpy(), rec_b.cpu().numpy())
    np.testing.assert_array_equal(a.states(), b.states())
    # shard: experiments 4..5 on their own handle
    s0c, tpc, tec = make_experiments(2, P, cfg, seed=2, experiment_offset=4)
    c = Fleet(2, K, T, noise="philox", seed=9, experiment_offset=4)
    c.reset(s0c)
    rec_c = torch.zeros((P, 2, 16), device="cuda")
    c.run(P, torch.from_numpy(tpc).cuda(), torch.from_numpy(tec).cuda(), None, rec_c)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(rec_c.cpu().numpy(), rec_a.cpu().numpy()[:, 4:])
    assert (np.diff(te[:, 0]) != 0).sum() >= 2   # the equilibrium really flipped during the run
    for f in (a, b, c):
        f.close()


def test_fleet_single_experiment_equals_engine_solve():
    """One period of a one-experiment fleet is the same solve as cps_mppi_step (different block geometry only)."""
    from cartpolesimulation_b200.core import Engine
    from cartpolesimulation_b200.fleet import Fleet
    from cartpolesimulation_b200 import _lib as L
    K, T = 2000, 50
    fl = Fleet(1, K, T, integrator="ODE_v0", noise="philox", seed=5)
    a = np.pi - 1e-3
    s = np.array([[a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0]], dtype=np.float32)
    fl.reset(s)
    fl.run(1)
    _, u_nom, u_prev = fl.states(with_u_nom=True)
    eng = Engine(K, T, integrator="ODE_v0", cost="quadratic_boundary_grad_minimal")
    u = eng.mppi_step(torch.from_numpy(s[0]).cuda(), fl.noise(0)[0].contiguous(), L.TIME_MAJOR, 0.0)
    torch.cuda.synchronize()
    np.testing.assert_allclose(eng.get_u_nom(), u_nom[0], atol=2e-6)
    assert abs(float(u.cpu()[0]) - u_prev[0]) < 2e-6
    fl.close(); eng.close()


def test_fleet_errors():
    from cartpolesimulation_b200.fleet import Fleet
    fl = Fleet(2, 64, 10, noise="supplied")
    with pytest.raises(ValueError):
        fl.run(1)                       # supplied fleet without noise
    with pytest.raises(ValueError):
        fl.reset(np.zeros((3, 6), np.float32))
    fl.close()
    fl = Fleet(2, 64, 10, noise="philox")
    with pytest.raises(ValueError):
        fl.run(1, noise=torch.zeros((1, 2, 2, 64), device="cuda"))
    fl.close()
    with pytest.raises(NotImplementedError):
        Fleet(2, 64, 10, integrator="neural")


@pytest.mark.parametrize("integ", ["ODE", "ODE_v0"])
def test_packed_fleet_solve_matches_one_rollout_per_thread(integ):
    """From 65536 rollouts per launch on the fleet's solve runs two rollouts per thread in packed FP32: per rollout the
    same arithmetic (bit-identical costs), the control equal to summation order; closed loop and relabel mode."""
    from cartpolesimulation_b200.fleet import Fleet, make_experiments
    from cartpolesimulation_b200.relabel import Relabeller
    E, K, T, P = 40, 2000, 30, 3
    s0, tp, te = make_experiments(E, P, seed=3)
    out = []
    for no_pairs in (True, False):
        fl = Fleet(E, K, T, integrator=integ, noise="philox", seed=7, device=0, no_pairs=no_pairs)
        fl.reset(s0)
        rec = torch.zeros((P, E, 16), device="cuda")
        J = torch.zeros((P, E, K), device="cuda")
        fl.run(P, torch.from_numpy(tp).cuda(), torch.from_numpy(te).cuda(), record=rec, J_out=J)
        out.append((rec.cpu().numpy(), J.cpu().numpy(), fl.states()))
        fl.close()
    (r1, J1, s1), (r2, J2, s2) = out
    np.testing.assert_array_equal(J1[0], J2[0])                       # first period: identical inputs -> identical costs
    np.testing.assert_allclose(r2[..., 9], r1[..., 9], rtol=0, atol=5e-6)   # Q_calculated of every period
    np.testing.assert_allclose(s2, s1, rtol=0, atol=2e-4)             # closed loop over three periods
    rng = np.random.default_rng(1)
    ang = rng.uniform(-np.pi, np.pi, (4, E))
    st = np.stack([ang, rng.uniform(-3, 3, (4, E)), np.cos(ang), np.sin(ang), rng.uniform(-0.1, 0.1, (4, E)),
                   rng.uniform(-0.3, 0.3, (4, E))], axis=2).astype(np.float32)
    Lr = rng.uniform(0.25, 0.55, (4, E)).astype(np.float32)
    q = [Relabeller(E, K, T, integrator=integ, noise="philox", seed=5, device=0, no_pairs=n).relabel(st, pole_length=Lr)
         for n in (True, False)]
    np.testing.assert_allclose(q[1], q[0], rtol=0, atol=5e-6)


def test_fleet_full_size_is_independent_of_the_sharding():
    """BASELINE configs[4] size: 8192 experiments x (K = 2000, T = 50).  Size-independent property: the record rows of a
    period do not depend on how the experiments are split over devices (in-kernel Philox streams are keyed by the global
    experiment index) -- one fleet of 8192 against four fleets of 2048 with their offsets, bit for bit; every control
    inside the limits and finite."""
    from cartpolesimulation_b200.fleet import Fleet, make_experiments
    E, K, T, P = 8192, 2000, 50, 2
    s0, tp, te = make_experiments(E, P, seed=11)
    tp_d, te_d = torch.from_numpy(tp).cuda(), torch.from_numpy(te).cuda()
    whole = Fleet(E, K, T, noise="philox", seed=4, device=0)
    whole.reset(s0)
    rec = torch.zeros((P, E, 16), device="cuda")
    whole.run(P, tp_d, te_d, record=rec)
    rec = rec.cpu().numpy()
    whole.close()
    assert np.isfinite(rec).all() and (np.abs(rec[..., 9]) <= 1.0).all()
    for part in range(4):
        lo, hi = part * 2048, (part + 1) * 2048
        fl = Fleet(2048, K, T, noise="philox", seed=4, device=0, experiment_offset=lo)
        fl.reset(s0[lo:hi])
        r = torch.zeros((P, 2048, 16), device="cuda")
        fl.run(P, tp_d[:, lo:hi].contiguous(), te_d[:, lo:hi].contiguous(), record=r)
        np.testing.assert_array_equal(r.cpu().numpy(), rec[:, lo:hi])
        fl.close()


@pytest.mark.parametrize("case", ["noisy", "latency"])
def test_fleet_plant_side_models_reproduce_reference(case):
    """Control disturbance, measurement noise and latency (CartPole/noise_control_signal.py, noise_adder.py,
    latency_adder.py; all OFF in the shipped configuration) against the unmodified CartPole + optimizer_mppi in closed loop
    with the models ON and every draw injected: what the controller saw in every period (delayed, interpolated, noisy),
    the control it chose, the control the plant received, and the true state."""
    from cartpolesimulation_b200.fleet import Fleet
    z, m = load_golden(f"closed_loop_{case}")
    K, T = m["K"], m["T"]
    P = len(z["Q"]) - 1
    fl = Fleet(1, K, T, integrator="ODE", cost=m["cost"], noise="supplied")
    fl.set_plant_models(control_noise_mode=m["control_noise_mode"], control_noise=m["control_noise"], control_bias=m["control_bias"],
                        noise_mode="ON" if m["measurement_noise"] else "OFF", sigma_angle=m["sigma_angle"],
                        sigma_position=m["sigma_position"], sigma_angleD=m["sigma_angleD"], sigma_positionD=m["sigma_positionD"],
                        latency=m["latency"])
    fl.reset(z["states"][0:1].copy())
    np.testing.assert_array_equal(fl.observed()[0], z["states"][0])     # t = 0: the controller sees the true state
    eps = torch.from_numpy(np.ascontiguousarray(z["eps"][:P].transpose(0, 2, 1))[:, None]).cuda().contiguous()  # [P,1,n_ind,K]
    tp = torch.from_numpy(z["ctrl_tp"][:P].astype(np.float32)[:, None].copy()).cuda()
    te = torch.from_numpy(z["ctrl_te"][:P].astype(np.float32)[:, None].copy()).cuda()
    meas = torch.from_numpy(z["meas_draws"].reshape(P, 10, 1, 4).copy()).cuda()
    cd = torch.from_numpy(z["ctrl_draws"][:P].reshape(P, 1).copy()).cuda()
    rec = torch.zeros((P, 1, 16), device="cuda")
    obs = []
    for j in range(P):
        fl.run(1, tp[j:j + 1], te[j:j + 1], eps[j:j + 1], rec[j:j + 1], ctrl_draws=cd[j:j + 1], meas_draws=meas[j:j + 1])
        obs.append(fl.observed()[0])
    rec = rec.cpu().numpy()[:, 0]
    assert state_err(np.array(obs), z["ctrl_s"][1:P + 1]) < 2e-5               # the controller's view, period by period
    np.testing.assert_allclose(rec[:, 9], z["Q"][:P], rtol=0, atol=1e-4)       # Q_calculated (north_star: 1e-4)
    np.testing.assert_allclose(rec[:, 10], z["Q_applied_tick"][0:10 * P:10], rtol=0, atol=1e-4)   # what drove the plant
    assert state_err(rec[:, [1, 2, 4, 5, 6, 7]], z["states"][0:10 * P:10]) < 2e-5   # the TRUE state stays in the record
    assert state_err(fl.states()[0], z["states"][10 * P]) < 2e-5
    if case == "noisy":
        assert np.abs(np.array(obs) - z["states"][10:10 * P + 1:10]).max() > 1e-2   # the view really differs from the truth
    fl.close()


def test_fleet_plant_side_models_philox_and_errors():
    """A Philox fleet draws the plant-side noise itself, reproducibly and independently of the sharding; truncnorm keeps
    the applied control inside [-1, 1]; a supplied-noise fleet refuses to run without the plant-side draws."""
    from cartpolesimulation_b200.fleet import Fleet
    E, K, T, P = 4, 128, 20, 6
    ang = np.array([3.0, 0.2, -1.0, 2.0])
    s0 = np.stack([ang, np.zeros(E), np.cos(ang), np.sin(ang), np.zeros(E), np.zeros(E)], 1).astype(np.float32)
    kw = dict(control_noise_mode="truncnorm", control_noise=0.5, control_bias=0.0, noise_mode="ON", sigma_angle=0.01,
              sigma_position=0.002, sigma_angleD=0.05, sigma_positionD=0.02, latency=0.004)
    recs = []
    for off, sl in ((0, slice(0, 4)), (0, slice(0, 4)), (2, slice(2, 4))):
        fl = Fleet(sl.stop - sl.start, K, T, noise="philox", seed=3, experiment_offset=off)
        fl.set_plant_models(**kw)
        fl.reset(s0[sl])
        rec = torch.zeros((P, sl.stop - sl.start, 16), device="cuda")
        fl.run(P, record=rec)
        recs.append(rec.cpu().numpy())
        fl.close()
    np.testing.assert_array_equal(recs[0], recs[1])            # reproducible
    np.testing.assert_array_equal(recs[0][:, 2:4], recs[2])    # streams keyed by the global experiment index
    assert np.abs(recs[0][:, :, 10]).max() <= 1.0 and np.abs(recs[0][:, :, 10] - recs[0][:, :, 9]).max() > 0.05
    fl = Fleet(1, K, T, noise="supplied")
    fl.set_plant_models(**kw)
    fl.reset(s0[:1])
    with pytest.raises(Exception):
        fl.run(1, noise=torch.zeros((1, 1, fl.n_ind, K), device="cuda"))
    fl.close()
