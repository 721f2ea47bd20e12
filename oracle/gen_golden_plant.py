"""Generate tests/golden/plant_*.npz and closed_loop_*.npz by running the UNMODIFIED reference plant
(CartPole.update_state, CartPole/__init__.py:283-324) in this container.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (read-only; imported through oracle/ref_loader.py).

plant_*.npz        the plant alone: a stub controller returns a prescribed Q sequence; the reference CartPole is
                   stepped at dt_simulation = 2 ms with a controller period of 20 ms, exactly as
                   run_cartpole_random_experiment does (CartPole/__init__.py:659-739); the state after EVERY plant
                   tick, the second derivatives, the controller's inputs (s, time, target_position,
                   target_equilibrium) and the target-equilibrium flips are stored.
closed_loop_*.npz  plant + the reference optimizer_mppi (torch library, injected noise) in closed loop for a few
                   control periods: per period the state seen by the controller, the noise draws, the control.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")

from oracle import ref_loader as R  # noqa: E402


class StubController:
    """Stands in for controller_mpc: returns pre-supplied controls and records what the plant hands it
    (CartPole/__init__.py:506-517: s, time, updated_attributes)."""

    def __init__(self, Qs):
        self.Qs, self.i, self.calls = list(Qs), 0, []
        self.has_optimizer = False

    def step(self, s, time=None, updated_attributes=None):
        ua = updated_attributes or {}
        self.calls.append((np.array(s, dtype=np.float32).copy(), float(time), float(ua["target_position"]),
                           float(ua["target_equilibrium"])))
        q = self.Qs[self.i]
        self.i += 1
        return q

    def controller_reset(self):
        pass


def make_cartpole(controller, s0, target_position_f, target_equilibrium, keep_up, keep_down, length):
    from CartPole import CartPole
    with contextlib.redirect_stdout(io.StringIO()):
        cp = CartPole()
    cp.dt_simulation, cp.dt_controller, cp.dt_save = 0.002, 0.02, 0.02
    cp.controller_name, cp.controller = "mpc", controller
    cp.use_pregenerated_target_position = 1
    cp.random_track_f = target_position_f
    cp.length_of_experiment = length
    cp.target_equilibrium = target_equilibrium
    cp.keep_target_equilibrium_x_seconds_up = keep_up
    cp.keep_target_equilibrium_x_seconds_down = keep_down
    cp.save_data_in_cart = True
    cp.s = np.array(s0, dtype=np.float32)
    cp.target_position = target_position_f(0.0)
    cp.set_cartpole_state_at_t0(reset_mode=2, s=cp.s, target_position=cp.target_position)
    cp.cartpole_ode()  # run_cartpole_random_experiment, CartPole/__init__.py:679
    return cp


def run_plant(s0, Qs, n_ticks, tp_f, te0=1.0, keep_up=np.inf, keep_down=np.inf):
    ctrl = StubController(Qs)
    cp = make_cartpole(ctrl, s0, tp_f, te0, keep_up, keep_down, length=n_ticks * 0.002 + 1.0)
    states = [cp.s.copy()]
    dd = [(float(cp.angleDD), float(cp.positionDD))]
    qs = [float(cp.Q)]
    te = [float(cp.target_equilibrium)]
    tp = [float(cp.target_position)]
    for _ in range(n_ticks):
        cp.update_state()
        states.append(cp.s.copy())
        dd.append((float(cp.angleDD), float(cp.positionDD)))
        qs.append(float(cp.Q))
        te.append(float(cp.target_equilibrium))
        tp.append(float(cp.target_position))
    return dict(states=np.array(states, np.float32), dd=np.array(dd, np.float64), Q_tick=np.array(qs, np.float64),
                te_tick=np.array(te, np.float32), tp_tick=np.array(tp, np.float64),
                ctrl_s=np.array([c[0] for c in ctrl.calls], np.float32),
                ctrl_time=np.array([c[1] for c in ctrl.calls], np.float64),
                ctrl_tp=np.array([c[2] for c in ctrl.calls], np.float64),
                ctrl_te=np.array([c[3] for c in ctrl.calls], np.float32))


def gen_plant():
    rng = np.random.default_rng(7)
    cases = {
        # name: (s0, n control periods, Q generator, te0, keep_up, keep_down)
        "hanging": ([np.pi - 1e-3, 0.0, 0, 0, 0.0, 0.0], 60, lambda n: rng.uniform(-1, 1, n), 1.0, np.inf, np.inf),
        "upright": ([0.05, -0.2, 0, 0, 0.02, 0.1], 60, lambda n: rng.uniform(-0.4, 0.4, n), 1.0, np.inf, np.inf),
        # full power to one side: the cart reaches the track end -> edge_bounce (CartPole/__init__.py:460-470)
        "bounce": ([2.0, 1.0, 0, 0, 0.15, 0.5], 40, lambda n: np.full(n, 1.0), 1.0, np.inf, np.inf),
        # crossing +-pi repeatedly (fast spinning pole) -> wrap_angle_rad (fmod form)
        "spin": ([3.0, 12.0, 0, 0, 0.0, 0.0], 40, lambda n: rng.uniform(-0.2, 0.2, n), 1.0, np.inf, np.inf),
        # target equilibrium flips: 0.1 s up / 0.05 s down (update_target_equilibrium, :380-388)
        "flip": ([np.pi - 1e-3, 0.0, 0, 0, 0.0, 0.0], 30, lambda n: rng.uniform(-1, 1, n), 1.0, 0.1, 0.05),
    }
    for name, (s0, n_ctrl, qgen, te0, ku, kd) in cases.items():
        Qs = np.asarray(qgen(n_ctrl + 1), dtype=np.float64)
        Qs = np.float32(Qs).astype(np.float64)  # the controller returns float32 controls
        tp_f = (lambda t: 0.05 * np.sin(3.0 * t)) if name in ("upright", "flip") else (lambda t: 0.0)
        out = run_plant(s0, Qs, n_ctrl * 10, tp_f, te0, ku, kd)
        meta = dict(entry="CartPole.update_state (CartPole/__init__.py:283-324) at dt_simulation=0.002, dt_controller=0.02, "
                          "stub controller with prescribed Q", case=name)
        np.savez_compressed(os.path.join(GOLDEN, f"plant_{name}.npz"), s0=np.array(s0, np.float32), Q=Qs,
                            meta=json.dumps(meta), **out)
        print(f"plant_{name}: {out['states'].shape[0]} ticks, {len(out['ctrl_time'])} controller calls, "
              f"max|x| = {np.abs(out['states'][:, 4]).max():.4f}, te flips = {int((np.diff(out['te_tick']) != 0).sum())}")


class MppiController:
    """controller_mpc reduced to what the closed loop needs: reference optimizer_mppi + reference predictor_ODE +
    reference cost plugin under the torch library, noise injected (Control_Toolkit/Controllers/controller_mpc.py:95-112)."""

    def __init__(self, K, T, cost_name, draws):
        self.lib = R.torch_lib()
        self.vp = R.variable_parameters(self.lib)
        pred = R.ODECoreAdapter(T, K, variable_parameters=self.vp)
        cost = R.cost_function(cost_name, self.lib, self.vp, K, T)
        with contextlib.redirect_stdout(io.StringIO()):
            self.opt = R.optimizer_mppi(pred, cost, K, T, logging=False)
        self.opt.rng = R.InjectedNormal(draws)
        self.calls, self.us, self.u_noms = [], [], []
        self.has_optimizer = True

    def step(self, s, time=None, updated_attributes=None):
        ua = updated_attributes or {}
        self.vp.set_attributes({k: float(v) for k, v in ua.items() if k in ("target_position", "target_equilibrium", "L", "m_pole")})
        self.calls.append((np.array(s, dtype=np.float32).copy(), float(time), float(ua["target_position"]),
                           float(ua["target_equilibrium"])))
        u = self.opt.step(np.array(s, dtype=np.float32), time)
        self.us.append(float(u))
        self.u_noms.append(np.array(self.opt.u_nom).reshape(-1).astype(np.float32).copy())
        return u

    def controller_reset(self):
        pass


def gen_closed_loop():
    import torch
    K, T, n_ctrl = 256, 30, 25
    n_ind = int(np.ceil((T - 1) / 10)) + 1
    for name, cost_name, s0, keep in (("gradmin", "quadratic_boundary_grad_minimal", [np.pi - 1e-3, 0.0, 0, 0, 0.0, 0.0], (np.inf, np.inf)),
                                      ("gradmin_flip", "quadratic_boundary_grad_minimal", [0.3, 0.0, 0, 0, 0.05, 0.0], (0.2, 0.1))):
        g = torch.Generator().manual_seed(11)
        draws = [torch.randn((K, n_ind, 1), generator=g, dtype=torch.float32) for _ in range(n_ctrl + 1)]
        ctrl = MppiController(K, T, cost_name, draws)
        tp_f = lambda t: 0.04 * np.sin(2.0 * t)  # noqa: E731
        cp = make_cartpole(ctrl, s0, tp_f, 1.0, keep[0], keep[1], length=n_ctrl * 0.02 + 1.0)
        states = [cp.s.copy()]
        for _ in range(n_ctrl * 10):
            cp.update_state()
            states.append(cp.s.copy())
        meta = dict(entry="CartPole.update_state + optimizer_mppi.step (torch library, predictor_ODE, injected noise)",
                    K=K, T=T, cost=cost_name)
        np.savez_compressed(
            os.path.join(GOLDEN, f"closed_loop_{name}.npz"), s0=np.array(s0, np.float32),
            eps=np.stack([d.numpy()[:, :, 0] for d in draws]).astype(np.float32),
            states=np.array(states, np.float32), ctrl_s=np.array([c[0] for c in ctrl.calls], np.float32),
            ctrl_time=np.array([c[1] for c in ctrl.calls]), ctrl_tp=np.array([c[2] for c in ctrl.calls]),
            ctrl_te=np.array([c[3] for c in ctrl.calls], np.float32),
            Q=np.array(ctrl.us, np.float32), u_nom=np.array(ctrl.u_noms, np.float32),
            meta=json.dumps(meta))
        print(f"closed_loop_{name}: {len(ctrl.calls)} solves, final state {states[-1]}")


class SeqRng:
    """Stands in for a numpy Generator: hands out prepared float32 standard normals in call order."""

    def __init__(self, values):
        self.v, self.i = np.asarray(values, dtype=np.float32), 0

    def standard_normal(self, size=None, dtype=np.float32):
        x = self.v[self.i]
        self.i += 1
        return np.float32(x) if size is None or size == () else np.full(size, x, dtype=np.float32)


def gen_closed_loop_noisy():
    """Closed loop with the plant-side models ON: additive control disturbance, measurement noise, 5 ms latency (2.5 plant
    ticks) -- CartPole/noise_control_signal.py, noise_adder.py, latency_adder.py driven by the unmodified CartPole."""
    import torch
    import CartPole as CPmod
    import CartPole.noise_adder as NA
    K, T, n_ctrl = 256, 30, 20
    n_ind = int(np.ceil((T - 1) / 10)) + 1
    sig = dict(sigma_angle=0.01, sigma_position=0.002, sigma_angleD=0.075, sigma_positionD=0.02)
    cases = (("noisy", 0.005, "additive", 0.1, 0.02, True), ("latency", 0.013, "OFF", 0.0, 0.0, False))
    for name, latency, cmode, cmult, cadd, meas in cases:
        g = torch.Generator().manual_seed(21)
        draws = [torch.randn((K, n_ind, 1), generator=g, dtype=torch.float32) for _ in range(n_ctrl + 1)]
        rng = np.random.default_rng(5)
        meas_draws = rng.standard_normal((n_ctrl * 10, 4)).astype(np.float32)
        ctrl_draws = rng.standard_normal(n_ctrl + 8).astype(np.float32)
        ctrl = MppiController(K, T, "quadratic_boundary_grad_minimal", draws)
        tp_f = lambda t: 0.04 * np.sin(2.0 * t)  # noqa: E731
        saved = (CPmod.rng, CPmod.controlDisturbance_mode, float(CPmod.controlDisturbance), float(CPmod.controlBias),
                 NA.sigma_angle, NA.sigma_position, NA.sigma_angleD, NA.sigma_positionD)
        try:
            CPmod.rng = SeqRng(ctrl_draws)
            CPmod.controlDisturbance_mode = cmode
            CPmod.controlDisturbance[...] = cmult
            CPmod.controlBias[...] = cadd
            NA.sigma_angle, NA.sigma_position = sig["sigma_angle"], sig["sigma_position"]
            NA.sigma_angleD, NA.sigma_positionD = sig["sigma_angleD"], sig["sigma_positionD"]
            cp = make_cartpole(ctrl, [np.pi - 1e-3, 0.0, 0, 0, 0.0, 0.0], tp_f, 1.0, np.inf, np.inf, length=n_ctrl * 0.02 + 1.0)
            # make_cartpole made the controller call of t = 0 (true state); its control-noise draw is the last one consumed
            i0 = CPmod.rng.i - 1
            cp.NoiseAdderInstance.noise_mode = "ON" if meas else "OFF"
            cp.NoiseAdderInstance.rng_noise_adder = SeqRng(meas_draws.reshape(-1))
            cp.LatencyAdderInstance.dt_sampling = 0.002
            cp.LatencyAdderInstance.set_latency(latency)
            states, q_applied = [cp.s.copy()], [float(cp.Q_applied)]
            for _ in range(n_ctrl * 10):
                cp.update_state()
                states.append(cp.s.copy())
                q_applied.append(float(cp.Q_applied))
        finally:
            (CPmod.rng, CPmod.controlDisturbance_mode) = saved[0], saved[1]
            CPmod.controlDisturbance[...] = saved[2]
            CPmod.controlBias[...] = saved[3]
            NA.sigma_angle, NA.sigma_position, NA.sigma_angleD, NA.sigma_positionD = saved[4:]
        meta = dict(entry="CartPole.update_state + optimizer_mppi.step with add_control_noise / NoiseAdder / LatencyAdder "
                          "(CartPole/__init__.py:336-340,523-524)", K=K, T=T, cost="quadratic_boundary_grad_minimal",
                    latency=latency, control_noise_mode=cmode, control_noise=cmult, control_bias=cadd,
                    measurement_noise=bool(meas), **sig)
        np.savez_compressed(
            os.path.join(GOLDEN, f"closed_loop_{name}.npz"), eps=np.stack([d.numpy()[:, :, 0] for d in draws]).astype(np.float32),
            states=np.array(states, np.float32), ctrl_s=np.array([c[0] for c in ctrl.calls], np.float32),
            ctrl_time=np.array([c[1] for c in ctrl.calls]), ctrl_tp=np.array([c[2] for c in ctrl.calls]),
            ctrl_te=np.array([c[3] for c in ctrl.calls], np.float32), Q=np.array(ctrl.us, np.float32),
            Q_applied_tick=np.array(q_applied, np.float32), u_nom=np.array(ctrl.u_noms, np.float32),
            meas_draws=meas_draws, ctrl_draws=ctrl_draws[i0:i0 + n_ctrl + 1], meta=json.dumps(meta))
        print(f"closed_loop_{name}: {len(ctrl.calls)} solves, max |observed - true| = "
              f"{np.abs(np.array([c[0] for c in ctrl.calls])[1:] - np.array(states)[10::10][:len(ctrl.calls) - 1]).max():.4f}")


def gen_datagen():
    """Host-side pieces of the data generator: generate_random_initial_state (CartPole/data_generator.py:238-275) and
    Generate_Random_Trace_Function (CartPole/random_target_generator.py:9-87) with seeded numpy Generators."""
    from CartPole.data_generator import generate_random_initial_state
    from CartPole.random_target_generator import Generate_Random_Trace_Function
    from CartPole.state_utilities import create_cartpole_state
    out = {}
    times = np.concatenate([[0.0], np.cumsum(np.full(1800, 0.002))])[::10]
    for seed in range(6):
        rng = np.random.default_rng([5, seed])
        stub = create_cartpole_state()
        stub[:] = np.nan
        s0 = generate_random_initial_state(stub, init_limits=[0.8, 0.5, [0.0, 180.0], 1200.0], rng=rng)
        itype = ("previous", "0-derivative-smooth", "linear")[seed % 3]
        end_at = 1.0 * 0.198 * rng.uniform(-1.0, 1.0)
        f = Generate_Random_Trace_Function(length_of_experiment=3.6 if seed < 4 else 0.9, rtf_rng=rng,
                                           track_relative_complexity=2.0, interpolation_type=itype, turning_points=None,
                                           turning_points_period="regular" if seed % 2 == 0 else "random",
                                           start_random_target_position_at=float(s0[4]),
                                           end_random_target_position_at=end_at, used_track_fraction=1.0)
        out[f"s0_{seed}"] = np.asarray(s0, np.float32)
        out[f"tp_{seed}"] = np.asarray(f(np.minimum(times, 3.6 if seed < 4 else 0.9)), np.float64)
    meta = dict(entry="generate_random_initial_state + Generate_Random_Trace_Function with np.random.default_rng([5, seed])",
                n=6)
    np.savez_compressed(os.path.join(GOLDEN, "datagen_host.npz"), times=times, meta=json.dumps(meta), **out)
    print("datagen_host: 6 seeds")


if __name__ == "__main__":
    R.load()
    os.makedirs(GOLDEN, exist_ok=True)
    if "noisy" in sys.argv[1:]:   # only the plant-side-model fixtures
        gen_closed_loop_noisy()
        raise SystemExit(0)
    gen_plant()
    gen_datagen()
    if "--no-closed-loop" not in sys.argv:
        gen_closed_loop()
