"""Generate tests/golden/*.npz by running the UNMODIFIED reference code in this container.

TEST INFRASTRUCTURE ONLY.  Run from anywhere:  python oracle/gen_golden.py
Needs /root/reference (read-only) -- it is imported through oracle/ref_loader.py (stubs for packages
missing from the image; nothing is written into the reference tree).  The fixtures travel to the GPU
box; this script and the reference do not need to.

Each fixture stores inputs AND the reference outputs, plus `meta` naming the reference entry point.
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


def make_states(rng, B, kind="random"):
    if kind == "random":
        ang = rng.uniform(-np.pi, np.pi, B)
        angD = rng.uniform(-5, 5, B)
        pos = rng.uniform(-0.8 * 0.198, 0.8 * 0.198, B)
        posD = rng.uniform(-0.5, 0.5, B)
    elif kind == "edge":  # heading into the track end -> edge_bounce fires (ODE_v0 only models it)
        ang = rng.uniform(-np.pi, np.pi, B)
        angD = rng.uniform(-3, 3, B)
        sign = np.where(rng.uniform(size=B) < 0.5, -1.0, 1.0)
        pos = sign * rng.uniform(0.17, 0.1975, B)
        posD = sign * rng.uniform(0.2, 1.0, B)
    elif kind == "wrap":  # crossing +-pi
        sign = np.where(rng.uniform(size=B) < 0.5, -1.0, 1.0)
        ang = sign * rng.uniform(np.pi - 0.05, np.pi, B)
        angD = sign * rng.uniform(0.5, 8, B)
        pos = rng.uniform(-0.1, 0.1, B)
        posD = rng.uniform(-0.2, 0.2, B)
    elif kind == "upright":
        ang = rng.uniform(-0.3, 0.3, B)
        angD = rng.uniform(-1, 1, B)
        pos = rng.uniform(-0.1, 0.1, B)
        posD = rng.uniform(-0.2, 0.2, B)
    else:
        raise ValueError(kind)
    s = np.stack([ang, angD, np.cos(ang), np.sin(ang), pos, posD], 1)
    return s.astype(np.float32)


def hanging_state(eps=1e-3):
    a = np.pi - eps
    return np.array([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], dtype=np.float32)


def save(name, meta, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def gen_rollouts():
    import torch
    rng = np.random.default_rng(1234)
    T, n, dt = 50, 10, 0.02
    cases = {}
    # Appendix-D style known answer
    s_ka = np.array([[0.1, 0, np.cos(0.1), np.sin(0.1), 0, 0]], dtype=np.float32)
    cases["known"] = (s_ka, np.full((1, 3), 0.5, dtype=np.float32), {})
    for kind, B in (("random", 192), ("edge", 64), ("wrap", 64), ("upright", 64)):
        cases[kind] = (make_states(rng, B, kind), rng.uniform(-1, 1, (B, T)).astype(np.float32), {})
    # one state tiled over K control sequences, MPPI-like small perturbations (config C1)
    cases["tiled"] = (hanging_state()[None], np.clip(rng.normal(0, 0.2121, (128, T)), -1, 1).astype(np.float32), {})
    # variable parameters
    cases["varL"] = (make_states(rng, 32, "random"), rng.uniform(-1, 1, (32, T)).astype(np.float32),
                     {"L": 0.3, "m_pole": 0.12})
    # long horizon (config 4 shape) and a coarse substep count
    cases["T100"] = (make_states(rng, 32, "random"), rng.uniform(-1, 1, (32, 100)).astype(np.float32), {})
    cases["n1"] = (make_states(rng, 32, "random"), rng.uniform(-1, 1, (32, 20)).astype(np.float32), {"n": 1})

    out_v0, out_ode = {}, {}
    for name, (s0, Q, var) in cases.items():
        B, Tc = Q.shape
        nn = var.get("n", n)
        lib = R.torch_lib()
        vp = None
        if "L" in var:
            vp_np = type("VP", (), {})()
            vp_np.L = np.float32(var["L"])  # predictors_customization_v0.py:47-48 reads .L as is
            vp = R.variable_parameters(lib, L=var["L"], m_pole=var["m_pole"])
        else:
            vp_np = None
        p0 = R.predictor_ODE_v0(Tc, dt, nn, B, variable_parameters=vp_np)
        ref0 = np.asarray(p0.predict(s0.copy(), Q[:, :, None].copy())).reshape(B, Tc + 1, 6)
        p1 = R.predictor_ODE(Tc, dt, nn, B, variable_parameters=vp)
        s_in = torch.from_numpy(np.tile(s0, (B, 1)) if s0.shape[0] == 1 and B > 1 else s0.copy())
        ref1 = p1.predict_core(s_in, torch.from_numpy(Q[:, :, None].copy())).numpy()
        for out, ref in ((out_v0, ref0), (out_ode, ref1)):
            out[f"{name}__s0"] = s0
            out[f"{name}__Q"] = Q
            out[f"{name}__traj"] = ref.astype(np.float32)
            out[f"{name}__var"] = np.array([var.get("L", 0.395), var.get("m_pole", 0.087), nn], dtype=np.float64)
    meta = dict(dt=dt, cases=list(cases.keys()))
    save("rollout_ode_v0", dict(meta, ref="SI_Toolkit/Predictors/predictor_ODE_v0.py:42-74 predict (numba, explicit Euler)"),
         **out_v0)
    save("rollout_ode", dict(meta, ref="SI_Toolkit/Predictors/predictor_ODE.py:86-97 _predict_core (torch fp32, Euler-Cromer)"),
         **out_ode)


def _patch_torch_lib_for_grad(lib):
    """quadratic_boundary_grad is TF-only in the reference: PyTorchLibrary has no stop_gradient/cond
    (computation_library.py:501).  Forward-pass semantics of both are trivial; supply them."""
    lib.stop_gradient = lambda x: x
    lib.cond = lambda c, true_fn, false_fn: true_fn() if bool(c) else false_fn()
    return lib


def gen_costs():
    import torch
    from oracle import oracle as O
    rng = np.random.default_rng(77)
    T = 20
    arrays = {}
    names = ["default", "quadratic_boundary", "quadratic_boundary_grad_minimal", "quadratic_boundary_grad"]
    # trajectories: real rollouts (oracle only generates INPUTS here) + synthetic states that cross every
    # threshold (track boundary fractions 0.85/0.90/0.95, terminal |angle|>0.2, |x-x*|>0.1*THL)
    s0 = make_states(rng, 96, "random")
    Q = rng.uniform(-1, 1, (96, T)).astype(np.float32)
    traj = O.rollout("ODE_v0", s0, Q)
    synth = np.zeros((96, T + 1, 6), dtype=np.float32)
    ang = rng.uniform(-np.pi, np.pi, (96, T + 1))
    ang[:32, -1] = rng.uniform(-0.3, 0.3, 32)  # terminal indicator both sides
    pos = rng.uniform(-0.21, 0.21, (96, T + 1))
    pos[:32, -1] = rng.uniform(-0.03, 0.03, 32)
    synth[..., 0] = ang
    synth[..., 1] = rng.uniform(-12, 12, (96, T + 1))
    synth[..., 2] = np.cos(ang)
    synth[..., 3] = np.sin(ang)
    synth[..., 4] = pos
    synth[..., 5] = rng.uniform(-1, 1, (96, T + 1))
    traj = np.concatenate([traj, synth], 0).astype(np.float32)
    Q = np.concatenate([Q, rng.uniform(-1, 1, (96, T)).astype(np.float32)], 0)
    arrays["traj"], arrays["Q"] = traj, Q
    settings = [(0.0, 1.0, 0.0), (0.05, 1.0, 0.2), (-0.1, -1.0, -0.4)]
    arrays["settings"] = np.array(settings, dtype=np.float64)
    for name in names:
        for i, (tp, te, up) in enumerate(settings):
            lib = R.torch_lib()
            if name == "quadratic_boundary_grad":
                _patch_torch_lib_for_grad(lib)
            vp = R.variable_parameters(lib, tp, te)
            cw = R.cost_function(name, lib, vp, traj.shape[0], T)
            tt, qq = torch.from_numpy(traj), torch.from_numpy(Q[:, :, None])
            u_prev = np.float32(up)
            stage = cw.get_stage_cost(tt[:, :-1, :], qq, u_prev)
            J = cw.get_trajectory_cost(tt, qq, u_prev)
            arrays[f"{name}__{i}__stage"] = stage.numpy().astype(np.float32)
            arrays[f"{name}__{i}__J"] = J.numpy().astype(np.float32)
            arrays[f"{name}__{i}__terminal"] = cw.get_terminal_cost(tt[:, -1, :]).numpy().astype(np.float32).reshape(-1)
            if hasattr(cw.cost_function, "MAX_COST"):
                arrays[f"{name}__max_cost"] = np.array(float(cw.cost_function.MAX_COST), dtype=np.float64)
    save("costs", dict(ref="Control_Toolkit/Cost_Functions/__init__.py:49-93 + Control_Toolkit_ASF/Cost_Functions/CartPole/*.py "
                           "(torch lib; quadratic_boundary_grad with identity stop_gradient/cond shims)",
                       names=names, T=T), **arrays)


def gen_interp():
    import torch
    R.load()
    from Control_Toolkit.others.Interpolator import Interpolator
    rng = np.random.default_rng(5)
    arrays = {}
    combos = [(50, 10), (35, 10), (100, 10), (51, 10), (41, 10), (7, 10), (20, 1), (50, 7), (2, 10), (11, 5)]
    for (T, p) in combos:
        with contextlib.redirect_stdout(io.StringIO()):
            it = Interpolator(T, p, 1, R.torch_lib())
        n_ind = it.number_of_interpolation_inducing_points
        eps = rng.normal(size=(16, n_ind, 1)).astype(np.float32)
        stdev = np.float32(0.03 / np.sqrt(0.02))
        y = it.interpolate(torch.from_numpy(eps) * torch.as_tensor(stdev))
        arrays[f"T{T}_p{p}__W"] = it.interp_mat.numpy()[:, :, 0].astype(np.float32)  # [n_ind, T]
        arrays[f"T{T}_p{p}__eps"] = eps[:, :, 0]
        arrays[f"T{T}_p{p}__delta_u"] = y.numpy()[:, :, 0].astype(np.float32)
    save("interp", dict(ref="Control_Toolkit/others/Interpolator.py:53-106", combos=combos, stdev=float(stdev)),
         **arrays)


def gen_mppi():
    import torch
    from oracle import oracle as O
    runs = [
        # name, predictor, cost, K, T, steps, target_position, target_equilibrium
        ("ode_gradmin", "ODE", "quadratic_boundary_grad_minimal", 512, 50, 4, 0.0, 1.0),
        ("v0_gradmin", "ODE_v0", "quadratic_boundary_grad_minimal", 512, 50, 4, 0.0, 1.0),
        ("ode_gradmin_K2000", "ODE", "quadratic_boundary_grad_minimal", 2000, 50, 2, 0.0, 1.0),
        ("ode_grad", "ODE", "quadratic_boundary_grad", 256, 35, 3, 0.05, 1.0),
        ("ode_grad_down", "ODE", "quadratic_boundary_grad", 256, 35, 3, -0.05, -1.0),
        ("ode_qb", "ODE", "quadratic_boundary", 256, 50, 3, 0.0, 1.0),
        ("ode_default", "ODE", "default", 256, 50, 3, 0.0, 1.0),
        ("ode_gradmin_T100", "ODE", "quadratic_boundary_grad_minimal", 256, 100, 2, 0.1, 1.0),
        ("ode_gradmin_T51", "ODE", "quadratic_boundary_grad_minimal", 128, 51, 2, 0.0, 1.0),
        # BASELINE.json configs[0] exactly: predictor_ODE_v0 (explicit Euler, numba), K = 2000, T = 50
        ("v0_gradmin_K2000", "ODE_v0", "quadratic_boundary_grad_minimal", 2000, 50, 2, 0.0, 1.0),
        # BASELINE.json configs[3] at full size: K = 65536, T = 100, barrier cost (and the shift-free plugin).  The
        # draws are not stored (2.9 MB per solve): tests regenerate them from the seed and check their digest.
        ("ode_qb_K65536_T100", "ODE", "quadratic_boundary", 65536, 100, 1, 0.0, 1.0),
        ("ode_gradmin_K65536_T100", "ODE", "quadratic_boundary_grad_minimal", 65536, 100, 1, 0.0, 1.0),
    ]
    only = [x for x in os.environ.get("CPS_GOLDEN_ONLY", "").split(",") if x]
    for (name, pred, cost, K, T, steps, tp, te) in runs:
        if only and name not in only:
            continue
        big = K >= 65536
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
        gen = torch.Generator().manual_seed(1)
        lib = R.torch_lib()
        if cost == "quadratic_boundary_grad":
            _patch_torch_lib_for_grad(lib)
        vp = R.variable_parameters(lib, tp, te)
        cw = R.cost_function(cost, lib, vp, K, T)
        Pred = R.ODEv0CoreAdapter if pred == "ODE_v0" else R.ODECoreAdapter
        vp_pred = None if pred == "ODE_v0" else vp
        predictor = Pred(T, K, 0.02, 10, vp_pred)
        opt = R.optimizer_mppi(predictor, cw, K, T, logging=True)
        n_ind = opt.Interpolator.number_of_interpolation_inducing_points
        draws = [torch.normal(0.0, 1.0, size=(K, n_ind, 1), generator=gen, dtype=torch.float32) for _ in range(steps)]
        opt.rng = R.InjectedNormal(draws)
        s = hanging_state()
        eps_all = np.stack([d.numpy()[:, :, 0] for d in draws], 0)
        arrays = {} if big else {"eps": eps_all}
        S, U, UNOM, JJ, UPREV = [], [], [], [], []
        for i in range(steps):
            UPREV.append(np.float32(opt.u))
            with torch.inference_mode():
                u = opt.step(s.copy())
            S.append(s.copy())
            U.append(np.float32(u))
            UNOM.append(opt.u_nom.numpy().reshape(-1).astype(np.float32))
            JJ.append(opt.logging_values["J_logged"].astype(np.float32))
            if i == 0:
                if not big:
                    arrays["u_run0"] = opt.logging_values["Q_logged"][:, :, 0].astype(np.float32)
                arrays["traj0"] = opt.logging_values["rollout_trajectories_logged"][:32].astype(np.float32)
            # advance the "plant" one control step with the applied control (input generation only)
            s = O.rollout("ODE", s, np.array([[u]], dtype=np.float32), n=10, dt=0.02)[0, 1]
        arrays.update(s=np.stack(S), u=np.array(U), u_nom=np.stack(UNOM), J=np.stack(JJ), u_prev=np.array(UPREV))
        meta = dict(ref="Control_Toolkit/Optimizers/optimizer_mppi.py:180-224 (torch lib, injected rng.normal draws)",
                    predictor=pred, cost=cost, K=K, T=T, steps=steps, target_position=tp, target_equilibrium=te,
                    n=10, dt=0.02, p=10, cc_weight=1.0, R=1.0, LBD=100.0, NU=1000.0, SQRTRHOINV=0.03)
        if big:
            import hashlib
            meta.update(eps_seed=1, eps_sha256=hashlib.sha256(eps_all.tobytes()).hexdigest(),
                        eps_recipe="gen = torch.Generator().manual_seed(eps_seed); per solve torch.normal(0.0, 1.0, "
                                   "size=(K, n_ind, 1), generator=gen, dtype=torch.float32)[:, :, 0]")
        save("mppi_" + name, meta, **arrays)


def main(which=None):
    if not R.available():
        raise SystemExit("reference tree not available; fixtures can only be regenerated in the build container")
    R.load()
    todo = dict(rollouts=gen_rollouts, costs=gen_costs, interp=gen_interp, mppi=gen_mppi)
    try:
        from oracle import gen_golden_net
        todo["net"] = gen_golden_net.gen_net
        todo["mppi_net"] = gen_golden_net.gen_mppi_net
        todo["net_diff"] = gen_golden_net.gen_net_diff
    except ImportError:
        pass
    for k, fn in todo.items():
        if which and k not in which:
            continue
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            try:
                fn()
            finally:
                txt = buf.getvalue()
        print("\n".join(l for l in txt.splitlines() if l.startswith("wrote")))


if __name__ == "__main__":
    main(sys.argv[1:] or None)
