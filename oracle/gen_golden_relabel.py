"""Golden vectors for offline relabelling: tests/golden/relabel_*.npz.

TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED add_control_along_trajectories
(SI_Toolkit/src/SI_Toolkit/General/preprocess_data_add_control_along_trajectories.py:53-140) over small recordings,
one call per file as the reference does, with its `controller_creator` hook returning a controller that does what
controller_mpc.step does (Control_Toolkit/Controllers/controller_mpc.py:102-109: update the variable parameters from
`updated_attributes`, then optimizer.step(s)) around the reference's optimizer_mppi (torch library, injected noise).
Every controller.step call is recorded (state, attributes, the draws it consumed, the control) so that the CUDA path
can be fed the same sequence; the returned DataFrame's label column is stored as the expected output.

    python oracle/gen_golden_relabel.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import ref_loader as R  # noqa: E402
from oracle.gen_golden import make_states, save  # noqa: E402

STATE_COLUMNS = ["angle", "angleD", "angle_cos", "angle_sin", "position", "positionD"]  # CartPole/state_utilities.py:5-23


class RecordingController:
    """controller_mpc.step around the reference optimizer_mppi, recording every call."""

    def __init__(self, opt, vp, draws, vp_np=None):
        self.opt, self.vp, self.draws, self.vp_np = opt, vp, draws, vp_np
        self.calls = []

    def reset(self):
        self.opt.optimizer_reset()

    def step(self, s, time=None, updated_attributes=None):
        import torch
        ua = {k: float(v) for k, v in (updated_attributes or {}).items()}
        self.vp.update_attributes({k: v for k, v in ua.items() if k in ("target_position", "target_equilibrium", "L", "m_pole")})
        if self.vp_np is not None and "L" in ua:  # predictor_ODE_v0 reads .L as is (predictors_customization_v0.py:47-48)
            self.vp_np.L = np.float32(ua["L"])
        s32 = np.asarray(s, dtype=np.float32)
        i = self.opt.rng.i
        with torch.inference_mode():
            u = self.opt.step(s32.copy())
        self.calls.append(dict(s=s32, tp=ua.get("target_position", 0.0), te=ua.get("target_equilibrium", 1.0),
                               L=ua.get("L", 0.395), eps=self.draws[i].numpy()[:, :, 0], u=np.float32(u)))
        return u


def make_file(rng, n_rows):
    import pandas as pd
    s = make_states(rng, n_rows, "random")
    df = pd.DataFrame({c: s[:, i].astype(np.float64) for i, c in enumerate(STATE_COLUMNS)})
    df.insert(0, "time", np.arange(n_rows) * 0.02)
    df["target_position"] = np.float32(rng.uniform(-0.15, 0.15)).astype(np.float64)
    df["target_equilibrium"] = np.where(np.arange(n_rows) < n_rows // 2, 1.0, -1.0)
    df["L"] = rng.uniform(0.25, 0.55, n_rows).astype(np.float32).astype(np.float64)
    df["Q_applied_-1"] = rng.uniform(-1, 1, n_rows)
    return df


def gen(name, pred, cost, K, T, n_files, n_rows, env_attrs, n_evals, method="monte_carlo"):
    import torch
    if "numdifftools" not in sys.modules:
        try:
            import numdifftools  # noqa: F401
        except Exception:
            sys.modules["numdifftools"] = types.ModuleType("numdifftools")
    from SI_Toolkit.General.preprocess_data_add_control_along_trajectories import add_control_along_trajectories
    import zlib
    # (the fixtures before nquad_* were drawn with a per-process salted str hash; they carry their own inputs)
    rng = np.random.default_rng(zlib.crc32(name.encode()) if method == "nquad" else abs(hash(name)) % (2 ** 31))
    arrays, n_calls = {}, None
    for f in range(n_files):
        df = make_file(rng, n_rows)
        lib = R.torch_lib()
        if cost == "quadratic_boundary_grad":
            from oracle.gen_golden import _patch_torch_lib_for_grad
            _patch_torch_lib_for_grad(lib)
        vp = R.variable_parameters(lib, 0.0, 1.0)
        cw = R.cost_function(cost, lib, vp, K, T)
        vp_np = None
        if pred == "ODE_v0":
            vp_np = type("VP", (), {})()
            vp_np.L = np.float32(0.395)
            predictor = R.ODEv0CoreAdapter(T, K, 0.02, 10, vp_np)
        else:
            predictor = R.ODECoreAdapter(T, K, 0.02, 10, vp)
        opt = R.optimizer_mppi(predictor, cw, K, T, logging=False)
        n_ind = opt.Interpolator.number_of_interpolation_inducing_points
        gen_ = torch.Generator().manual_seed(100 + f)
        draws = [torch.normal(0.0, 1.0, size=(K, n_ind, 1), generator=gen_, dtype=torch.float32)
                 for _ in range(n_rows * (21 * 2 * max(n_evals, 1) if method == "nquad" else max(n_evals, 1, 5)))]
        opt.rng = R.InjectedNormal(draws)
        ctrl = RecordingController(opt, vp, draws, vp_np)
        cfg = dict(state_components=STATE_COLUMNS, environment_attributes_dict=dict(env_attrs))
        out = add_control_along_trajectories(df.copy(), cfg, controller_creator=lambda c, a: ctrl,
                                             controller_output_variable_name="Q_calculated_offline",
                                             integration_method=method, integration_num_evals=n_evals)
        n_diff = sum(1 for v in env_attrs.values() if v.endswith("_differentiate_"))
        per_row = 5 * n_diff if n_diff else max(n_evals, 1)   # savgol window 5 per differentiated feature (:365-372)
        if method == "nquad":   # adaptive: at least one 21-point Gauss-Kronrod rule per row
            assert len(ctrl.calls) >= 21 * n_rows and len(ctrl.calls) % 21 == 0
            assert not np.isnan(out["Q_calculated_offline"].to_numpy(dtype=np.float64)).any()
            n_calls = max(n_calls or 0, len(ctrl.calls))
            arrays[f"f{f}__n_calls"] = np.array(len(ctrl.calls))
        else:
            assert len(ctrl.calls) == n_rows * per_row, "the reference swallowed an exception (it prints and goes on)"
            n_calls = len(ctrl.calls)
        arrays[f"f{f}__s"] = np.stack([c["s"] for c in ctrl.calls])
        arrays[f"f{f}__tp"] = np.array([c["tp"] for c in ctrl.calls], dtype=np.float64)
        arrays[f"f{f}__te"] = np.array([c["te"] for c in ctrl.calls], dtype=np.float64)
        arrays[f"f{f}__L"] = np.array([c["L"] for c in ctrl.calls], dtype=np.float64)
        arrays[f"f{f}__eps"] = np.stack([c["eps"] for c in ctrl.calls])
        arrays[f"f{f}__u"] = np.array([c["u"] for c in ctrl.calls], dtype=np.float32)
        for col in out.columns:
            if "calculated_offline" in col:
                arrays[f"f{f}__{col}"] = out[col].to_numpy(dtype=np.float64)
        arrays[f"f{f}__table"] = df[["time"] + STATE_COLUMNS + ["target_position", "target_equilibrium", "L"]].to_numpy()
    save("relabel_" + name, dict(ref="SI_Toolkit/General/preprocess_data_add_control_along_trajectories.py:53-140 driving "
                                     "optimizer_mppi (torch lib, injected draws) through a controller_mpc.step-shaped hook",
                                 predictor=pred, cost=cost, K=K, T=T, files=n_files, rows=n_rows, calls=n_calls,
                                 evals=n_evals, method=method, environment_attributes_dict=env_attrs,
                                 columns=["time"] + STATE_COLUMNS + ["target_position", "target_equilibrium", "L"]),
         **arrays)


def main():
    if not R.available():
        raise SystemExit("reference tree not available; fixtures can only be regenerated in the build container")
    R.load()
    if "nquad" in sys.argv[1:]:   # only the adaptive-quadrature fixture (the others carry per-process random inputs)
        integ = {"target_position": "target_position", "target_equilibrium": "target_equilibrium",
                 "L": "L_integrate_0.25_0.55_", "Q_ccrc": "Q_applied_-1"}
        with contextlib.redirect_stdout(io.StringIO()) as buf, contextlib.redirect_stderr(io.StringIO()):
            try:
                gen("nquad_ode", "ODE", "quadratic_boundary_grad_minimal", 128, 20, 2, 3, integ, 3, method="nquad")
            finally:
                txt = buf.getvalue()
        print("\n".join(l for l in txt.splitlines() if l.startswith("wrote") or "Error" in l))
        return
    plain = {"target_position": "target_position", "target_equilibrium": "target_equilibrium", "L": "L"}
    integ = {"target_position": "target_position", "target_equilibrium": "target_equilibrium",
             "L": "L_integrate_0.25_0.55_", "Q_ccrc": "Q_applied_-1"}
    with contextlib.redirect_stdout(io.StringIO()) as buf, contextlib.redirect_stderr(io.StringIO()):
        try:
            gen("plain_ode", "ODE", "quadratic_boundary_grad_minimal", 256, 30, 3, 10, plain, 0)
            gen("plain_v0", "ODE_v0", "quadratic_boundary_grad_minimal", 128, 30, 2, 8, plain, 0)
            gen("integrate_ode", "ODE", "quadratic_boundary_grad", 128, 20, 2, 5, integ, 4)
            diff = {"target_position": "target_position", "target_equilibrium": "target_equilibrium", "L": "L_differentiate_"}
            gen("differentiate_ode", "ODE", "quadratic_boundary_grad_minimal", 128, 20, 2, 4, diff, 0)
        finally:
            txt = buf.getvalue()
    print("\n".join(l for l in txt.splitlines() if l.startswith("wrote") or "Error" in l))


if __name__ == "__main__":
    main()
