"""Golden vectors for the autoregressive neural predictor (GRU / Dense), produced by the UNMODIFIED reference
predictor_autoregressive_neural (SI_Toolkit/src/SI_Toolkit/Predictors/predictor_autoregressive_neural.py) with the
torch Sequence network (Functions/Pytorch/Network.py).

No dynamics model ships with the reference (config_predictors.yml:10-11,32-33 point to a directory that is not in
the tree), so the weights are synthetic: default torch init under torch.manual_seed, saved through
Sequence.state_dict() into a scratch model directory with a hand-written net-info .txt (format:
Functions/General/Initialization.py:35-104) and a normalisation table with the value ranges of the shipped
GymlikeCartPole/Dense-7IN-32H1-32H2-1OUT-0/NI_2024-08-17_22-23-01.csv.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import contextlib
import io
import os
import shutil
import tempfile

import numpy as np

from oracle import ref_loader as R
from oracle.gen_golden import hanging_state, make_states, save

INPUTS = ["Q", "angleD", "angle_cos", "angle_sin", "position", "positionD"]
OUTPUTS = ["angleD", "angle_cos", "angle_sin", "position", "positionD"]
# mean, std, max, min rows (Normalising.py header); minmax_sym uses max/min only
NORM = {"Q": (0, 0.5, 1.0, -1.0), "angle": (0, 1.8, np.pi, -np.pi), "angleD": (0, 4.0, 18.38, -18.38),
        "angle_cos": (0, 0.7, 1.0, -1.0), "angle_sin": (0, 0.7, 1.0, -1.0), "position": (0, 0.08, 0.198, -0.198),
        "positionD": (0, 0.3, 1.125, -1.125)}


# differential networks (outputs D_*): value ranges of the derivatives for the normalisation table
NORM_D = {"D_angle": (0, 4.0, 18.38, -18.38), "D_angleD": (0, 20.0, 90.0, -90.0), "D_angle_cos": (0, 3.0, 15.0, -15.0),
          "D_angle_sin": (0, 3.0, 15.0, -15.0), "D_position": (0, 0.3, 1.125, -1.125), "D_positionD": (0, 2.0, 9.0, -9.0)}


def write_model_dir(root, full_name, net_type_name, seed, weight_scale=1.0, INPUTS=INPUTS, OUTPUTS=OUTPUTS, NORM=NORM):
    import torch
    from SI_Toolkit.Functions.Pytorch.Network import Sequence
    d = os.path.join(root, full_name)
    os.makedirs(d, exist_ok=True)
    short = "-".join(p for p in full_name.split("-") if not (p.endswith("IN") or p.endswith("OUT")))[:-2]  # drop index
    ni = os.path.join(d, "NI.csv")
    cols = list(NORM.keys())
    with open(ni, "w") as f:
        f.write("," + ",".join(cols) + "\n")
        for r, rowname in enumerate(["mean", "std", "max", "min"]):
            f.write(rowname + "," + ",".join(repr(float(NORM[c][r])) for c in cols) + "\n")
    with open(os.path.join(d, full_name + ".txt"), "w") as f:
        f.write("\n".join([
            "CREATED:", "2026-01-01 00:00:00", "", "LIBRARY:", "Pytorch", "", "NET NAME:", short, "",
            "NET FULL NAME:", full_name, "", "INPUTS:", ", ".join(INPUTS), "", "OUTPUTS:", ", ".join(OUTPUTS), "",
            "TYPE:", net_type_name, "", "NORMALIZATION:", ni, "", "NORMALIZE:", "True", "",
            "WASH OUT LENGTH:", "10", "", "CONSTRUCT NETWORK:", "with cells", ""]))
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = Sequence(short, INPUTS, OUTPUTS, batch_size=1, construct_network="with cells")
    if weight_scale != 1.0:
        with torch.no_grad():
            for p in net.parameters():
                p.mul_(weight_scale)
    torch.save(net.state_dict(), os.path.join(d, "ckpt.pt"))
    return {k: v.detach().cpu().numpy().astype(np.float32) for k, v in net.state_dict().items()}


def gen_net():
    import torch
    R.load()
    from SI_Toolkit.Predictors.predictor_autoregressive_neural import predictor_autoregressive_neural
    rng = np.random.default_rng(99)
    root = tempfile.mkdtemp(prefix="cps_models_")
    try:
        for full_name, tname, seed, scale in (("GRU-6IN-64H1-64H2-5OUT-0", "GRU", 7, 1.0),
                                              ("GRU-6IN-32H1-32H2-5OUT-0", "GRU", 8, 2.0),
                                              ("Dense-6IN-32H1-32H2-5OUT-0", "Dense", 9, 1.5)):
            sd = write_model_dir(root, full_name, tname, seed, scale)
            K, T = 96, 50
            with contextlib.redirect_stdout(io.StringIO()), torch.inference_mode():
                pred = predictor_autoregressive_neural(model_name=full_name, path_to_model=root + os.sep, horizon=T,
                                                       dt=0.02, batch_size=K, disable_individual_compilation=True,
                                                       update_before_predicting=False)
                s_single = hanging_state()
                s0 = np.tile(s_single, (K, 1))
                Q = np.clip(rng.normal(0, 0.3, (K, T, 1)), -1, 1).astype(np.float32)
                arrays = {"state_dict_keys": np.array(list(sd.keys()))}
                for k, v in sd.items():
                    arrays["w__" + k] = v
                # (1) rollout from zero hidden state
                out1 = pred.predict_core(torch.from_numpy(s0), torch.from_numpy(Q)).numpy().astype(np.float32)
                # (2) advance the stored hidden state with 3 applied controls (update_internal_state_tf), then roll out
                s_seq, q_seq = [], []
                s_cur = s0.copy()
                for i in range(3):
                    q0 = np.full((K, 1, 1), 0.2 * (i + 1) - 0.3, dtype=np.float32)
                    pred.update_internal_state_tf(torch.from_numpy(q0), torch.from_numpy(s_cur))
                    s_seq.append(s_cur[0].copy())
                    q_seq.append(float(q0[0, 0, 0]))
                    s_cur = s_cur.copy()
                    s_cur[:, 1] += 0.05  # a slightly different measured state each tick
                h_after = None
                if tname == "GRU":
                    h_after = np.stack([h[0].numpy() for h in pred.memory_states_ref[0]], 0).astype(np.float32)
                out2 = pred.predict_core(torch.from_numpy(s_cur), torch.from_numpy(Q)).numpy().astype(np.float32)
                # (3) per-rollout different initial states
                s_rand = make_states(rng, K, "random")
                out3 = pred.predict_core(torch.from_numpy(s_rand), torch.from_numpy(Q)).numpy().astype(np.float32)
            arrays.update(s0=s_single, Q=Q[:, :, 0], traj_zero_h=out1, upd_s=np.stack(s_seq),
                          upd_q=np.array(q_seq, dtype=np.float32), s_after=s_cur[0], traj_after_updates=out2,
                          s_rand=s_rand, traj_rand=out3,
                          norm_table=np.array([[NORM[c][r] for c in NORM] for r in range(4)], dtype=np.float64),
                          norm_cols=np.array(list(NORM.keys())))
            if h_after is not None:
                arrays["h_after_updates"] = h_after
            meta = dict(ref="SI_Toolkit/Predictors/predictor_autoregressive_neural.py:266-313,332-352 (torch Sequence, "
                            "synthetic seeded weights)", net=full_name, type=tname, inputs=INPUTS, outputs=OUTPUTS,
                        K=K, T=T, seed=seed, weight_scale=scale)
            save("net_" + full_name.replace("-", "_"), meta, **arrays)
    finally:
        shutil.rmtree(root, ignore_errors=True)


DIFF_NETS = (
    # full name, type, seed, weight scale, inputs, outputs
    ("GRU-6IN-32H1-32H2-5OUT-1", "GRU", 21, 1.5,
     ["Q", "position", "positionD", "angle_cos", "angle_sin", "angleD"],            # permuted w.r.t. the outputs
     ["D_angleD", "D_angle_cos", "D_angle_sin", "D_position", "D_positionD"]),
    ("Dense-5IN-32H1-32H2-4OUT-1", "Dense", 22, 1.5,
     ["Q", "angle", "angleD", "position", "positionD"],                            # angle output: sin/cos are augmented
     ["D_angle", "D_angleD", "D_position", "D_positionD"]),
    ("GRU-6IN-64H1-64H2-5OUT-1", "GRU", 23, 1.0, INPUTS, ["D_" + o for o in OUTPUTS]),
)


def gen_net_diff():
    """Differential networks (outputs named D_*): predictor_autoregressive_neural with the
    differential_model_autoregression_helper (Predictors/autoregression.py:118-158, Normalising.py:111-186), the
    unmodified reference on synthetic seeded weights.  Also the horizon == 1 case, which the reference's loop routes
    around the helper (autoregression.py:49-70)."""
    import torch
    R.load()
    from SI_Toolkit.Predictors.predictor_autoregressive_neural import predictor_autoregressive_neural
    rng = np.random.default_rng(123)
    norm = dict(NORM)
    norm.update(NORM_D)
    root = tempfile.mkdtemp(prefix="cps_models_")
    try:
        for full_name, tname, seed, scale, inputs, outputs in DIFF_NETS:
            sd = write_model_dir(root, full_name, tname, seed, scale, inputs, outputs, norm)
            K, T = 64, 50
            arrays = {"state_dict_keys": np.array(list(sd.keys()))}
            for k, v in sd.items():
                arrays["w__" + k] = v
            with contextlib.redirect_stdout(io.StringIO()), torch.inference_mode():
                pred = predictor_autoregressive_neural(model_name=full_name, path_to_model=root + os.sep, horizon=T,
                                                       dt=0.02, batch_size=K, disable_individual_compilation=True,
                                                       update_before_predicting=False)
                assert pred.differential_network
                s_single = hanging_state()
                s0 = np.tile(s_single, (K, 1))
                Q = np.clip(rng.normal(0, 0.3, (K, T, 1)), -1, 1).astype(np.float32)
                out1 = pred.predict_core(torch.from_numpy(s0), torch.from_numpy(Q)).numpy().astype(np.float32)
                s_cur = s0.copy()
                s_seq, q_seq = [], []
                for i in range(2):
                    q0 = np.full((K, 1, 1), 0.25 * (i + 1) - 0.3, dtype=np.float32)
                    pred.update_internal_state_tf(torch.from_numpy(q0), torch.from_numpy(s_cur))
                    s_seq.append(s_cur[0].copy())
                    q_seq.append(float(q0[0, 0, 0]))
                    s_cur = s_cur.copy()
                    s_cur[:, 1] += 0.05
                s_rand = make_states(rng, K, "random")
                out3 = pred.predict_core(torch.from_numpy(s_rand), torch.from_numpy(Q)).numpy().astype(np.float32)
                h_after = None
                if tname == "GRU":
                    h_after = np.stack([h[0].numpy() for h in pred.memory_states_ref[0]], 0).astype(np.float32)
                # horizon 1: a second predictor object (the horizon is fixed at construction)
                pred1 = predictor_autoregressive_neural(model_name=full_name, path_to_model=root + os.sep, horizon=1,
                                                        dt=0.02, batch_size=K, disable_individual_compilation=True,
                                                        update_before_predicting=False)
                out_h1 = pred1.predict_core(torch.from_numpy(s_rand), torch.from_numpy(Q[:, :1])).numpy().astype(np.float32)
            arrays.update(s0=s_single, Q=Q[:, :, 0], traj_zero_h=out1, upd_s=np.stack(s_seq),
                          upd_q=np.array(q_seq, dtype=np.float32), s_rand=s_rand, traj_rand=out3, traj_rand_T1=out_h1,
                          norm_table=np.array([[norm[c][r] for c in norm] for r in range(4)], dtype=np.float64),
                          norm_cols=np.array(list(norm.keys())))
            if h_after is not None:
                arrays["h_after_updates"] = h_after
            meta = dict(ref="SI_Toolkit/Predictors/predictor_autoregressive_neural.py:266-313 + autoregression.py:118-158 + "
                            "Functions/General/Normalising.py:111-186 (torch Sequence, synthetic seeded weights, differential)",
                        net=full_name, type=tname, inputs=inputs, outputs=outputs, K=K, T=T, seed=seed, weight_scale=scale,
                        dt=0.02, differential=True)
            save("net_diff_" + full_name.replace("-", "_"), meta, **arrays)
    finally:
        shutil.rmtree(root, ignore_errors=True)


class NeuralCoreAdapter:
    """PredictorWrapper-shaped holder around the unmodified predictor_autoregressive_neural, with the wrapper's
    `update` (SI_Toolkit/src/SI_Toolkit/Predictors/predictor_wrapper.py:173-177)."""

    def __init__(self, predictor):
        self.predictor = predictor
        self.predictor_type = "neural"
        self.num_states, self.num_control_inputs = 6, 1
        self.horizon, self.batch_size = predictor.horizon, predictor.batch_size

    def predict_core(self, s, Q):
        return self.predictor.predict_core(s, Q)

    def update(self, Q0, s):
        import torch
        self.predictor.update_internal_state_tf(s=torch.as_tensor(s, dtype=torch.float32),
                                                Q0=torch.as_tensor(Q0, dtype=torch.float32))

    def copy(self):
        return R._NullPredictor()

    def configure(self, **kw):
        pass


def gen_mppi_net():
    """optimizer_mppi (unmodified, torch lib, injected noise) driving predictor_autoregressive_neural: a closed loop
    of several solves, so the hidden-state update after every solve (optimizer_mppi.py:191,194-196) is pinned too."""
    import torch
    from oracle import oracle as O
    R.load()
    from SI_Toolkit.Predictors.predictor_autoregressive_neural import predictor_autoregressive_neural
    root = tempfile.mkdtemp(prefix="cps_models_")
    try:
        norm_d = dict(NORM)
        norm_d.update(NORM_D)
        only = [x for x in os.environ.get("CPS_GOLDEN_ONLY", "").split(",") if x]
        for (run, full_name, tname, seed, scale, cost, K, T, steps, inputs, outputs, norm) in (
                ("gru64_gradmin", "GRU-6IN-64H1-64H2-5OUT-0", "GRU", 7, 1.0, "quadratic_boundary_grad_minimal", 256, 50, 4, INPUTS, OUTPUTS, NORM),
                ("gru32_grad", "GRU-6IN-32H1-32H2-5OUT-0", "GRU", 8, 2.0, "quadratic_boundary_grad", 128, 35, 3, INPUTS, OUTPUTS, NORM),
                ("dense32_gradmin", "Dense-6IN-32H1-32H2-5OUT-0", "Dense", 9, 1.5, "quadratic_boundary_grad_minimal", 128, 50, 3, INPUTS, OUTPUTS, NORM),
                # differential network (a15) under the optimizer, incl. the post-solve hidden-state update
                ("diff_gru32_gradmin", DIFF_NETS[0][0], "GRU", DIFF_NETS[0][2], DIFF_NETS[0][3], "quadratic_boundary_grad_minimal",
                 128, 50, 3, DIFF_NETS[0][4], DIFF_NETS[0][5], norm_d),
                # the MAX_COST plugin with the neural predictor (backend-ordered cost mean in net_kernel)
                ("gru32_qb", "GRU-6IN-32H1-32H2-5OUT-0", "GRU", 8, 2.0, "quadratic_boundary", 128, 50, 2, INPUTS, OUTPUTS, NORM)):
            if only and run not in only:
                continue
            NORM_RUN = norm
            sd = write_model_dir(root, full_name, tname, seed, scale, inputs, outputs, norm)
            lib = R.torch_lib()
            if cost == "quadratic_boundary_grad":
                from oracle.gen_golden import _patch_torch_lib_for_grad
                _patch_torch_lib_for_grad(lib)
            vp = R.variable_parameters(lib, 0.0, 1.0)
            cw = R.cost_function(cost, lib, vp, K, T)
            with contextlib.redirect_stdout(io.StringIO()):
                pred = predictor_autoregressive_neural(model_name=full_name, path_to_model=root + os.sep, horizon=T,
                                                       dt=0.02, batch_size=K, disable_individual_compilation=True,
                                                       update_before_predicting=False)
            opt = R.optimizer_mppi(NeuralCoreAdapter(pred), cw, K, T, logging=True)
            n_ind = opt.Interpolator.number_of_interpolation_inducing_points
            gen = torch.Generator().manual_seed(3)
            draws = [torch.normal(0.0, 1.0, size=(K, n_ind, 1), generator=gen, dtype=torch.float32) for _ in range(steps)]
            opt.rng = R.InjectedNormal(draws)
            s = hanging_state()
            arrays = {"eps": np.stack([d.numpy()[:, :, 0] for d in draws], 0), "state_dict_keys": np.array(list(sd.keys()))}
            for k, v in sd.items():
                arrays["w__" + k] = v
            S, U, UNOM, JJ, UPREV, HH = [], [], [], [], [], []
            for i in range(steps):
                UPREV.append(np.float32(opt.u))
                with torch.inference_mode(), contextlib.redirect_stdout(io.StringIO()):
                    u = opt.step(s.copy())
                S.append(s.copy())
                U.append(np.float32(u))
                UNOM.append(opt.u_nom.numpy().reshape(-1).astype(np.float32))
                JJ.append(opt.logging_values["J_logged"].astype(np.float32))
                if tname == "GRU":  # hidden state AFTER the post-solve update, row 0 (all rows are identical)
                    HH.append(np.concatenate([h[0].numpy().reshape(-1) for h in pred.memory_states_ref[0]]).astype(np.float32))
                if i == 0:
                    arrays["u_run0"] = opt.logging_values["Q_logged"][:, :, 0].astype(np.float32)
                    arrays["traj0"] = opt.logging_values["rollout_trajectories_logged"][:32].astype(np.float32)
                s = O.rollout("ODE", s, np.array([[u]], dtype=np.float32), n=10, dt=0.02)[0, 1]
            arrays.update(s=np.stack(S), u=np.array(U), u_nom=np.stack(UNOM), J=np.stack(JJ), u_prev=np.array(UPREV),
                          norm_table=np.array([[NORM_RUN[c][r] for c in NORM_RUN] for r in range(4)], dtype=np.float64),
                          norm_cols=np.array(list(NORM_RUN.keys())))
            if HH:
                arrays["h_after"] = np.stack(HH)
            meta = dict(ref="Control_Toolkit/Optimizers/optimizer_mppi.py:180-224 + predictor_autoregressive_neural.py:266-352 "
                            "(torch lib, injected rng.normal draws, synthetic seeded weights)",
                        net=full_name, type=tname, inputs=inputs, outputs=outputs, predictor="neural", cost=cost, K=K, T=T,
                        differential=any(o.startswith("D_") for o in outputs),
                        steps=steps, target_position=0.0, target_equilibrium=1.0, dt=0.02, p=10, cc_weight=1.0, R=1.0,
                        LBD=100.0, NU=1000.0, SQRTRHOINV=0.03, seed=seed, weight_scale=scale)
            save("mppi_net_" + run, meta, **arrays)
    finally:
        shutil.rmtree(root, ignore_errors=True)
