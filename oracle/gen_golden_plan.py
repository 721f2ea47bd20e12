"""Golden vectors for the forward-only optimizers: tests/golden/plan_*.npz.

TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference classes optimizer_random_action_tf and optimizer_cem_tf
(Control_Toolkit/Optimizers/optimizer_random_action_tf.py, optimizer_cem_tf.py) in this container.  TensorFlow is not
installed here, so the handful of tf.* ops these two files call is supplied by oracle/tf_shim.py (torch CPU float32);
the optimizers' own Python runs as is, on the reference's torch predictor_ODE (or numba predictor_ODE_v0) and the
reference's cost plugins, with the generator's draws injected so that the CUDA path can be fed the same numbers.

    python oracle/gen_golden_plan.py [ra] [cem] [gmm] [rpgd]      (default: gmm)
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import tf_shim  # noqa: E402

tf_shim.install()  # before the reference is imported

from oracle import ref_loader as R  # noqa: E402
from oracle.gen_golden import hanging_state, make_states, save  # noqa: E402


def _make(cls_name, pred, cost, K, T, tp, te, rng=None, configure_kw=None, **params):
    import importlib
    import torch
    from SI_Toolkit.computation_library import TensorFlowLibrary
    lib = R.torch_lib()
    if cost == "quadratic_boundary_grad":
        # TF-only plugin in the reference (PyTorchLibrary has no stop_gradient / cond, computation_library.py:501): their torch
        # counterparts -- detach() is what tf.stop_gradient does to the tape
        lib.stop_gradient = lambda x: x.detach() if hasattr(x, "detach") else x
        lib.cond = lambda c, true_fn, false_fn: true_fn() if bool(c) else false_fn()
    vp = R.variable_parameters(lib, tp, te)
    cw = R.cost_function(cost, lib, vp, K, T)
    Pred = R.ODEv0CoreAdapter if pred == "ODE_v0" else R.ODECoreAdapter
    predictor = Pred(T, K, 0.02, 10, None if pred == "ODE_v0" else vp)
    mod = importlib.import_module("Control_Toolkit.Optimizers." + cls_name)
    cls = getattr(mod, cls_name)
    opt = cls(predictor=predictor, cost_function=cw,
              control_limits=(np.array([-1.0], dtype=np.float32), np.array([1.0], dtype=np.float32)),
              computation_library=TensorFlowLibrary(), seed=1, mpc_horizon=T, num_rollouts=K,
              optimizer_logging=True, calculate_optimal_trajectory=False, **params)
    opt.rng = rng if rng is not None else tf_shim.InjectedDraws([torch.zeros(K, T, 1)])  # optimizer_reset of random-action draws once
    opt.configure(num_states=6, num_control_inputs=1, **(configure_kw or {}))
    return opt


def gen_random_action():
    from oracle import oracle as O
    runs = [("ra_ode_gradmin", "ODE", "quadratic_boundary_grad_minimal", 640, 35, 4, 0.0, 1.0),
            ("ra_v0_qb", "ODE_v0", "quadratic_boundary", 256, 35, 3, 0.05, 1.0),
            ("ra_ode_default", "ODE", "default", 200, 20, 3, 0.0, -1.0)]
    for (name, pred, cost, K, T, steps, tp, te) in runs:
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
        opt = _make("optimizer_random_action_tf", pred, cost, K, T, tp, te)
        Qs = [rng.uniform(-1, 1, (K, T, 1)).astype(np.float32) for _ in range(steps)]
        opt.rng = tf_shim.InjectedDraws(Qs)
        s = make_states(rng, 1, "random")[0] if "default" in name else hanging_state()
        S, U, JJ, UP = [], [], [], []
        for i in range(steps):
            UP.append(np.float32(opt.u))
            u = opt.step(s.copy())
            S.append(s.copy()); U.append(np.float32(u)); JJ.append(opt.logging_values["J_logged"].astype(np.float32))
            s = O.rollout("ODE", s, np.array([[u]], dtype=np.float32))[0, 1]
        save("plan_" + name, dict(ref="Control_Toolkit/Optimizers/optimizer_random_action_tf.py:52-79 (tf ops from "
                                      "oracle/tf_shim.py, injected uniform draws)", predictor=pred, cost=cost, K=K, T=T,
                                  steps=steps, target_position=tp, target_equilibrium=te),
             Q=np.stack([q[:, :, 0] for q in Qs]), s=np.stack(S), u=np.array(U), J=np.stack(JJ), u_prev=np.array(UP))


def gen_cem():
    from oracle import oracle as O
    runs = [  # name, predictor, cost, K, T, steps, outer iterations, best_k, tp, te
        ("cem_ode_gradmin", "ODE", "quadratic_boundary_grad_minimal", 200, 35, 4, 3, 40, 0.0, 1.0),
        ("cem_v0_gradmin", "ODE_v0", "quadratic_boundary_grad_minimal", 200, 35, 3, 3, 40, 0.0, 1.0),
        ("cem_ode_qb", "ODE", "quadratic_boundary", 256, 20, 3, 2, 17, 0.05, 1.0),
        ("cem_ode_K1024", "ODE", "quadratic_boundary_grad_minimal", 1024, 40, 2, 2, 100, 0.0, 1.0),
    ]
    for (name, pred, cost, K, T, steps, iters, best_k, tp, te) in runs:
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
        opt = _make("optimizer_cem_tf", pred, cost, K, T, tp, te, cem_outer_it=iters, cem_initial_action_stdev=0.5,
                    cem_stdev_min=0.01, cem_best_k=best_k, warmup=False, warmup_iterations=250)
        eps = rng.standard_normal((steps, iters, K, T)).astype(np.float32)
        opt.rng = tf_shim.InjectedDraws([eps[i, j][:, :, None] for i in range(steps) for j in range(iters)])
        s = hanging_state()
        S, U, JJ, QQ, UP, MU, SD = [], [], [], [], [], [], []
        for i in range(steps):
            UP.append(np.float32(opt.u))
            u = opt.step(s.copy())
            S.append(s.copy()); U.append(np.float32(u))
            JJ.append(opt.logging_values["J_logged"].astype(np.float32))
            QQ.append(opt.logging_values["Q_logged"][:, :, 0].astype(np.float32))
            MU.append(opt.dist_mue.numpy().reshape(-1).astype(np.float32))
            SD.append(opt.stdev.numpy().reshape(-1).astype(np.float32))
            s = O.rollout("ODE", s, np.array([[u]], dtype=np.float32))[0, 1]
        save("plan_" + name, dict(ref="Control_Toolkit/Optimizers/optimizer_cem_tf.py:63-109 (tf ops from "
                                      "oracle/tf_shim.py, injected normal draws)", predictor=pred, cost=cost, K=K, T=T,
                                  steps=steps, iterations=iters, best_k=best_k, initial_stdev=0.5, stdev_min=0.01,
                                  target_position=tp, target_equilibrium=te),
             eps=eps, s=np.stack(S), u=np.array(U), J=np.stack(JJ), Q=np.stack(QQ), u_prev=np.array(UP),
             mean=np.stack(MU), stdev=np.stack(SD))


def gen_cem_gmm():
    import zlib
    from oracle import oracle as O
    runs = [  # name, predictor, cost, K, T, steps, outer iterations, best_k, tp, te
        ("gmm_ode_gradmin", "ODE", "quadratic_boundary_grad_minimal", 200, 35, 4, 3, 40, 0.0, 1.0),
        ("gmm_v0_gradmin", "ODE_v0", "quadratic_boundary_grad_minimal", 200, 35, 3, 3, 40, 0.0, 1.0),
        ("gmm_ode_K512", "ODE", "quadratic_boundary_grad_minimal", 512, 25, 3, 2, 33, 0.05, 1.0),
    ]
    for (name, pred, cost, K, T, steps, iters, best_k, tp, te) in runs:
        rng = np.random.default_rng(zlib.crc32(name.encode()))   # reproducible (str hashes are salted per process)
        opt = _make("optimizer_cem_gmm_tf", pred, cost, K, T, tp, te, cem_outer_it=iters, cem_initial_action_stdev=0.5,
                    cem_stdev_min=0.01, cem_best_k=best_k)
        eps = rng.standard_normal((steps, iters, K, T, 2)).astype(np.float32)
        u01 = rng.uniform(0, 1, (steps, iters, K, T)).astype(np.float32)
        draws = []
        for i in range(steps):
            for j in range(iters):
                draws += [eps[i, j][:, :, None, :], u01[i, j][:, :, None]]
        tf_shim.GMM_DRAWS = tf_shim.InjectedDraws(draws)
        s = hanging_state()
        S, U, JJ, QQ, UP, LOC, SC, P1 = [], [], [], [], [], [], [], []
        for i in range(steps):
            UP.append(np.float32(opt.u))
            u = opt.step(s.copy())
            S.append(s.copy()); U.append(np.float32(u))
            JJ.append(opt.logging_values["J_logged"].astype(np.float32))
            QQ.append(opt.logging_values["Q_logged"][:, :, 0].astype(np.float32))
            cd = opt.sampling_dist.components_distribution
            LOC.append(cd.mean().numpy().reshape(T, 2).astype(np.float32))
            SC.append(cd.stddev().numpy().reshape(T, 2).astype(np.float32))
            P1.append(np.float32(opt.sampling_dist.mixture_distribution.probs[0]))
            s = O.rollout("ODE", s, np.array([[u]], dtype=np.float32))[0, 1]
        save("plan_" + name, dict(ref="Control_Toolkit/Optimizers/optimizer_cem_gmm_tf.py:58-131 (tf / tfpd ops from "
                                      "oracle/tf_shim.py, injected normal and uniform draws)", predictor=pred, cost=cost, K=K,
                                  T=T, steps=steps, iterations=iters, best_k=best_k, initial_stdev=0.5, stdev_min=0.01,
                                  target_position=tp, target_equilibrium=te),
             eps=eps, u01=u01, s=np.stack(S), u=np.array(U), J=np.stack(JJ), Q=np.stack(QQ), u_prev=np.array(UP),
             loc=np.stack(LOC), scale=np.stack(SC), p1=np.array(P1))


def gen_rpgd():
    """optimizer_rpgd_tf (:15-420) on the shim: torch autograd stands in for the GradientTape, _LegacyAdam for Keras' Adam."""
    import zlib
    import torch
    from oracle import oracle as O
    GM, GR = "quadratic_boundary_grad_minimal", "quadratic_boundary_grad"
    runs = [  # name, K, T, steps, outer_its, resamp_per, keep ratio, period of the inducing points, tp, te, cost plugin
        ("rpgd_default", 16, 35, 12, 4, 10, 0.75, 4, 0.0, 1.0, GM),      # the shipped configuration (config_optimizers.yml:63-85)
        ("rpgd_resamp3", 24, 20, 8, 2, 3, 0.5, 5, 0.05, 1.0, GM),
        ("rpgd_qbgrad", 16, 35, 6, 4, 4, 0.75, 4, 0.0, 1.0, GR),         # the same optimizer on the other gradient cost plugin
    ]
    only = os.environ.get("CPS_GOLDEN_ONLY")   # regenerate one fixture without touching the others
    for (name, K, T, steps, its, resamp, ratio, p, tp, te, cost) in runs:
        if only and name != only:
            continue
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        n_ind = int(np.ceil((T - 1) / p)) + 1
        keep = int(max(int(K * ratio), 1))
        n_resamp = sum(1 for c in range(steps) if c % resamp == 0)
        draws = [rng.standard_normal((K, n_ind, 1)).astype(np.float32)] + \
                [rng.standard_normal((K - keep, n_ind, 1)).astype(np.float32) for _ in range(n_resamp)]
        opt = _make("optimizer_rpgd_tf", "ODE", cost, K, T, tp, te,
                    rng=tf_shim.InjectedDraws(draws), configure_kw=dict(dt=0.02, predictor_specification="ODE"),   # optimizer_reset: first draw
                    outer_its=its, sample_stdev=0.5,
                    sample_mean=0.0, sample_whole_control_space=False, uniform_dist_min=-0.8, uniform_dist_max=0.8,
                    resamp_per=resamp, period_interpolation_inducing_points=p, SAMPLING_DISTRIBUTION="normal",
                    shift_previous=1, warmup=False, warmup_iterations=250, learning_rate=0.05, opt_keep_k_ratio=ratio,
                    gradmax_clip=5, rtol=1e-3, adam_beta_1=0.9, adam_beta_2=0.999, adam_epsilon=1e-8)
        s = hanging_state()
        S, U, JJ, QQ, UP, QN, M1, V1, IT = [], [], [], [], [], [], [], [], []
        for i in range(steps):
            UP.append(np.float32(np.asarray(opt.u).reshape(-1)[0]))
            u = opt.step(s.copy())
            S.append(s.copy()); U.append(np.float32(np.asarray(u).reshape(-1)[0]))
            JJ.append(opt.logging_values["J_logged"].astype(np.float32))
            QQ.append(opt.logging_values["Q_logged"][:, :, 0].astype(np.float32))      # plans after the gradient steps
            QN.append(opt.Q_tf.detach().numpy()[:, :, 0].astype(np.float32).copy())      # warm start for the next solve
            w = opt.opt.get_weights()
            IT.append(int(w[0])); M1.append(w[1].numpy()[:, :, 0].copy()); V1.append(w[2].numpy()[:, :, 0].copy())
            s = O.rollout("ODE", s, np.array([[float(np.asarray(u).reshape(-1)[0])]], dtype=np.float32))[0, 1]
        save("plan_" + name, dict(ref="Control_Toolkit/Optimizers/optimizer_rpgd_tf.py:149-408 (tf ops, GradientTape and Keras "
                                      "legacy Adam from oracle/tf_shim.py; injected normal draws)", predictor="ODE",
                                  cost=cost, K=K, T=T, steps=steps, outer_its=its, resamp_per=resamp,
                                  opt_keep_k_ratio=ratio, period_interpolation_inducing_points=p, learning_rate=0.05,
                                  gradmax_clip=5.0, sample_stdev=0.5, target_position=tp, target_equilibrium=te),
             draw0=draws[0][:, :, 0], resamp_draws=np.stack([d[:, :, 0] for d in draws[1:]]), s=np.stack(S), u=np.array(U),
             J=np.stack(JJ), Q=np.stack(QQ), Q_next=np.stack(QN), u_prev=np.array(UP), adam_m=np.stack(M1), adam_v=np.stack(V1),
             adam_iterations=np.array(IT))


def main():
    if not R.available():
        raise SystemExit("reference tree not available; fixtures can only be regenerated in the build container")
    R.load()
    # the ra_* / cem_* fixtures were drawn with per-process seeds (salted str hash): they carry their own inputs and stay
    # valid recordings, but regenerating them gives different draws -- pass their names to regenerate deliberately
    which = set(sys.argv[1:]) or {"gmm"}
    fns = [f for f, tag in ((gen_random_action, "ra"), (gen_cem, "cem"), (gen_cem_gmm, "gmm"), (gen_rpgd, "rpgd")) if tag in which]
    for fn in fns:
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            try:
                fn()
            finally:
                txt = buf.getvalue()
        print("\n".join(l for l in txt.splitlines() if l.startswith("wrote")))


if __name__ == "__main__":
    main()
