#!/usr/bin/env python
"""Timing of the gradient of predict_and_cost (cps_plan_cost_grad: forward / per-step Jacobians / reverse) and of
optimizer_rpgd_b200.step at the reference's shipped RPGD configuration (16 plans x 35 steps, 4 Adam steps per solve)."""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import cartpolesimulation_b200 as cps                                       # noqa: E402
from cartpolesimulation_b200.optimizer_forward_b200 import optimizer_rpgd_b200   # noqa: E402


def stream_time(fn, reps=20, rounds=15):
    ts = []
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return float(np.median(ts))


def main():
    a = np.pi - 1e-3
    s = np.array([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], np.float32)
    lim = (np.array([-1.0], np.float32), np.array([1.0], np.float32))
    for K, T in ((16, 35), (32, 35), (2000, 50), (65536, 50)):
        vp = cps.VariableParameters(target_position=0.0, target_equilibrium=1.0, L=0.395, m_pole=0.087)
        cost, pred = cps.CostFunctionWrapper(), cps.PredictorWrapper()
        opt = optimizer_rpgd_b200(predictor=pred, cost_function=cost, control_limits=lim, seed=1, mpc_horizon=T, num_rollouts=K,
                                  device=0, outer_its=4)
        pred.configure(batch_size=K, horizon=T, dt=0.02, variable_parameters=vp, predictor_specification="ODE")
        cost.configure(batch_size=K, horizon=T, variable_parameters=vp, environment_name="CartPole", computation_library=None,
                       cost_function_specification="quadratic_boundary_grad_minimal")
        opt.configure(num_states=6, num_control_inputs=1, dt=0.02, predictor_specification="ODE")
        for _ in range(20):
            opt.step(s)
        lat = []
        for _ in range(200):
            t0 = time.perf_counter()
            opt.step(s)
            lat.append((time.perf_counter() - t0) * 1e3)
        Q = torch.zeros((K, T), device=opt.device).uniform_(-0.5, 0.5)
        s_dev = torch.from_numpy(s).to(opt.device)
        g = stream_time(lambda: opt.engine.plan_cost_grad(s_dev, Q))
        print(f"RPGD K={K} T={T}: gradient {g * 1e3:.1f} us in stream ({K * T * 10 / g * 1e3:.3e} state-steps/s forward-equivalent), "
              f"optimizer step (4 Adam steps) median {np.median(lat):.3f} ms p99 {np.percentile(lat, 99):.3f} ms")
        opt.engine.close()


if __name__ == "__main__":
    main()
