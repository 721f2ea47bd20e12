#!/usr/bin/env python
"""Kernel timing of the forward-only planners (plan_kernel): one CEM solve = `it` launches of K plans x T steps; also the
ncu target for plan_kernel.  CUDA events on the launching stream."""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=2000)
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--it", type=int, default=3)
    ap.add_argument("--best_k", type=int, default=0)
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    import torch
    from bench import _event_times, N_SUB
    from cartpolesimulation_b200 import _lib as L
    from cartpolesimulation_b200.core import Engine
    K, T = args.K, args.T
    eng = Engine(K, T, integrator="ODE", cost="quadratic_boundary_grad_minimal", device=0)
    eng.cem_configure(args.best_k or max(1, K // 10), 0.5, 0.01)
    a = np.pi - 1e-3
    s = torch.tensor([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], device="cuda", dtype=torch.float32)
    eps = torch.randn((args.it, T, K), device="cuda")
    Q = torch.empty((T, K), device="cuda").uniform_(-1, 1)
    cem = float(np.median(_event_times(lambda: eng.cem_step(s, eps, L.TIME_MAJOR, 0.0), args.iters)))
    ra = float(np.median(_event_times(lambda: eng.plan_random_action(s, Q, L.TIME_MAJOR, 0.0), args.iters)))
    pc = float(np.median(_event_times(lambda: eng.plan_cost(s, Q, L.TIME_MAJOR, 0.0), args.iters)))
    print(json.dumps({"K": K, "T": T, "cem_outer_it": args.it, "cem_solve_kernel_ms": cem,
                      "cem_state_steps_per_s": args.it * K * T * N_SUB / (cem * 1e-3),
                      "random_action_kernel_ms": ra, "plan_cost_kernel_ms": pc,
                      "plan_cost_state_steps_per_s": K * T * N_SUB / (pc * 1e-3)}))


if __name__ == "__main__":
    main()
