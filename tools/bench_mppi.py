#!/usr/bin/env python
"""Kernel timing of one ODE MPPI solve (cps_mppi_step) at a given K, T -- the ncu target for mppi_kernel."""
import argparse
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=2000)
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--integrator", default="ODE")
    ap.add_argument("--cost", default="quadratic_boundary_grad_minimal")
    ap.add_argument("--no-pairs", action="store_true", help="one rollout per thread at every K (A/B of the packed solve)")
    args = ap.parse_args()
    from cartpolesimulation_b200.core import Engine
    eng = Engine(args.K, args.T, integrator=args.integrator, cost=args.cost, device=0, no_pairs=args.no_pairs)
    a = np.pi - 1e-3
    s = torch.tensor([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], device=eng.device, dtype=torch.float32)
    noise = torch.randn((eng.n_ind, args.K), device=eng.device)
    for _ in range(5):
        eng.mppi_step(s, noise, 1, 0.0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.mppi_step(s, noise, 1, 0.0)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    # the same solve inside a stream of back-to-back solves (20 per event pair): the device time per solve without the
    # host's launch latency, which an event pair around a single launch on an idle GPU includes
    tb = []
    for _ in range(max(3, args.iters // 10)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.mppi_step(s, noise, 1, 0.0)
        e1.record()
        e1.synchronize()
        tb.append(e0.elapsed_time(e1) / 20)
    mb = float(np.median(tb))
    print(f"MPPI solve K={args.K} T={args.T} {args.integrator} {args.cost}{' one-per-thread' if args.no_pairs else ''}: single launch {ms * 1e3:.1f} us median (events around one "
          f"launch on an idle GPU), {mb * 1e3:.1f} us per solve in a stream of 20, {args.K * args.T * 10 / mb * 1e3:.3e} state-steps/s")


if __name__ == "__main__":
    main()
