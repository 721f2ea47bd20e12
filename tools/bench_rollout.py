#!/usr/bin/env python
"""Kernel timing of the open-loop rollout (cps_rollout, time-major, trajectory materialised): packed-pair kernel vs
one cartpole per thread -- the ncu target for rollout_pair_kernel / rollout_kernel."""
import argparse
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1 << 20)
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--integrator", default="ODE_v0")
    ap.add_argument("--no-pairs", action="store_true")
    ap.add_argument("--no-traj", action="store_true")
    ap.add_argument("--no-pairs-skip", action="store_true", help="packed-pair kernel only")
    args = ap.parse_args()
    import bench
    from cartpolesimulation_b200 import _lib as L
    from cartpolesimulation_b200.core import Engine
    s0_np, Q_np = bench.make_inputs(args.B, args.T, 1234)
    s0, Q = torch.from_numpy(s0_np).cuda(), torch.from_numpy(Q_np).cuda()
    traj = None if args.no_traj else torch.empty((args.T + 1, 6, args.B), device="cuda")
    fin = torch.empty((args.B, 6), device="cuda") if args.no_traj else None
    for no_pairs in ([True] if args.no_pairs else [False] if args.no_pairs_skip else [False, True]):
        eng = Engine(args.B, args.T, integrator=args.integrator, cost=None, device=0, no_pairs=no_pairs)
        run = lambda: eng.rollout(s0, Q, q_layout=L.TIME_MAJOR, traj_layout=L.TIME_MAJOR, want_traj=not args.no_traj,
                                  want_final=args.no_traj, traj_out=traj, final_out=fin)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(f"rollout {args.integrator} B={args.B} T={args.T} traj={not args.no_traj} "
              f"{'one-per-thread' if no_pairs else 'packed pairs  '}: {ms * 1e3:.1f} us median, "
              f"{args.B * args.T * 10 / ms * 1e3:.3e} state-steps/s")
        eng.close()


if __name__ == "__main__":
    main()
