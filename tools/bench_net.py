#!/usr/bin/env python
"""Timing of the neural-predictor MPPI solve (BASELINE.json configs[2]: GRU 2x64, K=2000, T=50) and of large-K
network rollouts.  Synthetic weights (torch GRUCell default init U(-1/sqrt(H), 1/sqrt(H)), seeded)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from cartpolesimulation_b200.neural import net_flops_per_step, synthetic_net_spec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=2000)
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--kernel", default=None, choices=[None, "tensor", "fp32"])
    ap.add_argument("--hidden", type=int, default=64, help="units per GRU layer (64: tensor-core kernel; else FP32 kernel)")
    args = ap.parse_args()
    from cartpolesimulation_b200.core import Engine
    spec = synthetic_net_spec((args.hidden, args.hidden))
    K, T = args.K, args.T
    eng = Engine(K, T, integrator="neural", cost="quadratic_boundary_grad_minimal", device=0, net_kernel=args.kernel)
    eng.net_load(spec)
    dev = eng.device
    a = np.pi - 1e-3
    s = torch.tensor([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], device=dev, dtype=torch.float32)
    noise = torch.randn((eng.n_ind, K), device=dev)
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
    fl = net_flops_per_step(spec) * K * T
    print(f"neural MPPI solve K={K} T={T} GRU 2x{args.hidden}: kernel {ms * 1e3:.1f} us median, {K * T / ms * 1e3:.3e} net-steps/s, "
          f"{fl / ms / 1e9:.2f} TFLOP/s fp32 ({fl / 1e9:.2f} GFLOP per solve)")
    # host-to-host latency through step_host
    s_np = s.cpu().numpy()
    for _ in range(20):
        eng.mppi_step_host(s_np, noise, 1, 0.0)
    lat = []
    for _ in range(200):
        t0 = time.perf_counter()
        eng.mppi_step_host(s_np, noise, 1, 0.0)
        lat.append(time.perf_counter() - t0)
    print(f"  host-to-host: median {np.median(lat) * 1e3:.3f} ms, p99 {np.percentile(lat, 99) * 1e3:.3f} ms")


if __name__ == "__main__":
    main()
