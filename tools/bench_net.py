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

NORM = {"Q": (1.0, -1.0), "angle": (np.pi, -np.pi), "angleD": (18.38, -18.38), "angle_cos": (1.0, -1.0),
        "angle_sin": (1.0, -1.0), "position": (0.198, -0.198), "positionD": (1.125, -1.125)}
INPUTS = ["Q", "angleD", "angle_cos", "angle_sin", "position", "positionD"]
OUTPUTS = INPUTS[1:]


def synthetic_spec(hsz=(64, 64), net_type="GRU", seed=0):
    from cartpolesimulation_b200.neural import build_net_spec
    rng = np.random.default_rng(seed)
    parts, n_in = [], len(INPUTS)
    G = 3 if net_type == "GRU" else 1
    for H in hsz:
        k = 1.0 / np.sqrt(H)
        parts.append(rng.uniform(-k, k, (G * H, n_in)))
        if net_type == "GRU":
            parts += [rng.uniform(-k, k, (3 * H, H)), rng.uniform(-k, k, 3 * H), rng.uniform(-k, k, 3 * H)]
        else:
            parts.append(rng.uniform(-k, k, H))
        n_in = H
    k = 1.0 / np.sqrt(n_in)
    parts += [rng.uniform(-k, k, (len(OUTPUTS), n_in)), rng.uniform(-k, k, len(OUTPUTS))]
    w = np.concatenate([p.reshape(-1) for p in parts]).astype(np.float32)
    cols = list(NORM)
    table = np.array([[0.0] * len(cols), [1.0] * len(cols), [NORM[c][0] for c in cols], [NORM[c][1] for c in cols]])
    return build_net_spec(net_type, INPUTS, OUTPUTS, list(hsz), w, (cols, table))


def flops_per_step(hsz, n_in=6, n_out=5, G=3):
    f, i = 0, n_in
    for H in hsz:
        f += 2 * G * H * (i + (H if G == 3 else 0))
        i = H
    return f + 2 * n_out * i


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=2000)
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    from cartpolesimulation_b200.core import Engine
    spec = synthetic_spec()
    K, T = args.K, args.T
    eng = Engine(K, T, integrator="neural", cost="quadratic_boundary_grad_minimal", device=0)
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
    fl = flops_per_step((64, 64)) * K * T
    print(f"neural MPPI solve K={K} T={T} GRU 2x64: kernel {ms * 1e3:.1f} us median, {K * T / ms * 1e3:.3e} net-steps/s, "
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
