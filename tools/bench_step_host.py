#!/usr/bin/env python
"""Host-to-host latency of one MPPI solve through cps_mppi_step_host (numpy state in, float control out), K = 2000, T = 50."""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cartpolesimulation_b200.core import Engine  # noqa: E402

for integ in ("ODE", "ODE_v0"):
    eng = Engine(2000, 50, integrator=integ, cost="quadratic_boundary_grad_minimal", device=0)
    a = np.pi - 1e-3
    s = np.array([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], dtype=np.float32)
    noise = torch.randn((eng.n_ind, 2000), device="cuda")
    for _ in range(50):
        eng.mppi_step_host(s, noise, 1, 0.0)
    lat = []
    for _ in range(2000):
        t0 = time.perf_counter()
        eng.mppi_step_host(s, noise, 1, 0.0)
        lat.append((time.perf_counter() - t0) * 1e6)
    print(f"{integ}: cps_mppi_step_host median {np.median(lat):.1f} us, p99 {np.percentile(lat, 99):.1f} us")
