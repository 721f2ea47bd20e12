#!/usr/bin/env python
"""A few gradient evaluations (cps_plan_cost_grad) at K = 2000, T = 50 for an ncu capture of plan_grad_fwd_kernel /
plan_grad_jacrev_kernel:  ncu --set full -k regex:plan_grad -s 4 -c 2 -o gpurun_out/r02_plan_grad python tools/profile_grad.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartpolesimulation_b200.core import Engine   # noqa: E402

K, T = 2000, 50
eng = Engine(K, T, integrator="ODE", cost="quadratic_boundary_grad_minimal", device=0)
a = np.pi - 1e-3
s = torch.tensor([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], device=eng.device, dtype=torch.float32)
Q = torch.zeros((K, T), device=eng.device).uniform_(-0.5, 0.5)
for _ in range(4):
    J, G = eng.plan_cost_grad(s, Q)
torch.cuda.synchronize()
print("finite:", bool(torch.isfinite(G).all()))
