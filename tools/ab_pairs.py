import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from bench import _event_times
from cartpolesimulation_b200.core import Engine
a = np.pi - 1e-3
s = torch.tensor([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], device="cuda", dtype=torch.float32)
for K in (16384, 32768, 65536):
    for T in (50, 100):
        row = []
        for no_pairs in (True, False):
            eng = Engine(K, T, integrator="ODE", cost="quadratic_boundary_grad_minimal", device=0, no_pairs=no_pairs)
            noise = torch.randn((eng.n_ind, K), device="cuda")
            Q = torch.empty((T, K), device="cuda").uniform_(-1, 1)
            m = float(np.median(_event_times(lambda: eng.mppi_step(s, noise, 1, 0.0), 30)))
            p = float(np.median(_event_times(lambda: eng.plan_cost(s, Q, 1, 0.0), 30)))
            row.append((m * 1e3, p * 1e3))
            eng.close()
        print(f"K={K} T={T}: mppi one/two per thread {row[0][0]:.1f} / {row[1][0]:.1f} us; plan_cost {row[0][1]:.1f} / {row[1][1]:.1f} us")
