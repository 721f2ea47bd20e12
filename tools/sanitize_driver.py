#!/usr/bin/env python
"""Small invocations of the kernels with cross-thread / cross-block hand-offs, meant to run under compute-sanitizer
(tools/sanitize.sh): the ticket + last-block merge of mppi_kernel / mppi_pair_kernel, the cooperative cem_select_kernel
(grid.sync between radix passes), the mbarrier / tensor-memory pipeline of net_tc_kernel, net_kernel, fleet_kernel and the
open-loop rollout kernels.  Sizes are tiny: the sanitizer slows kernels down by two to three orders of magnitude."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from cartpolesimulation_b200 import _lib as L          # noqa: E402
from cartpolesimulation_b200.core import Engine         # noqa: E402
from cartpolesimulation_b200.neural import synthetic_net_spec   # noqa: E402


def state(dev):
    a = np.pi - 1e-3
    return torch.tensor([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], device=dev, dtype=torch.float32)


def main(which):
    done = []
    if "mppi" in which:
        for K, T, cost in ((2000, 10, "quadratic_boundary_grad_minimal"), (4096, 33, "quadratic_boundary"), (65536, 4, "quadratic_boundary")):
            eng = Engine(K, T, integrator="ODE", cost=cost, device=0)
            noise = torch.randn((eng.n_ind, K), device=eng.device)
            for _ in range(2):
                u = eng.mppi_step(state(eng.device), noise, L.TIME_MAJOR, 0.0)
            torch.cuda.synchronize()
            assert np.isfinite(float(u.cpu()[0]))
            eng.close()
            done.append(f"mppi K={K} T={T} {cost}")
    if "cem" in which:
        for K in (2000, 16384):   # > 8192 plans: selection by the whole grid (cooperative launch)
            T = 8
            eng = Engine(K, T, integrator="ODE", cost="quadratic_boundary_grad_minimal", device=0)
            eng.cem_configure(max(1, K // 10), 0.5, 0.01)
            eps = torch.randn((2, T, K), device=eng.device)
            eng.cem_step(state(eng.device), eps, L.TIME_MAJOR, 0.0)
            torch.cuda.synchronize()
            eng.close()
            done.append(f"cem K={K}")
    if "net" in which:
        for kern, K, hid in (("tensor", 200, 64), ("tensor", 4800, 64), ("tensor", 9600, 64), ("fp32", 200, 64),
                             ("tensor", 200, 32), ("tensor", 9600, 32), ("tensor", 200, 48)):   # 32 / 64 / 128 live rollouts per CTA; narrow / padded nets
            spec = synthetic_net_spec((hid, hid), "GRU", seed=0)
            T = 4
            eng = Engine(K, T, integrator="neural", cost="quadratic_boundary_grad_minimal", device=0, net_kernel=kern)
            eng.net_load(spec)
            noise = torch.randn((eng.n_ind, K), device=eng.device)
            for _ in range(2):
                u = eng.mppi_step(state(eng.device), noise, L.TIME_MAJOR, 0.0)
            torch.cuda.synchronize()
            assert np.isfinite(float(u.cpu()[0])) and eng.net_last_kernel() == kern
            eng.close()
            done.append(f"net {kern} K={K} H={hid}")
    if "rollout" in which:
        for B in (4096, 1 << 18):
            T = 4
            eng = Engine(B, T, integrator="ODE_v0", cost=None, device=0)
            s0 = state(eng.device).repeat(B, 1).contiguous()
            Q = torch.empty((T, B), device=eng.device).uniform_(-1, 1)
            traj = torch.empty((T + 1, 6, B), device=eng.device)
            eng.rollout(s0, Q, q_layout=L.TIME_MAJOR, traj_layout=L.TIME_MAJOR, traj_out=traj)
            torch.cuda.synchronize()
            done.append(f"rollout B={B} ({eng.rollout_last_kernel()})")
            eng.close()
    if "fleet" in which:
        from cartpolesimulation_b200.fleet import DataGenConfig, Fleet, make_experiments
        E, K, T, periods = 4, 512, 8, 2
        cfg = DataGenConfig(length_of_experiment=2.0)
        s0, tp, te = make_experiments(E, periods, cfg, seed=0, experiment_offset=0)
        fl = Fleet(E, K, T, integrator="ODE", cost="quadratic_boundary_grad_minimal", noise="philox", seed=0, device=0)
        fl.reset(s0)
        fl.run(periods, torch.from_numpy(tp).to(fl.device), torch.from_numpy(te).to(fl.device), None, None)
        torch.cuda.synchronize()
        fl.close()
        done.append("fleet E=4")
    if "grad" in which:
        # fused Jacobian + reverse kernel (records in shared memory, block barrier), small and multi-plan blocks, and the
        # global-record path of large batches
        for K, T in ((16, 35), (2000, 12), (20000, 10)):
            eng = Engine(K, T, integrator="ODE", cost="quadratic_boundary_grad_minimal", device=0)
            Q = torch.empty((K, T), device=eng.device).uniform_(-0.5, 0.5)
            J, G = eng.plan_cost_grad(state(eng.device), Q)
            torch.cuda.synchronize()
            assert bool(torch.isfinite(G).all())
            if K <= 4096:   # get_action + bookkeeping kernel, resampling and plain solves
                keep = max(1, (3 * K) // 4)
                ages = torch.zeros(K, dtype=torch.int32, device=eng.device)
                eng.rpgd_reset()
                for fresh in (torch.zeros((K - keep, T), device=eng.device), None):
                    u_nom = eng.rpgd_finish(J, Q, fresh, keep, 1, ages)
                    assert np.isfinite(u_nom).all()
            eng.close()
            done.append(f"grad K={K} T={T}")
    print("sanitize_driver finished:", "; ".join(done))


if __name__ == "__main__":
    main(sys.argv[1:] or ["mppi", "cem", "net", "rollout", "fleet"])
