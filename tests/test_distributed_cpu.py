"""Host-side logic of the multi-GPU forms (cartpolesimulation_b200/distributed.py) on CPU: world_size-2 gloo.

The CUDA engine is replaced by a stand-in built on the CPU oracle that produces the same partial record
(min J, sum w, sum w*eps[.]) and applies the same merge rule as merge_and_finish (csrc/cps_device.cuh), so the test
covers the partition, the all-gather plumbing and the algebra of the sharded update: the sharded result must equal
the oracle's single-process solve over all K rollouts."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cartpolesimulation_b200.distributed import ShardedMPPI, shard_bounds


def test_shard_bounds_partition():
    for n, w in [(2000, 8), (65536, 8), (7, 3), (8192, 5), (3, 3)]:
        cuts = [shard_bounds(n, w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


class OracleEngine:
    """CPU stand-in for core.Engine in shard mode."""

    def __init__(self, K, T, integrator="ODE", cost="quadratic_boundary_grad_minimal", p=10, dt=0.02):
        from oracle import oracle as O
        self.O, self.K, self.T, self.p, self.dt = O, K, T, p, dt
        self.integrator, self.cost = integrator, cost
        self.device = torch.device("cpu")
        self.n_ind = O.num_inducing(T, p)
        self.mp = O.mppi_params(dt=dt)
        self.u_nom = np.zeros(T, np.float32)
        self.net_htot = 0

    def partial_size(self):
        return self.n_ind + 2

    def set_shard(self, buf):
        self.buf = buf

    def mppi_step(self, s, noise, layout, u_prev):
        eps = noise.numpy().T if layout == 1 else noise.numpy()
        out = self.O.mppi_step(self.integrator, self.cost, s.numpy(), self.u_nom, eps=eps, u_prev=u_prev, dt=self.dt, p=self.p)
        J = out["J"]
        m = J.min()
        w = np.exp(-(J - m) / self.mp[2]).astype(np.float32)
        rec = np.concatenate([[m, w.sum(dtype=np.float32)], (w[:, None] * eps).sum(0, dtype=np.float32)]).astype(np.float32)
        self.buf.copy_(torch.from_numpy(rec))

    def mppi_finalize(self, gathered):
        g = gathered.numpy().reshape(-1, self.n_ind + 2)
        m = g[:, 0].min()
        f = np.exp(-(g[:, 0] - m) / self.mp[2]).astype(np.float32)
        S = (g[:, 1] * f).sum(dtype=np.float32)
        E = (g[:, 2:] * f[:, None]).sum(0, dtype=np.float32)
        W = self.O.interp_matrix(self.T, self.p)                 # [n_ind, T]
        delta = self.mp[4] * (E @ W) / S
        shifted = np.concatenate([self.u_nom[1:], self.u_nom[-1:]])
        self.u_nom = np.clip(shifted + delta, self.mp[5], self.mp[6]).astype(np.float32)
        return torch.from_numpy(self.u_nom[:1].copy())

    def get_u_nom(self):
        return self.u_nom.copy()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, K, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)   # same on every rank: replicated inputs
        sm = ShardedMPPI(K, T, engine_factory=lambda k: OracleEngine(k, T))
        eps = rng.standard_normal((K, sm.engine.n_ind)).astype(np.float32)
        a = np.pi - 1e-3
        s = torch.tensor([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], dtype=torch.float32)
        us = []
        for step in range(2):
            noise_full = torch.from_numpy(np.ascontiguousarray(eps.T)) * (1.0 if step == 0 else -0.5)
            u = sm.step(s, sm.noise_slice(noise_full, time_major=True), 1, u_prev=0.1 * step)
            us.append(float(u[0]))
        q.put((rank, sm.lo, sm.hi, us, sm.get_u_nom()))
    finally:
        dist.destroy_process_group()


def test_sharded_mppi_world2_gloo_matches_single_process_oracle():
    from oracle import oracle as O
    K, T, world = 301, 35, 2   # ragged split: 151 + 150
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 151), (151, 301)]
    np.testing.assert_array_equal(res[0][4], res[1][4])          # identical u_nom on both ranks
    assert res[0][3] == res[1][3]
    # single-process reference over all K rollouts
    rng = np.random.default_rng(11)
    eps = rng.standard_normal((K, O.num_inducing(T, 10))).astype(np.float32)
    a = np.pi - 1e-3
    s = np.array([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], np.float32)
    u_nom = np.zeros(T, np.float32)
    for step in range(2):
        out = O.mppi_step("ODE", "quadratic_boundary_grad_minimal", s, u_nom, eps=eps * (1.0 if step == 0 else -0.5),
                          u_prev=0.1 * step)
        u_nom = out["u_nom"]
        assert abs(res[0][3][step] - float(out["u"])) < 2e-6
    np.testing.assert_allclose(res[0][4], u_nom, rtol=0, atol=2e-6)
