"""K sharded over 2 GPUs -- records exchanged inside the solve launch over peer memory (cps_mppi_set_peers), or by an NCCL
all-gather -- must give the same u_nom / u as one GPU solving all K rollouts.  Needs 2 GPUs (gpurun --gpus 2); skipped
otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, K, T, integ, exchange, q):
    import torch.distributed as dist
    from cartpolesimulation_b200.distributed import ShardedMPPI
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sm = ShardedMPPI(K, T, integrator=integ, cost="quadratic_boundary_grad_minimal", device=rank, exchange=exchange)
        assert sm.exchange == exchange
        rng = np.random.default_rng(3)
        eps = rng.standard_normal((sm.engine.n_ind, K)).astype(np.float32)
        a = np.pi - 1e-3
        s = torch.tensor([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], device=sm.device, dtype=torch.float32)
        us = []
        for step in range(3):
            full = torch.from_numpy(eps * (1.0 - 0.3 * step)).to(sm.device)
            u = sm.step(s, sm.noise_slice(full), 1, u_prev=0.05 * step)
            us.append(float(u.cpu()[0]))
        if exchange == "peer":
            assert sm.engine.peer_timeouts() == 0
        q.put((rank, us, sm.get_u_nom()))
    except Exception as ex:  # report instead of leaving the parent to time out
        q.put((rank, "error", repr(ex)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "allgather"])
@pytest.mark.parametrize("integ", ["ODE", "ODE_v0"])
def test_sharded_mppi_two_gpus_matches_one_gpu(integ, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from cartpolesimulation_b200.core import Engine
    K, T, world = 4001, 50, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, T, integ, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    assert all(r[1] != "error" for r in res), res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(res[0][2], res[1][2])   # bit-identical on both ranks
    assert res[0][1] == res[1][1]
    eng = Engine(K, T, integrator=integ, cost="quadratic_boundary_grad_minimal", device=0)
    rng = np.random.default_rng(3)
    eps = rng.standard_normal((eng.n_ind, K)).astype(np.float32)
    a = np.pi - 1e-3
    s = torch.tensor([a, 0.0, np.cos(a), np.sin(a), 0.0, 0.0], device=eng.device, dtype=torch.float32)
    for step in range(3):
        u = eng.mppi_step(s, torch.from_numpy(eps * (1.0 - 0.3 * step)).to(eng.device), 1, 0.05 * step)
        assert abs(float(u.cpu()[0]) - res[0][1][step]) < 2e-6
    np.testing.assert_allclose(eng.get_u_nom(), res[0][2], rtol=0, atol=2e-6)
