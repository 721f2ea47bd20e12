"""Multi-GPU forms of the path, one process per GPU over torch.distributed.

  ShardedMPPI        ONE MPPI solve with its K rollouts partitioned over the ranks.  Every rank rolls out its
                     contiguous slice of rollouts (its own slice of the noise draws; s, u_nom and the scalars are
                     replicated) and reduces it to a partial record (min J, sum w, sum w*eps[.]) = n_ind + 2 floats.
                     The only exchange is that record (32 B per rank at n_ind = 6), merged on every rank with the
                     online-softmax rule, so that all ranks hold the identical u_nom / u.  Two transports:
                       exchange="peer"       inside the solve launch: the last block pushes the record into every
                                             rank's symmetric-memory buffer over NVLink, waits for the others' and
                                             finishes the update itself (cps_mppi_set_peers) -- one launch per solve,
                                             no collective call;
                       exchange="allgather"  cps_mppi_set_shard + dist.all_gather_into_tensor + cps_mppi_finalize
                                             (three stream operations; any backend, also gloo on CPU in the tests).
  replica_slice      independent experiments / open-loop batches: plain partition, no collective at all.

The reference has no counterpart: it parallelises by running one process per experiment on a SLURM array
(others/EulerClusterScripts/ParallelDataGeneration.sh).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index: int) -> list | None:
    """Pin the calling process to the CPUs NVML reports as local to the GPU (its NUMA node), so that pinned host
    buffers allocated afterwards land in memory next to the GPU's PCIe root.  One process per GPU on a two-socket box
    otherwise sends half of the host<->device traffic across the socket interconnect.  Returns the CPU list, or None if
    NVML / the cpuset does not allow it (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous partition of n items over `world` ranks; the first n % world ranks get one more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    base, rem = divmod(int(n), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def replica_slice(n: int, group=None) -> tuple[int, int]:
    """This rank's share of n independent units (experiments, cartpoles of an open-loop batch)."""
    if dist.is_available() and dist.is_initialized():
        return shard_bounds(n, dist.get_world_size(group), dist.get_rank(group))
    return 0, int(n)


class ShardedMPPI:
    """One MPPI solve over K_total rollouts sharded across the ranks of `group`.

    engine_factory(K_local) must return an object with the Engine methods used here (mppi_step, set_shard,
    partial_size, mppi_finalize, n_ind, device, and for recurrent predictors net_update / net_htot); the default
    builds a cartpolesimulation_b200.core.Engine from **engine_kwargs.
    """

    def __init__(self, num_rollouts: int, horizon: int, group=None, engine_factory=None, exchange: str = "auto",
                 **engine_kwargs):
        if exchange not in ("auto", "peer", "allgather"):
            raise ValueError("exchange must be 'auto', 'peer' or 'allgather'")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.K_total, self.T = int(num_rollouts), int(horizon)
        if self.K_total < self.world:
            raise ValueError(f"cannot shard {self.K_total} rollouts over {self.world} ranks")
        self.lo, self.hi = shard_bounds(self.K_total, self.world, self.rank)
        self.K_local = self.hi - self.lo
        if engine_factory is None:
            from .core import Engine
            engine_factory = lambda k: Engine(num_rollouts=k, horizon=self.T, **engine_kwargs)
        self.engine = engine_factory(self.K_local)
        self.device = self.engine.device
        self.rec = int(self.engine.partial_size())
        self.recurrent = bool(getattr(self.engine, "net_htot", 0))
        self.exchange = "allgather"
        self._symm = None
        if exchange != "allgather" and self.world > 1:
            try:
                self._setup_peers()
                self.exchange = "peer"
            except Exception:
                if exchange == "peer":
                    raise
        if self.exchange == "allgather":
            self._partial = torch.zeros(self.rec, device=self.device, dtype=torch.float32)
            self._gathered = torch.zeros(self.world * self.rec, device=self.device, dtype=torch.float32)
            self.engine.set_shard(self._partial)

    def _setup_peers(self):
        """Symmetric-memory exchange buffers (torch.distributed._symmetric_memory: CUDA virtual-memory handles exchanged
        over the process group's store) mapped into every rank; their device addresses go to cps_mppi_set_peers."""
        if not hasattr(self.engine, "set_peers") or self.device.type != "cuda":
            raise RuntimeError("peer exchange needs the CUDA engine")
        import torch.distributed._symmetric_memory as symm_mem
        n = int(self.engine.peer_buffer_floats(self.world))
        buf = symm_mem.empty(n, dtype=torch.float32, device=self.device)
        buf.zero_()
        group = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(buf, group)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)   # every rank's buffer is zeroed before anybody's first solve pushes into it
        self.engine.set_peers(self.world, self.rank, [int(p) for p in hdl.buffer_ptrs])
        self._symm = (buf, hdl)

    def noise_slice(self, noise_full: torch.Tensor, time_major: bool = True) -> torch.Tensor:
        """This rank's rollouts of a full noise tensor ([n_ind, K_total] time-major, or [K_total, n_ind])."""
        part = noise_full[:, self.lo:self.hi] if time_major else noise_full[self.lo:self.hi]
        return part.contiguous()

    def step(self, s: torch.Tensor, noise_local: torch.Tensor, noise_layout: int = 1, u_prev: float = 0.0) -> torch.Tensor:
        """s: device tensor [6] (identical on all ranks); noise_local: this rank's draws.  Returns the device tensor
        [1] holding u -- identical on every rank.  No host synchronisation."""
        if self.exchange == "peer":   # one launch: local rollouts, record exchange over peer memory, update (+ GRU state)
            return self.engine.mppi_step(s, noise_local, noise_layout, u_prev)
        self.engine.mppi_step(s, noise_local, noise_layout, u_prev)   # -> self._partial (stream-ordered)
        if self.world > 1:
            dist.all_gather_into_tensor(self._gathered, self._partial, group=self.group)
        else:
            self._gathered.copy_(self._partial)
        u = self.engine.mppi_finalize(self._gathered)
        if self.recurrent:  # the fused kernel skips the hidden-state update when K is sharded (u is not known yet)
            self.engine.net_update(s, u)
        return u

    def get_u_nom(self) -> np.ndarray:
        return self.engine.get_u_nom()

    def reset(self, value: float = 0.0):
        self.engine.mppi_reset(value)
