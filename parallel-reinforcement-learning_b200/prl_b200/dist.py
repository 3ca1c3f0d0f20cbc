"""Multi-GPU plumbing: one process per GPU, envs sharded across ranks, weights replicated.

The reference has no distributed layer at all; the data-parallel design is section 8(e):
rank g owns envs [g*E/G, (g+1)*E/G), their fp64 state and their rollout buffer, and runs rollout + old-policy
evaluation + GAE locally with no communication.  The exchanges are
  (1) one allreduce(sum) of the flat float32 gradient per minibatch (36-53 KB; latency-bound over NVLink/NVSwitch),
  (2) one 4-double allreduce per learn() for the global advantage mean / std (PPO.py:199),
  (3) one allgather of the local row counts per learn() so every rank derives the same minibatch schedule.
torch.distributed is the transport (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as td

_ACTIVE = None


class Comm:
    def __init__(self, group=None):
        self.group = group
        self.rank = td.get_rank(group)
        self.world_size = td.get_world_size(group)

    def allreduce_(self, tensor: torch.Tensor) -> torch.Tensor:
        td.all_reduce(tensor, op=td.ReduceOp.SUM, group=self.group)
        return tensor

    def allreduce_max_(self, tensor: torch.Tensor) -> torch.Tensor:
        td.all_reduce(tensor, op=td.ReduceOp.MAX, group=self.group)
        return tensor

    def allgather_int(self, value: int) -> list[int]:
        out = [None] * self.world_size
        td.all_gather_object(out, int(value), group=self.group)
        return [int(v) for v in out]

    def barrier(self):
        td.barrier(group=self.group)


def shard_bounds(num_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block of env indices owned by `rank` (the remainder goes to the lowest ranks)."""
    base, rem = divmod(int(num_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def minibatch_schedule(n_all: list[int], mini_batch_size: int):
    """Global minibatch k = union of every rank's k-th local chunk of ceil(mb / world) rows.
    Returns (mb_local, n_minibatches, global row count of each minibatch)."""
    world = len(n_all)
    mb_local = -(-int(mini_batch_size) // world)
    n_mb = max(-(-n // mb_local) for n in n_all)
    counts = [sum(min(max(n - k * mb_local, 0), mb_local) for n in n_all) for k in range(n_mb)]
    return mb_local, n_mb, counts


def init_from_env(backend: str | None = None) -> Comm | None:
    """torchrun entry: reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*; returns None for a single process."""
    global _ACTIVE
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        _ACTIVE = None
        return None
    if not td.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group(backend=backend)
    _ACTIVE = Comm()
    return _ACTIVE


def activate(comm: Comm | None):
    global _ACTIVE
    _ACTIVE = comm


def active() -> Comm | None:
    return _ACTIVE
