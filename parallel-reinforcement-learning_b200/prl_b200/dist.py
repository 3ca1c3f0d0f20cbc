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

    def broadcast_(self, tensor: torch.Tensor, src: int = 0) -> torch.Tensor:
        td.broadcast(tensor, src=src, group=self.group)
        return tensor

    def allgather_obj(self, value) -> list:
        out = [None] * self.world_size
        td.all_gather_object(out, value, group=self.group)
        return out

    def allgather_int(self, value: int) -> list[int]:
        """One integer per rank.  A tensor all_gather (one small collective) rather than all_gather_object, which pickles and
        runs two collectives with host synchronisations in between - this sits on the critical path of every sharded learn()."""
        cuda = td.get_backend(self.group) == "nccl"
        mine = torch.tensor([int(value)], dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()) if cuda else "cpu")
        out = torch.empty(self.world_size, dtype=torch.int64, device=mine.device)
        td.all_gather_into_tensor(out, mine, group=self.group)
        return [int(v) for v in out.tolist()]

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
        # PRL_SAME_DEVICE_PEERS=1 (tests only): every rank runs on GPU 0 - two processes time-slicing one device, gloo for the
        # collectives (NCCL refuses two ranks on one GPU), CUDA IPC for the exchange buffers.  It lets a 1-GPU box exercise the
        # sharded update end to end; each cross-rank wait then costs a context time slice, so it is no way to run anything else.
        same = os.environ.get("PRL_SAME_DEVICE_PEERS") == "1"
        if backend is None:
            backend = os.environ.get("PRL_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() and not same else "gloo")
        if torch.cuda.is_available():
            torch.cuda.set_device(0 if same else int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group(backend=backend)
    _ACTIVE = Comm()
    return _ACTIVE


def activate(comm: Comm | None):
    global _ACTIVE
    _ACTIVE = comm


def active() -> Comm | None:
    return _ACTIVE


RANK_STRIDE = 0x9E3779B97F4A7C15   # odd 64-bit constant: rank r's random streams are keyed seed + r * RANK_STRIDE


def rank_seed(seed: int, rank: int) -> int:
    """The seed of rank `rank`'s Philox streams (action sampling, env resets) derived from the job's seed: identically seeded
    ranks must not draw identical noise for their shards."""
    return (int(seed) + int(rank) * RANK_STRIDE) & ((1 << 62) - 1)


def peer_access_possible(comm: Comm) -> bool:
    """True when the in-kernel gradient exchange can be used: every rank of the group runs on this host, every pair of
    their GPUs has peer access, and they are distinct devices - agreed by all ranks (one allgather), so that either
    every rank takes the peer-memory path or every rank falls back to the NCCL allreduce."""
    import socket

    dev = torch.cuda.current_device()
    info = comm.allgather_obj((socket.gethostname(), dev))
    same = os.environ.get("PRL_SAME_DEVICE_PEERS") == "1"   # (tests only: see init_from_env)
    ok = len({h for h, _ in info}) == 1 and (same or len({d for _, d in info}) == len(info))
    if ok:
        ok = all(d == dev or torch.cuda.can_device_access_peer(dev, d) for _, d in info)
    return all(comm.allgather_obj(bool(ok)))


class PeerExchange:
    """Per-rank gradient exchange buffers shared across the GPUs of one node through CUDA IPC (NVLink / NVSwitch peer
    memory): what prl_ppo_step_tc_p2p sums over instead of calling an NCCL allreduce.  `table` is a device int64 tensor
    holding the `world_size` buffer addresses as seen from this process (own buffer at [rank]).  Only construct it when
    `peer_access_possible(comm)`; `close()` (collective) unmaps the peers' buffers and frees the own one."""

    def __init__(self, comm: Comm, is_continuous: bool, observ_dim: int, action_dim: int):
        import ctypes as C

        from . import _lib

        self.comm = comm
        self.rank, self.world_size = comm.rank, comm.world_size
        nbytes = int(_lib.fn("prl_p2p_exchange_bytes")(int(is_continuous), observ_dim, action_dim, self.world_size))
        own = C.c_void_p()
        _lib.call("prl_p2p_alloc", nbytes, C.byref(own))
        handle = C.create_string_buffer(64)
        _lib.call("prl_p2p_get_handle", own, handle)
        handles = [None] * self.world_size
        td.all_gather_object(handles, bytes(handle.raw), group=comm.group)
        self._own, self._opened, ptrs = own, [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(own.value)
            else:
                p = C.c_void_p()
                _lib.call("prl_p2p_open_handle", C.create_string_buffer(h, 64), C.byref(p))
                self._opened.append(p)
                ptrs.append(p.value)
        self.table = torch.tensor(ptrs, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
        comm.barrier()   # every rank has mapped every buffer before anybody signals through them

    def close(self):
        """Collective: after a barrier (nobody is still signalling through the buffers) close the peers' mappings and free
        the own buffer.  Idempotent."""
        from . import _lib

        if self._own is None:
            return
        torch.cuda.synchronize()
        try:
            self.comm.barrier()
        except Exception:   # the process group is already gone (interpreter shutdown): still release the local resources
            pass
        for p in self._opened:
            _lib.call("prl_p2p_close_handle", p)
        self._opened = []
        _lib.call("prl_p2p_free", self._own)
        self._own, self.table = None, None

    def __del__(self):
        # no collective in a finaliser: only release what belongs to this process if close() was never called
        try:
            from . import _lib

            if getattr(self, "_own", None) is not None:
                for p in self._opened:
                    _lib.fn("prl_p2p_close_handle")(p)
                _lib.fn("prl_p2p_free")(self._own)
                self._own = None
        except Exception:
            pass
