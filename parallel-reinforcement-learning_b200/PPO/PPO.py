"""Drop-in `PPO` (reference: /root/reference/PPO/PPO.py:13-283) on the sm_100a kernels of libprl_b200.so.

Same constructor arguments, attributes and methods as the reference.  What changes is where the work happens:

  get_action   one fused launch: trunk + policy head + Categorical / tanh-Gaussian sampling (prl_policy_act)
  learn        stays on the device end to end: old-policy evaluation (1 launch over all N rows), RND reward mixing and
               predictor updates, float32 GAE in the reference's operation order (prl_gae), advantage normalisation
               (prl_adv_normalize), then k_epochs x ceil(N / mini_batch_size) strictly sequential steps of
               [fused forward+loss+backward -> flat gradient (prl_ppo_grad)] -> [NCCL allreduce when sharded] ->
               [clip_grad_norm_(2.0) + AdamW in one kernel (prl_adamw_step)], no host sync inside the loop.

There is no CPU fallback: constructing a PPO without the shared library or without a CUDA device raises.
"""
from __future__ import annotations

import numpy as np
import torch as t
from torch import nn
from tqdm import tqdm

from prl_b200 import dist as pdist
from prl_b200 import ops
from prl_b200._lib import require_cuda
from prl_b200.optim import FusedAdamW

from .ActorCritic import ActorCritic
from .Memory import Memory
from .RND import RND


class PPO:
    def __init__(
            self,
            is_continuous: bool,
            observ_dim: int,
            action_dim: int,
            action_scaling: float = None,
            lr: float = 0.001,
            k_epochs: int = 7,
            policy_clip: float = 0.2,
            GAE_lambda: float = 0.95,
            gamma: float = 0.995,
            batch_size: int = 1024,
            mini_batch_size: int = 64,
            use_RND: bool = False,
            beta: int = 0.001
    ):
        self.device = require_cuda()
        self.policy = ActorCritic(is_continuous, observ_dim, action_dim)
        self.policy_old = ActorCritic(is_continuous, observ_dim, action_dim)
        if use_RND:
            self.rnd = RND(in_features=observ_dim, out_features=observ_dim, beta=beta)

        self.memory = Memory()
        self.policy_old.flat.copy_(self.policy.flat)  # policy_old.load_state_dict(policy.state_dict())  PPO.py:44
        self.policy.train()
        self.policy_old.eval()

        self.loss_fn = nn.SmoothL1Loss()  # API attribute; the fused kernel evaluates SmoothL1(mean, beta=1) itself
        self.optimizer = FusedAdamW(self.policy.flat, lr=lr, weight_decay=0.01, max_norm=2.0)

        self.is_continuous = is_continuous
        self.action_scaling = action_scaling
        self.use_RND = use_RND
        self.beta = beta
        self.lr = lr
        self.policy_clip = policy_clip
        self.k_epochs = k_epochs
        self.GAE_lambda = GAE_lambda
        self.gamma = gamma
        self.batch_size = batch_size
        self.mini_batch_size = mini_batch_size
        self.observ_dim = observ_dim
        self.action_dim = action_dim

        # sampling: counter-based Philox keyed by (seed; row, call) - the seed comes from torch's global generator so
        # that torch.manual_seed(...) makes runs repeatable, like the reference's dist.sample()
        self._seed = int(t.randint(0, 2 ** 62, (1,)).item())
        self._calls = 0
        # progress / loss reporting: the reference syncs once per minibatch for the tqdm label (PPO.py:254-255)
        self.show_progress = True
        self.report_loss = False
        self.last_losses = None   # device [steps, 4] float64: sums of (policy term, SmoothL1 term, entropy), rows
        self._grad = t.zeros_like(self.policy.flat)
        self._ws = None
        self._rows = None
        # minibatch-gradient kernel: "tensor" = tcgen05 kernel (csrc/update_tc.cu; discrete policies in one fused launch per step,
        # continuous ones as a pre-pass + two passes of the same kernel), "fp32" = CUDA-core FMA kernel (csrc/update_ppo.cu; every
        # configuration).  Same math and tolerances, two members of one family.
        self.peer_exchange = True     # sharded tensor path: gradient exchange over peer memory inside the step kernel
        self._xch = None
        self.fused_optimizer = True   # tensor path, single process: one launch per optimiser step (prl_ppo_step_tc)
        self.graph_collectives = False  # also capture the per-minibatch gradient allreduce (NCCL) in the epoch graph
        self.use_cuda_graph = False   # capture each epoch of learn() in a CUDA graph (single process, >= 16 optimiser steps)
        # AsyncPPO's fused worker also records log-prob / value of every transition under the acting policy (and, without RND, the
        # GAE returns): learn() then skips its old-policy pass - same bits, one pass over the rows less
        self.fuse_evaluation = True
        self.update_path = "tensor" if ops.tc_supported(is_continuous, observ_dim, action_dim) else "fp32"
        self._ws_owner = None         # which update path's header lives in self._ws
        self._p2p_ok = None           # sharded: can the ranks exchange gradients over peer memory (agreed once, collectively)
        # sharded (one process per GPU, prl_b200.dist active): only gradients are exchanged afterwards, so every replica must
        # START identical - rank 0's networks, optimiser state and seed are broadcast once, here or at the first learn()
        self._replicas_synced = False
        if pdist.active() is not None:
            self.sync_replicas(pdist.active())

    # ------------------------------------------------------------------------------------------------ sharding
    def sync_replicas(self, comm) -> None:
        """Collective.  Make every rank's replica identical to rank 0's: policy, policy_old, RND target / predictor, AdamW
        moments and step clocks.  The sampling seed becomes rank_seed(rank 0's seed, rank): the same on-device Philox generator
        keyed differently per shard, so identically seeded processes do not draw identical action noise."""
        bufs = [self.policy.flat, self.policy_old.flat, self.optimizer.exp_avg, self.optimizer.exp_avg_sq, self.optimizer.step_dev]
        if self.use_RND:
            bufs += [self.rnd.target_flat, self.rnd.pred_flat, self.rnd.optimizer.exp_avg, self.rnd.optimizer.exp_avg_sq, self.rnd.optimizer.step_dev]
        for b in bufs:
            comm.broadcast_(b, src=0)
        meta = t.tensor([self._seed, self.optimizer.step_count, self.rnd.optimizer.step_count if self.use_RND else 0],
                        dtype=t.int64, device=self.device)
        comm.broadcast_(meta, src=0)
        seed0, self.optimizer.step_count, rnd_steps = (int(x) for x in meta.cpu())
        if self.use_RND:
            self.rnd.optimizer.step_count = rnd_steps
        self._seed = pdist.rank_seed(seed0, comm.rank)
        self._replicas_synced = True

    def close(self) -> None:
        """Release the peer-memory exchange buffers of a sharded run (collective; see prl_b200.dist.PeerExchange.close)."""
        if self._xch is not None:
            self._xch.close()
            self._xch = None

    def _eval_tag(self):
        """What the fused worker's by-products (log-prob, value, returns) were computed from; learn() uses them only if this still
        holds: the acting network's buffer and its in-place version counter (any load / copy / step bumps it), gamma, lambda."""
        f = self.policy_old.flat
        return (f.data_ptr(), f._version, float(self.gamma), float(self.GAE_lambda), bool(self.use_RND))

    # ------------------------------------------------------------------------------------------------ acting
    def _action_scale(self) -> float:
        if not self.is_continuous:
            return 1.0
        if self.action_scaling is None:
            raise TypeError("continuous PPO needs action_scaling (the reference multiplies tanh(action) by it)")
        return float(self.action_scaling)

    @t.no_grad()
    def get_action_device(self, states: t.Tensor, row_ids: t.Tensor | None = None, call_index: int | None = None) -> t.Tensor:
        """Actions for `states` [n, O] (CUDA float32) as a CUDA tensor: int64 [n] / float32 [n, A]."""
        if call_index is None:
            call_index = self._calls
            self._calls += 1
        return ops.policy_act(self.policy_old.flat, self.is_continuous, self.observ_dim, self.action_dim, self._action_scale(),
                              states, self._seed, call_index, row_ids=row_ids)

    @t.no_grad()
    def get_action(self, state: t.Tensor, _row_ids: t.Tensor | None = None, _call_index: int | None = None) -> np.ndarray:
        """PPO.py:82-96: host tensor of any float dtype in, host numpy out."""
        state = state.to(dtype=t.float32, device=self.device).reshape(-1, self.observ_dim).contiguous()
        return self.get_action_device(state, _row_ids, _call_index).cpu().numpy()

    def batch_packer(self, values, batch_size: int):
        """PPO.py:98-105: sequential chunks (what list(DataLoader(tensor, batch_size)) yields)."""
        if isinstance(values, t.Tensor):
            batch = list(t.split(values, batch_size))
        elif isinstance(values, list):
            batch = [list(t.split(value, batch_size)) for value in values]
        return batch

    # ------------------------------------------------------------------------------------------------ GAE
    def compute_gae(self, rewards: np.ndarray, dones: np.ndarray, state_values: np.ndarray, next_value: np.ndarray):
        """PPO.py:107-120 on the device (prl_gae: float32, the reference's operation order); returns a list."""
        dev = lambda x: t.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1))).to(self.device)  # noqa: E731
        nv = dev(np.asarray(next_value, dtype=np.float32).reshape(-1)[:1])
        out = ops.gae(dev(rewards), dev(dones), dev(state_values), self.gamma, self.GAE_lambda, next_value=nv)
        return list(out.cpu().numpy())

    # ------------------------------------------------------------------------------------------------ learn
    @t.no_grad()
    def learn(self):
        comm = pdist.active()
        if comm is None:
            if len(self.memory.states) < self.batch_size:   # PPO.py:123-124
                return
        else:
            # sharded: the decision has to be the same on every rank (learn() is collective) - batch_size counts the rows of
            # all shards together
            n_all = comm.allgather_int(len(self.memory.states))
            if sum(n_all) < self.batch_size:
                return
            if min(n_all) == 0:
                raise RuntimeError(f"sharded learn(): a rank holds no transitions (rows per rank: {n_all})")
            if not self._replicas_synced:
                self.sync_replicas(comm)
        O, A, cont = self.observ_dim, self.action_dim, self.is_continuous
        AW = A if cont else 1
        states, actions, rewards, dones = self.memory.device_view(O, AW, self.device)
        N = states.shape[0]
        mb = int(self.mini_batch_size)

        # per-row scratch, allocated once at the store's capacity and reused by every learn() (no allocator traffic)
        cap = max(self.memory._dev_cap, N)
        if self._rows is None or self._rows["cap"] < cap:
            f32 = lambda: t.empty(cap, dtype=t.float32, device=self.device)  # noqa: E731
            self._rows = dict(cap=cap, logp=f32(), values=f32(), returns=f32(), adv=f32(), rewards=f32(),
                              gae_ws=t.empty(int(ops._lib.fn("prl_gae_ws_bytes")(cap)), dtype=t.uint8, device=self.device),
                              ent=t.zeros(1, dtype=t.float64, device=self.device), stats=t.zeros(4, dtype=t.float64, device=self.device))
        R = self._rows

        # old-policy evaluation (PPO.py:134-154): taken from the fused worker when it recorded it for exactly these rows under the
        # policy_old that is still in place; otherwise row-independent, so one launch over all N rows
        pre = self.memory.evaluated(N, self._eval_tag()) if self.fuse_evaluation else None
        if pre is not None:
            old_logp, old_values, pre_returns = pre
        else:
            pre_returns = None
            old_logp, old_values, _ = ops.policy_evaluate(self.policy_old.flat, cont, O, A, states, actions, entropy_sum=R["ent"],
                                                          logp=R["logp"][:N], value=R["values"][:N])

        if self.use_RND:  # PPO.py:157-178: rewards + intrinsic, THEN one predictor pass over the same chunks
            rewards = self.rnd.intrinsic_reward_device(states, add_to=rewards, out=R["rewards"][:N])
            if comm is None:
                for i in range(0, N, mb):
                    self.rnd.update_pred_chunk(states[i:i + mb])
            else:   # global chunk k = union of every rank's k-th local chunk, as for the policy minibatches below
                mb_r, n_r, counts_r = pdist.minibatch_schedule(n_all, mb)
                for k in range(n_r):
                    self.rnd.update_pred_chunk(states[min(k * mb_r, N):min((k + 1) * mb_r, N)], comm=comm, global_rows=counts_r[k])
        self.memory.clear()  # PPO.py:184 (the device rows stay valid until the next transfer)

        # next_value = V(last stored state), PPO.py:188
        # (every episode of a worker() ends with done = 1, so the flat scan factorises over envs: the fused worker's column scan
        # over the time-major planes gives the same bits - tests/test_gpu_kernels.py::test_full_size_c2_properties)
        if pre_returns is not None:
            returns = pre_returns
        else:
            returns = ops.gae(rewards, dones, old_values, self.gamma, self.GAE_lambda, out=R["returns"][:N], ws=R["gae_ws"])
        # advantages = returns - values; (adv - mean) / (std + 1e-8) over ALL rows of ALL ranks (PPO.py:198-199)
        stats = R["stats"].zero_()
        ops.adv_normalize(returns, old_values, stats=stats, phase=1)
        if comm is not None:
            comm.allreduce_(stats)
        adv, _ = ops.adv_normalize(returns, old_values, stats=stats, phase=2, out=R["adv"][:N])

        # minibatch schedule: sequential chunks of the flat env-major buffer, same order every epoch (PPO.py:202-211).
        # Sharded: global minibatch k = union of every rank's k-th local chunk.
        if comm is not None:
            mb_local, n_mb, counts = pdist.minibatch_schedule(n_all, mb)
        else:
            mb_local, n_mb = mb, -(-N // mb)
            counts = [min(N - k * mb, mb) for k in range(n_mb)]

        use_tc = self.update_path == "tensor"
        tc_level = ops.tc_supported(cont, O, A)
        if use_tc and not tc_level:
            raise ValueError("update_path='tensor' supports policies with observ_dim <= 16 and action_dim <= 8")
        grad_fn = ops.ppo_grad_tc if use_tc else ops.ppo_grad
        need = (ops.update_tc_ws_floats if use_tc else ops.update_ws_floats)(cont, O, A, min(mb_local, N))
        if self._ws is None or self._ws.numel() < need:
            self._ws = t.zeros(need, dtype=t.float32, device=self.device)
        elif self._ws_owner != self.update_path:
            self._ws[:4].zero_()   # the tensor path keeps its status word / arrival counter / launch counter in the header; the fp32 path writes partials there
        self._ws_owner = self.update_path
        steps = self.k_epochs * n_mb

        fused = use_tc and tc_level == 1 and comm is None and self.fused_optimizer   # gradient + clip + AdamW in one cooperative launch
        # sharded: the same single launch, with the gradient exchange over NVLink peer memory inside it (no NCCL per step) -
        # when every rank sits on this host and all GPU pairs have peer access; otherwise the NCCL allreduce path below
        p2p = use_tc and tc_level == 1 and comm is not None and self.fused_optimizer and self.peer_exchange
        if p2p:
            if self._p2p_ok is None:
                self._p2p_ok = pdist.peer_access_possible(comm)
            p2p = self._p2p_ok
        if p2p and self._xch is None:
            self._xch = pdist.PeerExchange(comm, cont, O, A)
        # a failed step (an in-kernel wait that timed out) must not leave a half-applied update behind: keep what is needed to roll back
        snap = (self.optimizer.exp_avg.clone(), self.optimizer.exp_avg_sq.clone(), self.optimizer.step_dev.clone(), self.optimizer.step_count) if use_tc else None

        def minibatch_step(k, loss_slot):
            lo, hi = min(k * mb_local, N), min((k + 1) * mb_local, N)
            if p2p:
                rows = (states[lo:hi], actions[lo:hi], old_logp[lo:hi], adv[lo:hi], returns[lo:hi]) if hi > lo else (None,) * 5
                ops.ppo_step_tc_p2p(self.policy.flat, cont, O, A, *rows, self.policy_clip, 1.0 / counts[k], self._grad, loss_slot,
                                    self.optimizer, self._ws, self._xch)
                self.optimizer.step_count += 1
                return hi - lo
            if fused:
                ops.ppo_step_tc(self.policy.flat, cont, O, A, states[lo:hi], actions[lo:hi], old_logp[lo:hi], adv[lo:hi],
                                returns[lo:hi], self.policy_clip, 1.0 / counts[k], self._grad, loss_slot, self.optimizer, self._ws)
                self.optimizer.step_count += 1
                return hi - lo
            if hi > lo:
                grad_fn(self.policy.flat, cont, O, A, states[lo:hi], actions[lo:hi], old_logp[lo:hi], adv[lo:hi],
                        returns[lo:hi], self.policy_clip, 1.0 / counts[k], self._grad, loss_slot, self._ws)
            else:
                self._grad.zero_()  # this rank has no rows in minibatch k; it still joins the allreduce
            if comm is not None:
                comm.allreduce_(self._grad)
            self.optimizer.step(self._grad)
            return hi - lo

        # (sharded runs capture the NCCL allreduce inside the graph as well when graph_collectives is set)
        ev = None
        if getattr(self, "update_events", None) is not None:   # (bench.py: CUDA events around the optimisation loop, on the launching stream)
            ev = (t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True))
            ev[0].record()
        use_graph = self.use_cuda_graph and (comm is None or self.graph_collectives or p2p) and not self.report_loss and steps >= 16 and n_mb >= 2
        pbar = tqdm(total=N * self.k_epochs, leave=False) if self.show_progress else None
        if use_graph:
            # one epoch = n_mb launch sets with fixed pointers: captured once, replayed k_epochs times (no host work between the
            # launches); last_losses[k] then ACCUMULATES minibatch k's loss sums over the epochs.  The capture covers the first
            # n_mb - 1 minibatches - full ones, the same launches whatever N is - and is KEPT across learn() calls (key: every
            # pointer and size that went into it); the last minibatch, whose row count follows N, is launched eagerly after each
            # replay.  In steady state learn() therefore captures nothing: the capture cost (~4 ms of host time for 128 launches)
            # would otherwise sit exposed between the rollout and the first optimiser step.
            full = n_mb - 1
            spans = tuple((min(k * mb_local, N), min((k + 1) * mb_local, N), counts[k]) for k in range(full))
            o = self.optimizer
            key = (tuple(x.data_ptr() for x in (states, actions, old_logp, adv, returns, self.policy.flat, self._ws, self._grad, o.exp_avg,
                                                o.exp_avg_sq, o.step_dev, o.grad_norm)),
                   spans, bool(p2p), bool(fused), float(self.policy_clip), o.lr, o.weight_decay, o.max_norm, id(self._xch), comm is not None)
            cache = self.__dict__.setdefault("_graph_cache", {})
            entry = cache.get(key)
            first_step = self.optimizer.step_count
            if entry is None:
                losses = t.zeros(n_mb, 4, dtype=t.float64, device=self.device)
                self.last_losses = losses
                side = self._side_stream = getattr(self, "_side_stream", None) or t.cuda.Stream()
                side.wait_stream(t.cuda.current_stream())
                graph = t.cuda.CUDAGraph()
                counts0 = dict(ops._lib.CALL_COUNTS)
                with t.cuda.stream(side):
                    graph.capture_begin()   # (torch.cuda.graph() would also synchronise, collect garbage and empty the allocator cache)
                    try:
                        for k in range(full):
                            minibatch_step(k, losses[k])
                    finally:
                        graph.capture_end()
                t.cuda.current_stream().wait_stream(side)
                captured = {k: v - counts0.get(k, 0) for k, v in ops._lib.CALL_COUNTS.items() if v != counts0.get(k, 0)}
                for k, v in captured.items():
                    ops._lib.CALL_COUNTS[k] -= v          # the capture itself launched nothing
                while len(cache) >= 4:
                    cache.pop(next(iter(cache)))
                entry = cache[key] = (graph, losses, captured)
            graph, losses, captured = entry
            self.last_losses = losses.zero_()
            self.optimizer.step_count = first_step   # (minibatch_step counted during a capture)
            for _ in range(self.k_epochs):
                if full:
                    graph.replay()   # AdamW's step number is device-resident, so every replay advances it
                    for k, v in captured.items():
                        ops._lib.CALL_COUNTS[k] += v
                minibatch_step(n_mb - 1, losses[n_mb - 1])
                if pbar is not None:
                    pbar.update(N)
            self.optimizer.step_count = first_step + steps
            self._graphs = graph  # keep alive until the replays have run
        else:
            self.last_losses = t.zeros(steps, 4, dtype=t.float64, device=self.device)
            step = 0
            for _ in range(self.k_epochs):
                for k in range(n_mb):
                    n_rows = minibatch_step(k, self.last_losses[step])
                    if pbar is not None:
                        pbar.update(n_rows)
                        if self.report_loss:
                            l = self.last_losses[step].cpu().numpy()
                            pbar.set_description(f"Loss: {(l[0] + 0.5 * l[1] - 0.01 * l[2]) / max(l[3], 1.0): .6f}")
                    step += 1
        if pbar is not None:
            pbar.close()
        if ev is not None:
            ev[1].record()
            self.update_events.append((ev[0], ev[1], steps))
        status = ops.ppo_grad_tc_status(self._ws) if use_tc else 0
        self.memory.verify_transfers()   # (the status read above synchronised the stream: the transfer totals are final)
        if comm is not None and use_tc:
            # every rank learns about a failure anywhere, so that all of them raise together instead of hanging in the next collective
            st = t.tensor([status], dtype=t.int32, device=self.device)
            comm.allreduce_max_(st)
            status = int(st.item())
        if status != 0:
            # roll back to the state before this learn(): weights from policy_old (PPO.py:258-260 has not run yet), moments and clocks
            self.policy.flat.copy_(self.policy_old.flat)
            self.optimizer.exp_avg.copy_(snap[0]); self.optimizer.exp_avg_sq.copy_(snap[1]); self.optimizer.step_dev.copy_(snap[2])
            self.optimizer.step_count = snap[3]
            self._ws[:4].zero_()
            raise RuntimeError("tensor-core update failed on some rank (weights and optimiser state rolled back to the start of learn()): " +
                               {1: "an MMA phase never completed (mbarrier timeout)", 2: "a wait on another CTA timed out",
                                3: "a peer rank never signalled its gradient"}.get(status, f"status {status}"))
        self.policy_old.flat.copy_(self.policy.flat)  # PPO.py:258-260

    # ------------------------------------------------------------------------------------------------ checkpoints
    def load_weights(self, path: str):
        """PPO.py:262-277: same files and state_dict keys as the reference; a missing file is ignored."""
        try:
            self.policy.load_state_dict(t.load(path + '/Policy_weights.pth', weights_only=True, map_location=self.device))
            self.policy_old.flat.copy_(self.policy.flat)
            if self.use_RND:
                self.rnd.load_state_dict(t.load(path + '/RND_weights.pth', weights_only=True, map_location=self.device))
        except FileNotFoundError:
            pass

    def save_weights(self, path: str):
        t.save(self.policy.state_dict(), path + '/Policy_weights.pth')
        if self.use_RND:
            t.save(self.rnd.state_dict(), path + '/RND_weights.pth')
