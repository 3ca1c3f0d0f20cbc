#!/usr/bin/env python
"""bench.py - the reference's headline metric on B200: env-steps/s end to end (rollout + GAE + PPO update).

Workload (BASELINE.json configs[1]): CartPole-v1 discrete, num_envs = 65 536 per GPU, one episode per env with
TimeLimit T = 128 ("synthetic rollout T=128"), then PPO.learn() with k_epochs = 11 and mini_batch_size = 65 536
(SURVEY.md section 8d).  A "step" = one AsyncPPO.worker() + one PPO.learn(): reset, fused rollout kernel, env-major
transfer, old-policy evaluation, float32 GAE, advantage normalisation, 11 x ceil(N / 65 536) optimiser steps.

  python bench.py [--gpus N] [--steps K] [--warmup W]          the B200 arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                          the CPU arm: the oracle port of the reference's
                                                                Python path on the host cores, bounded sample

One JSON line on stdout (rank 0).  `value`: device-resident (random resets drawn on the GPU); `e2e`: the same steps
through the public API with the start states coming from pinned host memory every step and the updated weights +
scores read back every step.  `roofline`: the dominant kernel, timed live with CUDA events on the launching stream.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "parallel-reinforcement-learning_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "env_steps_per_sec_end_to_end_rollout_gae_update"
UNIT = "env-steps/s"
# algorithmic FLOPs of one sample-epoch of the update, CartPole shapes: forward 2*(O*64 + 2*64^2 + 64*A
# + 64) = 17 280, forward + backward ~ 3x
FLOPS_PER_SAMPLE_EPOCH = 51_840.0
# bf16 MMA flops the tensor-core kernel issues per row: forward 2 heads x 24 MMAs (128x64x16), dgrad 2 x 24, wgrad 2 x 32,
# trunk wgrad 32 x (128x16x16), all x 2 flops, per 128-row tile
EXECUTED_BF16_FLOPS_PER_ROW = (2 * 24 * 128 * 64 * 16 * 2 + 2 * 24 * 128 * 64 * 16 * 2 + 2 * 32 * 128 * 64 * 16 * 2 + 32 * 128 * 16 * 16 * 2) / 128.0


# BASELINE.json `configs`, in its order (c1 = configs[0] ... c4 = configs[3]; configs[4], the size sweep, is the `configs.c5_*`
# block of the c2 line).  `envs` / `mini_batch` are per GPU (weak scaling); `ref_envs` = the env count the CPU reference arm
# really runs per step (a bounded sample; c1 is small enough to run whole).
CONFIGS = {
    "c1": dict(env_id="CartPole-v1", envs=32, horizon=500, k_epochs=11, mini_batch=512, batch=1024, ref_envs=32,
               ppo=dict(is_continuous=False, observ_dim=4, action_dim=2),
               note="README quick start (README.md:27-50): 32 envs, gymnasium's default 500-step TimeLimit, batch 1024 / mini 512"),
    "c2": dict(env_id="CartPole-v1", envs=65536, horizon=128, k_epochs=11, mini_batch=65536, batch=1024, ref_envs=512,
               ppo=dict(is_continuous=False, observ_dim=4, action_dim=2), note="the configuration the metric is quoted on"),
    "c3": dict(env_id="Pendulum-v1", envs=262144, horizon=200, k_epochs=11, mini_batch=262144, batch=1024, ref_envs=256,
               ppo=dict(is_continuous=True, observ_dim=3, action_dim=1, action_scaling=2.0),
               note="tanh-Gaussian policy, gymnasium's 200-step TimeLimit"),
    "c4": dict(env_id="Acrobot-v1", envs=65536, horizon=128, k_epochs=11, mini_batch=65536, batch=1024, ref_envs=512,
               ppo=dict(is_continuous=False, observ_dim=6, action_dim=3, use_RND=True, beta=0.001),
               note="RND predictor / target MLPs in the path"),
}
STATE_BOX = {"CartPole-v1": ([-0.05] * 4, [0.05] * 4), "Pendulum-v1": ([-3.141592653589793, -1.0], [3.141592653589793, 1.0]),
             "Acrobot-v1": ([-0.1] * 4, [0.1] * 4), "MountainCar-v0": ([-0.6, 0.0], [-0.4, 0.0]),
             "MountainCarContinuous-v0": ([-0.6, 0.0, 0.0], [-0.4, 0.0, 0.0])}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configs[i-1]")
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the config's)")
    ap.add_argument("--horizon", type=int, default=None)
    ap.add_argument("--k-epochs", type=int, default=None)
    ap.add_argument("--mini-batch", type=int, default=None, help="minibatch rows per GPU")
    ap.add_argument("--cpu-sample-envs", type=int, default=None, help="envs per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the `configs` block (c1, c3, c4, the c5 sweep, the ragged regime)")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    a.envs = a.envs or c["envs"]
    a.horizon = a.horizon or c["horizon"]
    a.k_epochs = a.k_epochs or c["k_epochs"]
    a.mini_batch = a.mini_batch or c["mini_batch"]
    a.cpu_sample_envs = a.cpu_sample_envs or c["ref_envs"]
    a.cfg = dict(c, envs=a.envs, horizon=a.horizon, k_epochs=a.k_epochs, mini_batch=a.mini_batch, ref_envs=a.cpu_sample_envs, name=a.config)
    return a


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_port_steps(envs: int, horizon: int, k_epochs: int, mini_batch: int, n_steps: int, warmup: int, threads: int):
    """The reference's CPU path restated (oracle/): per-env Python env objects stepped in a loop, list-of-lists
    VecMemory, torch-CPU ActorCritic for get_action, then learn() = torch-CPU float32 autograd + AdamW.  Returns
    (env steps, seconds) over the timed steps.  This is the checker code being TIMED as the baseline, nothing more."""
    import numpy as np
    import torch as t

    from oracle import envs as oenvs
    from oracle import ppo as oppo
    from oracle import vec as ovec

    t.set_num_threads(threads)
    t.manual_seed(0)
    np.random.seed(0)
    cont, O, A = False, 4, 2
    shapes = oppo.param_shapes(cont, O, A)
    params = {}
    for k in oppo.param_keys(cont):  # reference init (ActorCritic.py:66-80)
        w = t.empty(shapes[k])
        if k.endswith("0.weight") or k.endswith("3.weight"):
            t.nn.init.xavier_uniform_(w)
        elif k.endswith("3.bias"):
            t.nn.init.normal_(w, 0, 0.01)
        elif k.endswith("1.weight"):
            w.fill_(1.0)
        else:
            w.zero_()
        params[k] = w
    vec = ovec.Vectorizer(oenvs.make("CartPole-v1", max_episode_steps=horizon), envs)
    opt = oppo.AdamW([params[k] for k in oppo.param_keys(cont)], lr=1e-3)

    def act(states):
        with t.no_grad():
            _, (probs,) = oppo.dist_params(params, cont, t.from_numpy(states))
            return t.multinomial(probs, 1).squeeze(-1).numpy()

    total_steps, total_s = 0, 0.0
    for it in range(warmup + n_steps):
        t0 = time.perf_counter()
        mem = ovec.FlatMemory()
        _, ss = ovec.worker(vec, ovec.PerEnvMemory(envs), act, mem)
        m = dict(states=np.array(mem.states, np.float32), actions=np.array(mem.actions, np.float32),
                 rewards=np.array(mem.rewards, np.float32), dones=np.array(mem.dones, np.float32))
        oppo.learn(params, cont, m, lr=1e-3, k_epochs=k_epochs, policy_clip=0.2, gae_lambda=0.95, gamma=0.995,
                   mini_batch_size=mini_batch, opt=opt)
        dt = time.perf_counter() - t0
        if it >= warmup:
            total_steps += ss
            total_s += dt
    return total_steps, total_s


def cpu_reference(cfg, envs: int, n_steps: int, warmup: int, threads: int):
    """The CPU arm on `envs` envs of config `cfg`: the UNMODIFIED reference (oracle/_ref/reference, staged by build();
    kind "reference") in a CPU-only child process, else - where nothing is staged - the oracle port (kind "port", CartPole only).
    Returns (env steps, seconds, kind, note)."""
    from oracle import ref_runner, stage_reference

    if stage_reference.staged():
        ppo = dict(cfg["ppo"], lr=1e-3, k_epochs=cfg["k_epochs"], batch_size=cfg["batch"], mini_batch_size=cfg["mini_batch"])
        r = ref_runner.run_subprocess(dict(env_id=cfg["env_id"], envs=envs, horizon=cfg["horizon"], steps=n_steps, warmup=warmup,
                                           threads=threads, ppo=ppo))
        return r["env_steps"], r["seconds"], "reference", ("the unmodified reference (AsyncPPO.worker + PPO.learn, torch " + r["torch"] +
                                                           " on the CPU); physics: " + r["physics"])
    if cfg["env_id"] != "CartPole-v1":
        raise SystemExit("the reference is not staged (run build() where /root/reference exists) and the port arm covers CartPole only")
    steps, secs = cpu_port_steps(envs, cfg["horizon"], cfg["k_epochs"], cfg["mini_batch"], n_steps, warmup, threads)
    return steps, secs, "port", "oracle/ restatement of the reference's Python path"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cfg = args.cfg
    steps, secs, kind, note = cpu_reference(cfg, args.cpu_sample_envs, args.steps, args.warmup, threads)
    v = steps / secs
    sample = (f"{args.cpu_sample_envs} envs per step (the B200 arm runs {args.envs} per GPU), same env, T={args.horizon}, k_epochs={args.k_epochs}, "
              f"mini_batch={args.mini_batch}; {note}")
    conf = workload_config(cfg, 1)
    conf["num_envs_run_by_this_arm"] = args.cpu_sample_envs
    conf["workload"] += f" [this CPU arm ran num_envs={args.cpu_sample_envs} per step: a bounded sample]" if args.cpu_sample_envs != args.envs else ""
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 networks / f64 physics", "data": "synthetic",
        "config": conf,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(cfg, world):
    p = cfg["ppo"]
    kind = "continuous (tanh-Gaussian, action_scaling=%s)" % p.get("action_scaling") if p["is_continuous"] else "discrete"
    return {"workload": f"{cfg['env_id']} {kind}{', use_RND beta=%s' % p['beta'] if p.get('use_RND') else ''}, num_envs={cfg['envs']}/GPU, one episode per env "
                        f"with TimeLimit T={cfg['horizon']}, PPO.learn k_epochs={cfg['k_epochs']}, mini_batch_size={cfg['mini_batch']}/GPU, "
                        f"batch_size={cfg['batch']}, lr=1e-3 (BASELINE configs[{int(cfg['name'][1]) - 1}]: {cfg['note']})",
            "baseline_config": cfg["name"], "env_id": cfg["env_id"],
            "num_envs_per_gpu": cfg["envs"], "horizon": cfg["horizon"], "k_epochs": cfg["k_epochs"], "mini_batch_per_gpu": cfg["mini_batch"],
            "parallelism": f"env-sharded dp{world}",
            "l2": "256 MiB buffer rewritten between steps inside the timed region (L2 flush); at c2 the [T][C][E] rollout buffer alone is 235 MB > 126 MB L2"}


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def summary(self, t0, t1):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 <= ts <= t1 + 0.2):
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- B200 arm
class Runner:
    """One configuration through the drop-in API: AsyncPPO.worker() + PPO.learn() per step."""

    def __init__(self, cfg, comm, dev, flush, fresh_policy_every_step=False):
        import numpy as np
        import torch as t

        import prl_b200
        from AsyncTools.AsyncPPO import AsyncPPO
        from PPO import PPO

        self.t, self.cfg, self.comm, self.dev, self.flush = t, cfg, comm, dev, flush
        rank = comm.rank if comm else 0
        world = comm.world_size if comm else 1
        E = cfg["envs"]
        t.manual_seed(0)  # identical initial weights and sampling seed on every rank
        self.ppo = ppo = PPO(lr=1e-3, k_epochs=cfg["k_epochs"], batch_size=cfg["batch"], mini_batch_size=cfg["mini_batch"] * world, **cfg["ppo"])
        ppo.show_progress = False
        ppo.use_cuda_graph = True   # one captured epoch replayed k_epochs times (see PPO.learn)
        ppo.graph_collectives = os.environ.get("PRL_GRAPH_COLLECTIVES", "1") == "1"   # NCCL allreduce captured in the graph too
        # sharded runs: gradient exchange over NVLink peer memory inside the step kernel (0 = NCCL allreduce between grad and AdamW)
        ppo.peer_exchange = os.environ.get("PRL_PEER_EXCHANGE", "1") == "1"
        # (sharded: PPO broadcasts rank 0's replica at construction and keys the sampling / reset streams by rank - prl_b200.dist.rank_seed)
        self.ap = AsyncPPO(env=prl_b200.make(cfg["env_id"], max_episode_steps=cfg["horizon"]), ppo=ppo, num_envs=E, steps=1)
        lo, hi = STATE_BOX[cfg["env_id"]]
        rng = np.random.default_rng(1234 + rank)
        self.host_states = t.from_numpy(rng.uniform(lo, hi, (E, len(lo)))).pin_memory()       # e2e: start states from the host
        self.host_weights = t.empty(ppo.policy.flat.numel(), dtype=t.float32).pin_memory()   # e2e: results read back
        self.host_scores = t.empty(2, dtype=t.float64).pin_memory()
        # ragged regime: every step starts from the freshly initialised policy (short, unequal episodes: the shrinking-batch path)
        self.fresh = None
        if fresh_policy_every_step:
            o = ppo.optimizer
            self.fresh = [(x, x.clone()) for x in (ppo.policy.flat, ppo.policy_old.flat, o.exp_avg, o.exp_avg_sq, o.step_dev)]
            self.fresh_count = o.step_count

    def step(self, e2e: bool):
        t, ap, ppo = self.t, self.ap, self.ppo
        if self.fresh is not None:
            for dst, src in self.fresh:
                dst.copy_(src)
            ppo.optimizer.step_count = self.fresh_count
        ap.step_score = 0
        ap.reward_score = 0
        ap.worker(initial_states=self.host_states if e2e else None)
        n = int(ap.step_score)
        ppo.learn()
        if e2e:
            self.host_weights.copy_(ppo.policy.flat, non_blocking=True)
            self.host_scores.copy_(ap._scores, non_blocking=True)
            t.cuda.current_stream().synchronize()
        self.flush.fill_(1)
        return n

    def timed(self, e2e: bool, k: int, profile: bool = False):
        from prl_b200 import _lib

        t, comm = self.t, self.comm
        if comm:
            comm.barrier()
        t.cuda.synchronize()
        counts0 = dict(_lib.CALL_COUNTS)
        if profile:
            _lib.profile_calls(True)
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        n = sum(self.step(e2e) for _ in range(k))
        e1.record()
        t.cuda.synchronize()
        w1 = time.time()
        prof = _lib.profile_calls(False) if profile else None
        ms = t.tensor([e0.elapsed_time(e1)], dtype=t.float64, device=self.dev)
        tot = t.tensor([float(n)], dtype=t.float64, device=self.dev)
        if comm:
            comm.allreduce_max_(ms)
            comm.allreduce_(tot)
            comm.barrier()
        counts = {k_: v - counts0.get(k_, 0) for k_, v in _lib.CALL_COUNTS.items() if v - counts0.get(k_, 0)}
        return float(tot.item()), float(ms.item()), counts, prof, (w0, w1)

    def e2e_bytes(self):
        return int(self.host_states.numel() * 8), int(self.host_weights.numel() * 4 + 16 + 16)

    def close(self):
        self.ppo.close()
        self.ap.buffer = None
        self.ap = self.ppo = None
        import gc

        gc.collect()
        self.t.cuda.empty_cache()


def short_line(cfg, comm, dev, flush, steps, warmup, fresh=False):
    """One more configuration, measured like the headline (device-resident `value` + host-to-host `e2e`), in short form."""
    world = comm.world_size if comm else 1
    r = Runner(cfg, comm, dev, flush, fresh_policy_every_step=fresh)
    for _ in range(warmup):
        r.step(False)
    r.step(True)
    n_dev, ms_dev, counts, _, _ = r.timed(False, steps)
    n_e2e, ms_e2e, _, _, _ = r.timed(True, steps)
    h2d, d2h = r.e2e_bytes()
    from prl_b200 import _lib

    out = {"value": n_dev / (ms_dev * 1e-3), "unit": UNIT, "ms_per_step": ms_dev / steps, "env_steps_per_step": n_dev / steps, "steps": steps, "warmup": warmup,
           "n_gpus": world, "e2e": {"value": n_e2e / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "gpu_launches": _lib.launches(counts), "update_path": r.ppo.update_path, "config": workload_config(cfg, world)}
    r.close()
    return out


def run_b200(args):
    import numpy as np
    import torch as t

    from prl_b200 import _lib
    from prl_b200 import dist as pdist

    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        del os.environ["NCCL_DEBUG"]   # keep stdout to the one JSON line (at these levels NCCL prints a version banner there)
    comm = pdist.init_from_env()
    rank = comm.rank if comm else 0
    world = comm.world_size if comm else 1
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 under torch.distributed.run")
    dev = t.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    t.cuda.set_device(dev)

    cfg = args.cfg
    E, T = args.envs, args.horizon
    flush = t.empty(256 << 20, dtype=t.uint8, device=dev)
    run = Runner(cfg, comm, dev, flush)
    ppo, ap = run.ppo, run.ap
    step, timed = run.step, run.timed
    host_states, host_weights = run.host_states, run.host_weights

    # sharded parity, visible to the driver on every scaling run: ONE rollout consumed twice from the same replica state - by the
    # in-kernel peer-memory exchange (the timed path) and by the NCCL-allreduce path - must give the same weights (bit-identical
    # at 2 ranks: a two-addend sum is commutative; at more ranks NCCL's reduction order differs from the kernel's rank order)
    p2p_check = None
    if comm is not None and ppo.peer_exchange:
        ap.worker()
        n_rows = ppo.memory._dev_count
        keep = [x.clone() for x in (ppo.policy.flat, ppo.policy_old.flat, ppo.optimizer.exp_avg, ppo.optimizer.exp_avg_sq, ppo.optimizer.step_dev)]
        k0, cnt0, ppo.k_epochs = ppo.k_epochs, ppo.optimizer.step_count, 1
        res = []
        for peer in (True, False):
            for dst, src in zip((ppo.policy.flat, ppo.policy_old.flat, ppo.optimizer.exp_avg, ppo.optimizer.exp_avg_sq, ppo.optimizer.step_dev), keep):
                dst.copy_(src)
            ppo.optimizer.step_count = cnt0
            ppo.memory._dev_count = n_rows          # the same rows again (learn() only reads them)
            ppo.peer_exchange = peer
            ppo.learn()
            res.append(ppo.policy.flat.clone())
        ppo.peer_exchange, ppo.k_epochs = True, k0
        diff = (res[0] - res[1]).abs().max().reshape(1).double()
        comm.allreduce_max_(diff)
        p2p_check = {"optimizer_steps": int(ppo.optimizer.step_count - cnt0), "max_abs_weight_diff_vs_nccl_path": float(diff.item()),
                     "peer_path_used": bool(ppo._p2p_ok)}
    for _ in range(args.warmup):
        step(False)
    step(True)  # touch the e2e path once
    clocks = ClockSampler(dev.index) if rank == 0 else None
    time.sleep(0.3)
    ppo.update_events = []   # PPO.learn brackets its optimisation loop with CUDA events (2 per learn(): nothing next to 1 408 launches)
    n_dev, ms_dev, counts, _, (w0, w1) = timed(False, args.steps, profile=False)
    upd = ppo.update_events
    ppo.update_events = None
    upd_ms, upd_launches = sum(a.elapsed_time(b) for a, b, _ in upd), sum(n for _, _, n in upd)
    clk = clocks.summary(w0, w1) if clocks else None
    n_e2e, ms_e2e, _, _, _ = timed(True, args.steps, profile=False)
    # sharded: the replicas must still be identical after the timed region (weights and AdamW moments, bit for bit)
    replicas_identical = None
    if comm is not None:
        import torch.distributed as td

        mine = t.cat([ppo.policy.flat, ppo.policy_old.flat, ppo.optimizer.exp_avg, ppo.optimizer.exp_avg_sq])
        allw = [t.empty_like(mine) for _ in range(world)]
        td.all_gather(allw, mine)
        replicas_identical = all(t.equal(allw[0], w) for w in allw[1:]) and bool(t.isfinite(mine).all())
    # kernel attribution: ONE more step of the same workload, launch by launch (no CUDA graph), every entry-point call
    # bracketed by CUDA events on the launching stream
    ppo.use_cuda_graph = False
    n_prof, ms_prof, _, prof, _ = timed(False, 1, profile=True)
    ppo.use_cuda_graph = True

    # per-entry-point device time (CUDA events on the launching stream, this rank)
    per = {}
    for name, evs in prof.items():
        per[name] = {"calls": len(evs), "ms": sum(a.elapsed_time(b) for a, b in evs)}
    pk = peaks()
    roof = None
    gk = next((k_ for k_ in ("prl_ppo_step_tc_p2p", "prl_ppo_step_tc", "prl_ppo_grad_tc", "prl_ppo_grad") if k_ in per), None)
    kernel_names = {
        "prl_ppo_step_tc_p2p": "k_ppo_grad_tc (prl_ppo_step_tc_p2p: forward + loss + backward on tcgen05 bf16x3 MMAs, then gradient exchange over NVLink peer memory + clip + AdamW in the same cooperative launch)",
        "prl_ppo_step_tc": "k_ppo_grad_tc (prl_ppo_step_tc: forward + loss + backward on tcgen05 bf16x3 MMAs, then reduce + clip + AdamW in the same cooperative launch)",
        "prl_ppo_grad_tc": "k_ppo_grad_tc (prl_ppo_grad_tc: fused forward + loss + backward, tcgen05 bf16x3 MMAs + CUDA-core epilogues)",
        "prl_ppo_grad": "k_ppo_grad (prl_ppo_grad: fused forward + loss + backward, fp32 FMA path)"}
    if gk is not None:
        g = per[gk]
        rows_epochs = n_prof / world * args.k_epochs           # sample-epochs this rank pushed through the update kernel
        pp = cfg["ppo"]
        heads_out = 2 * pp["action_dim"] if pp["is_continuous"] else pp["action_dim"]
        flops_row = 3.0 * 2 * (pp["observ_dim"] * 64 + 2 * 64 * 64 + 64 * heads_out + 64)   # = 51 840 at CartPole shapes
        # the dominant kernel's launch duration INSIDE the timed region: CUDA events on the launching stream around the optimisation
        # loops of the timed learn() calls (graph replays: back-to-back launches) divided by the optimiser steps in them.  The
        # launch-by-launch profiling step below brackets every call with its own event pair, which adds ~5 us per launch.
        timed_launch_ms = upd_ms / max(upd_launches, 1)
        timed_rows_epochs = n_dev / world * args.k_epochs
        tf = timed_rows_epochs * flops_row / (upd_ms * 1e-3) / 1e12 if upd_ms > 0 else rows_epochs * flops_row / (g["ms"] * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r02_tc_traffic.json")   # dram__bytes_read + write of one launch (ncu --set full capture)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
        roof = {"kernel": kernel_names[gk], "bound": "tensor", "achieved": tf,
                "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_sustained"], "traffic": traffic,
                "traffic_unit": "bytes per launch (65 536 rows: 2.1 MB of row inputs + 36 KB of parameters are the algorithmic bytes)",
                "traffic_source": traffic_src,
                "peak_source": pk["source"] + " cuBLAS bf16, sustained figure (kernel timed inside a long step)",
                "algorithmic_flops_per_sample_epoch": flops_row, "avg_launch_ms": timed_launch_ms if upd_ms > 0 else g["ms"] / g["calls"],
                "launches_timed": upd_launches, "avg_launch_ms_profiling_step": g["ms"] / g["calls"],
                # what the tensor pipe really executes: every fp32-grade product is six bf16 MMAs (bf16x3 split) and the weight
                # gradients run on stacked piece windows - 44.0 MFLOP of bf16 MMAs per 128-row tile
                "executed_bf16_flops_per_sample_epoch": EXECUTED_BF16_FLOPS_PER_ROW,
                "executed_bf16_tflops": tf * EXECUTED_BF16_FLOPS_PER_ROW / flops_row if gk != "prl_ppo_grad" else None,
                "limiter": "serial hand-over between CUDA-core epilogues (GroupNorm / SiLU / loss forward + backward, bf16x3 splitting, column sums) and "
                           "the tensor core with one 17-warp CTA per SM: ncu issue slots 33 % busy (73 % inside the compute phases), tensor pipe 16 % "
                           "active, DRAM = algorithmic bytes (profiles/r02_k_ppo_grad_tc_ncu_summary.txt, DESIGN.md section 5)",
                "share_of_step": upd_ms / ms_dev if upd_ms > 0 else g["ms"] / ms_prof,
                "measured_on": "CUDA events on the launching stream around the optimisation loops of the timed steps (CUDA-graph replays); "
                               "avg_launch_ms_profiling_step: one launch-by-launch step after the timed region, one event pair per call"}
    hbm = {}
    # algorithmic bytes per transition (the fused worker adds two planes to the rollout and three fields to the transfer)
    bytes_per = {"prl_rollout": 102.0, "prl_rollout_eval": 110.0, "prl_gae": 16.0, "prl_gae_columns": 16.0, "prl_adv_normalize": 10.0,
                 "prl_buffer_transfer": 56.0 + 4.0 * E / max(n_prof / world, 1), "prl_buffer_transfer_ex": 80.0 + 4.0 * E / max(n_prof / world, 1)}
    for name, b in bytes_per.items():
        if name in per and per[name]["ms"] > 0:
            gbs = (n_prof / world) * b / (per[name]["ms"] * 1e-3) / 1e9
            hbm[name] = {"GB/s": gbs, "frac_of_" + pk["source"] + "_hbm": gbs / pk["hbm"], "bytes_per_transition": b, "ms": per[name]["ms"], "calls": per[name]["calls"]}

    # the two forward-only kernels are bound by the fp32 FMA pipe, not by HBM: report their TFLOP/s
    fma = {}
    for name, fl in (("prl_rollout", 8960.0), ("prl_rollout_eval", 17280.0), ("prl_policy_evaluate", 17280.0)):
        if name in per and per[name]["ms"] > 0:
            fma[name] = {"TFLOP/s_fp32": (n_prof / world) * fl / (per[name]["ms"] * 1e-3) / 1e12, "flops_per_row": fl, "ms": per[name]["ms"],
                         "fp32_fma_peak_TFLOP/s": 148 * 128 * 2 * 1.965e9 / 1e12}
    micro = hbm_microbench(pk) if rank == 0 and cfg["env_id"] == "CartPole-v1" else None
    out = {
        "metric": METRIC, "value": n_dev / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 networks / f64 physics", "data": "synthetic",
        "config": workload_config(cfg, world),
        "env_steps_per_step": n_dev / args.steps,
        "e2e": {"value": n_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(host_states.numel() * 8),
                "d2h_bytes_per_step": int(host_weights.numel() * 4 + 16 + 16), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": _lib.launches(counts),
        "replicas_identical": replicas_identical,
        "sharded_parity": p2p_check,
        "calls": counts,
        "clocks": clk,
        "roofline": roof,
        "hbm_kernels": hbm,
        "fma_kernels": fma,
        "hbm_micro": micro,
        "kernel_ms": {k: round(v["ms"], 3) for k, v in sorted(per.items(), key=lambda kv: -kv[1]["ms"])},
    }
    run.close()
    del run, ppo, ap, step, timed
    # ---- the other BASELINE configurations, measured the same way (short form), so that every config has a driver-visible number
    extras = {}
    if not args.no_extra_configs and args.config == "c2":
        def sub(name, **kw):
            return dict(CONFIGS[name], name=name, **kw)
        plan = []
        if world == 1:
            plan += [("c1", sub("c1"), 30, 5, False), ("c3", sub("c3"), 2, 1, False), ("c4", sub("c4"), 3, 2, False),
                     ("c2_ragged_fresh_policy", sub("c2", note="ragged regime: every step starts from the freshly initialised policy, so episodes "
                                                               "are short and unequal (the shrinking-batch path, SURVEY H5)"), 5, 3, True)]
        # configs[4]: CartPole sweep, envs per NODE = 2^16 .. 2^22 split over the GPUs of this run (2^16 per GPU = the headline line)
        for lg in (18, 20, 22):
            per_gpu = (1 << lg) // world
            if per_gpu >= 65536 and per_gpu != E:
                plan.append((f"c5_2^{lg}_envs_per_node", sub("c2", envs=per_gpu, note=f"configs[4] sweep point: 2^{lg} envs on {world} GPU(s)"), 2, 1, False))
        for name, c, k, w, fresh in plan:
            try:
                extras[name] = short_line(c, comm, dev, flush, k, w, fresh)
            except Exception as e:   # a failed extra must not lose the headline line
                extras[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
    if rank != 0:
        return
    if extras:
        out["configs"] = extras
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_cpu = args.cpu_sample_envs
        cs, csec, kind, note = cpu_reference(cfg, n_cpu, 1, 0, threads)
        out["cpu_baseline"] = {"value": cs / csec, "unit": UNIT, "cores": threads, "kind": kind,
                               "sample": f"1 step (worker + learn) of {n_cpu} envs (the B200 arm: {E}), same env, T={T}, k_epochs={args.k_epochs}, "
                                         f"mini_batch={args.mini_batch}; {cs} env-steps in {csec:.1f} s; {note}"}
        if extras and "c1" in extras and "value" in extras["c1"]:
            c1 = dict(CONFIGS["c1"], name="c1")
            cs, csec, kind, note = cpu_reference(c1, c1["envs"], 6, 1, threads)
            extras["c1"]["cpu_baseline"] = {"value": cs / csec, "unit": UNIT, "cores": threads, "kind": kind,
                                            "sample": f"the whole configuration (32 envs), 6 steps after 1 warm-up; {cs} env-steps in {csec:.1f} s; {note}"}
    print(json.dumps(out), flush=True)


def hbm_microbench(pk, E=65536, T=128):
    """The HBM-bound kernels of the path in isolation (outside the timed region): achieved GB/s at SURVEY 8(d)'s
    algorithmic bytes per unit at the C2 shapes (E = 65 536, T = 128).  Each figure is the average over ROT launches that
    run back to back on ROT DIFFERENT buffer sets (together several times the 126 MB L2, so every launch streams from
    HBM and launch latency is amortised as it is inside a step), CUDA events on the launching stream."""
    import torch as t

    from prl_b200 import ops

    dev = t.device("cuda", t.cuda.current_device())
    ROT = 4

    def timeit(fns, reps=5):
        for f in fns:
            f()
        t.cuda.synchronize()
        ms = 0.0
        for _ in range(reps):
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            for f in fns:
                f()
            e1.record()
            t.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        return ms / reps / len(fns)

    out = {}
    N = E * T
    sets = []
    for _ in range(ROT):
        sets.append(dict(r=t.ones(N, device=dev), v=t.rand(N, device=dev), ret=t.empty(N, device=dev), adv=t.empty(N, device=dev),
                         d=t.zeros(N, device=dev), stats=t.zeros(4, dtype=t.float64, device=dev)))
    ws = t.empty(64, dtype=t.uint8, device=dev)
    lens = t.full((E,), T, dtype=t.int32, device=dev)
    gen = t.Generator(device=dev); gen.manual_seed(7)
    for name, ep in (("gae_flat_T128", T), ("gae_flat_T20", 20), ("gae_flat_ragged_20_128", None)):
        for S in sets:
            S["d"].zero_()
            if ep is None:   # ragged episodes of 20..128 steps, as a partly trained CartPole policy produces
                ends = t.cumsum(t.randint(20, T + 1, (2 * E,), device=dev, generator=gen), 0) - 1
                S["d"][ends[ends < N]] = 1.0
                S["d"][N - 1] = 1.0
            else:
                S["d"][ep - 1::ep] = 1.0
        ms = timeit([lambda S=S: ops.gae(S["r"], S["d"], S["v"], 0.995, 0.95, out=S["ret"], ws=ws) for S in sets])
        out[name] = {"GB/s": N * 16 / ms / 1e6, "ms": ms, "bytes_per_transition": 16}
    for S in sets:
        S["d"].zero_(); S["d"].view(T, E)[T - 1] = 1.0
    ms = timeit([lambda S=S: ops.gae_columns(S["r"].view(T, E), S["d"].view(T, E), S["v"].view(T, E), lens, 0.995, 0.95, out=S["ret"].view(T, E))
                 for S in sets])
    out["gae_columns_T128"] = {"GB/s": N * 16 / ms / 1e6, "ms": ms, "bytes_per_transition": 16}
    ms = timeit([lambda S=S: (S["stats"].zero_(), ops.adv_normalize(S["ret"], S["v"], stats=S["stats"], phase=3, out=S["adv"])) for S in sets])
    out["adv_normalize"] = {"GB/s": N * 20 / ms / 1e6, "ms": ms, "bytes_per_transition": 20}
    del sets
    # device VecMemory -> PPO.memory (prl_buffer_transfer), every env with a full-length episode: 56 B per transition
    bufs = [ops.RolloutBuffer(E, T, 4, 1) for _ in range(2)]
    mems = [[t.empty(N, 4, device=dev), t.empty(N, 1, device=dev), t.empty(N, device=dev), t.empty(N, device=dev)] for _ in range(2)]
    total = t.zeros(1, dtype=t.int64, device=dev)

    def xfer(b, m):
        b.lengths.fill_(T)
        b.transfer(*m, 0, total)
    ms = timeit([lambda b=b, m=m: xfer(b, m) for b, m in zip(bufs, mems)])
    out["buffer_transfer"] = {"GB/s": N * 56 / ms / 1e6, "ms": ms, "bytes_per_transition": 56}
    del mems
    # standalone env step (per-step API), all envs active, 2^22 envs (428 MB per launch > L2): 102 B per env-step (CartPole)
    E2 = 1 << 22
    sim = ops.EnvState("CartPole-v1", E2, 1 << 30)
    sim.reset(1, 1)
    idx = t.arange(E2, dtype=t.int32, device=dev)
    acts = t.randint(0, 2, (E2,), dtype=t.int32, device=dev)
    ms = timeit([lambda: sim.step(idx, E2, acts)] * 3)
    out["env_step_cartpole"] = {"GB/s": E2 * 102 / ms / 1e6, "ms": ms, "bytes_per_env_step": 102, "envs": E2}
    del sim, idx, acts
    # teacher-forced fused rollout (physics + TimeLimit + mask + buffer write, no policy): E x T env-steps
    sims = [ops.EnvState("CartPole-v1", E, T) for _ in range(2)]
    tape = t.randint(0, 2, (T, E), dtype=t.int32, device=dev)
    scores = t.zeros(2, dtype=t.float64, device=dev)

    def taped(sim, buf):
        sim.reset(1, 1)
        ops.rollout(sim, buf, None, 1.0, 0, 1, scores, tape=tape)
    taped(sims[0], bufs[0]); scores.zero_(); taped(sims[0], bufs[0])
    n_steps = float(scores.cpu()[1])
    ms = timeit([lambda s_=s_, b=b: taped(s_, b) for s_, b in zip(sims, bufs)])
    out["rollout_taped_cartpole"] = {"GB/s": n_steps * 102 / ms / 1e6, "ms": ms, "bytes_per_env_step": 102, "env_steps": n_steps,
                                     "note": "65 536 envs = 14 warps per SM running ~22-step random-policy episodes of fp64 physics: latency-bound at this size"}
    del sims, bufs
    # the same at 2^20 envs (C5's range), where there are enough envs per SM for the buffer writes to matter
    E3 = 1 << 20
    sim3 = ops.EnvState("CartPole-v1", E3, T)
    buf3 = ops.RolloutBuffer(E3, T, 4, 1)
    tape3 = t.randint(0, 2, (T, E3), dtype=t.int32, device=dev)

    def taped3():
        sim3.reset(1, 1)
        ops.rollout(sim3, buf3, None, 1.0, 0, 1, scores, tape=tape3)
    scores.zero_(); taped3()
    n3 = float(scores.cpu()[1])
    ms = timeit([taped3] * 2, reps=3)
    out["rollout_taped_cartpole_1M_envs"] = {"GB/s": n3 * 102 / ms / 1e6, "ms": ms, "bytes_per_env_step": 102, "env_steps": n3}
    del sim3, buf3, tape3
    for k in out:
        out[k]["frac_of_" + pk["source"] + "_hbm"] = out[k]["GB/s"] / pk["hbm"]
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
        import torch.distributed as td

        if td.is_available() and td.is_initialized():
            td.destroy_process_group()


if __name__ == "__main__":
    main()
