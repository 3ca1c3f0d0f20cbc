"""ORACLE / test infrastructure: run the UNMODIFIED reference (the copies staged by oracle/stage_reference.py under
oracle/_ref/reference/) on the host CPU and time it - the `kind: "reference"` arm of bench.py.

What is executed verbatim: the reference's `AsyncPPO.worker()` (EnvVectorizer per-env Python loop, VecMemory lists, utils.*,
PPO.get_action) and `PPO.learn()` (old-policy evaluation through DataLoader batches, the O(N^2) list-insert GAE, k_epochs x
minibatch autograd + clip + AdamW) - /root/reference/AsyncTools/AsyncPPO.py:104-165 and PPO/PPO.py:82-260.  What is NOT the
reference: `gymnasium` (not installable here) is a stub module whose `make()` returns oracle/envs.py's per-env Python objects
- our restatement of gymnasium's classic-control physics with the same reset()/step() interface; it is labelled as such
in the output.  The reference picks CUDA when it sees one (PPO/PPO.py:11): the caller hides the GPUs
(CUDA_VISIBLE_DEVICES="") so that this is the reference's CPU path.

Runs as its own process (the reference's packages are called `PPO` and `AsyncTools`, like the drop-in's):
    python oracle/ref_runner.py '{"env_id": "CartPole-v1", "envs": 32, "horizon": 500, "steps": 3, "warmup": 1, "ppo": {...}}'
prints one JSON line {"env_steps", "seconds", "threads", "steps", "rows_per_step", ...}."""
from __future__ import annotations

import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGE = os.path.join(HERE, "_ref", "reference")


def load_reference():
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import envs as oenvs

    gym = types.ModuleType("gymnasium")

    class Env:
        pass

    gym.Env = Env
    gym.make = lambda env_id, max_episode_steps=None, **kw: oenvs.make(env_id, max_episode_steps=max_episode_steps)
    sys.modules["gymnasium"] = gym
    if not os.path.isdir(os.path.join(STAGE, "PPO")):
        raise SystemExit("the reference is not staged under oracle/_ref/reference (run build() where /root/reference exists)")
    sys.path.insert(0, STAGE)
    import AsyncTools.AsyncPPO as ref_async
    import PPO as ref_ppo

    assert os.path.abspath(ref_async.__file__).startswith(STAGE) and os.path.abspath(ref_ppo.__file__).startswith(STAGE)
    return gym, ref_async, ref_ppo


def run(cfg: dict) -> dict:
    import numpy as np
    import torch as t

    threads = int(cfg.get("threads") or os.cpu_count() or 1)
    t.set_num_threads(threads)
    gym, ref_async, ref_ppo = load_reference()
    assert str(ref_ppo.PPO.__module__).startswith("PPO") and not t.cuda.is_available(), "the reference arm must run on the CPU"
    t.manual_seed(int(cfg.get("seed", 0)))
    np.random.seed(int(cfg.get("seed", 0)))
    env = gym.make(cfg["env_id"], max_episode_steps=cfg.get("horizon"))
    ppo = ref_ppo.PPO(**cfg["ppo"])
    runner = ref_async.AsyncPPO(env=env, ppo=ppo, num_envs=int(cfg["envs"]), steps=10 ** 9)
    total_steps, total_s, rows = 0, 0.0, []
    for it in range(int(cfg.get("warmup", 0)) + int(cfg["steps"])):
        runner.step_score = 0
        runner.reward_score = 0
        t0 = time.perf_counter()
        runner.worker()                      # AsyncPPO.py:117-146
        n = int(runner.step_score)
        ppo.learn()                          # PPO.py:122-260 (returns early below batch_size, exactly like the reference)
        dt = time.perf_counter() - t0
        if it >= int(cfg.get("warmup", 0)):
            total_steps += n
            total_s += dt
            rows.append(n)
    return {"env_steps": total_steps, "seconds": total_s, "threads": threads, "steps": int(cfg["steps"]), "rows_per_step": rows,
            "envs": int(cfg["envs"]), "torch": t.__version__,
            "physics": "oracle/envs.py per-env Python objects behind a stub gymnasium (gymnasium itself is not installable here)"}


def run_subprocess(cfg: dict, timeout: float = 1800.0) -> dict:
    """Launch this file as a CPU-only child process (what bench.py calls)."""
    import subprocess

    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONDONTWRITEBYTECODE="1")
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), json.dumps(cfg)], env=env, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("reference runner failed: " + r.stderr[-2000:])
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


if __name__ == "__main__":
    print(json.dumps(run(json.loads(sys.argv[1]))), flush=True)
