"""CPU-side checks (no GPU): the C ABI library loads and exports every symbol the header declares, the product refuses
to run without a CUDA device (no CPU fallback), and the host-side logic (PPO.memory list semantics, env descriptors,
shard bounds, the sharded minibatch schedule, the gloo allreduce plumbing) behaves like the reference's."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch as t

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "parallel-reinforcement-learning_b200")


def header_symbols(name="prl_b200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(prl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from prl_b200 import _lib

    lib = _lib.load_library()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/prl_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
        assert not n.startswith("prl_test_"), "test hooks do not belong in the product library"
    # the parity-test hooks live in their own library with their own header; the product exports none of them
    tlib = _lib.load_test_library()
    tnames = header_symbols("prl_b200_test.h")
    assert sorted(tnames) == sorted(_lib.TEST_FUNCTIONS + ("prl_test_last_error",))
    for n in tnames:
        assert hasattr(tlib, n), f"{n} declared in include/prl_b200_test.h but not exported"
        assert not hasattr(lib, n), f"{n} leaked into the product library"
    assert sorted(_lib.PROTOTYPES) == sorted(names + list(_lib.TEST_FUNCTIONS))  # nothing is bound that no header declares
    assert lib.prl_version() == 100


def test_host_side_introspection_calls():
    """Entry points that do no GPU work: env facts and parameter counts (reference: 9 027 / 13 187 / 9 220 / 966)."""
    from prl_b200 import _lib, ops

    info = _lib.env_info("CartPole-v1")
    assert (info["S"], info["O"], info["A"], info["continuous"], info["max_steps"]) == (4, 4, 2, False, 500)
    info = _lib.env_info("Pendulum-v1")
    assert (info["S"], info["O"], info["A"], info["continuous"], info["max_steps"]) == (2, 3, 1, True, 200)
    info = _lib.env_info("Acrobot-v1")
    assert (info["S"], info["O"], info["A"], info["continuous"], info["max_steps"]) == (4, 6, 3, False, 500)
    assert ops.policy_param_count(False, 4, 2) == 9027
    assert ops.policy_param_count(True, 3, 1) == 13187
    assert ops.policy_param_count(False, 6, 3) == 9220
    assert ops.rnd_param_count(6, 6) == 966


@pytest.mark.skipif(t.cuda.is_available(), reason="checks the behaviour WITHOUT a GPU")
def test_product_fails_loudly_without_cuda():
    from prl_b200 import PrlError
    from PPO import PPO
    import AsyncTools

    with pytest.raises(PrlError):
        PPO(is_continuous=False, observ_dim=4, action_dim=2)
    with pytest.raises(PrlError):
        AsyncTools.utils.indexes_of_active_environments(4, np.zeros(4, bool))
    import prl_b200

    with pytest.raises(PrlError):
        AsyncTools.AsyncPPO.EnvVectorizer(prl_b200.make("CartPole-v1"), 4)


def test_missing_library_is_an_error(tmp_path):
    code = ("import sys; sys.path.insert(0, %r); from prl_b200 import _lib; _lib.LIB_PATH = %r\n"
            "try:\n    _lib.load_library()\nexcept _lib.PrlError as e:\n    print('RAISED', e)\n") % (PKG, str(tmp_path / "nope.so"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "RAISED" in out and "no CPU fallback" in out


def test_memory_list_semantics_match_reference():
    """PPO/Memory.py:7-30: four flat lists of float32 items; push casts; clear empties; `+=` extends."""
    from PPO import Memory

    m = Memory()
    for i in range(5):
        m.push(state=np.random.randn(4), action=np.random.randint(0, 2, size=(1,)), reward=np.random.rand(1), done=np.bool_(i == 4))
    assert len(m.states) == len(m.actions) == len(m.rewards) == len(m.dones) == 5
    assert m.states[0].dtype == np.float32 and m.states[0].shape == (4,)
    assert m.actions[0].dtype == np.float32 and m.actions[0].shape == (1,)
    assert m.dones[4].dtype == np.float32 and float(m.dones[4]) == 1.0
    m.rewards += [np.float32(2.0)]
    assert len(m.rewards) == 6 and float(list(m.rewards)[-1]) == 2.0
    m.clear()
    assert len(m.states) == 0 and len(m.rewards) == 0


def test_env_descriptor_stands_in_for_gym_make():
    import prl_b200
    from prl_b200.envs import describe

    e = prl_b200.make("CartPole-v1")
    assert e.observation_space.shape[0] == 4 and e.action_space.n == 2 and e.spec.max_episode_steps == 500
    p = prl_b200.make("Pendulum-v1", max_episode_steps=64)
    assert p.observation_space.shape[0] == 3 and p.action_space.shape == (1,) and p.max_episode_steps == 64

    class FakeGym:  # what a gymnasium env exposes
        class spec:
            id, max_episode_steps = "Acrobot-v1", 500
    assert describe(FakeGym()).env_id == "Acrobot-v1"
    m = prl_b200.make("MountainCar-v0")
    assert m.observation_space.shape[0] == 2 and m.action_space.n == 3 and m.spec.max_episode_steps == 200
    with pytest.raises(ValueError):
        prl_b200.make("LunarLander-v3")


def test_shard_bounds_and_minibatch_schedule():
    from prl_b200.dist import minibatch_schedule, shard_bounds

    for E, W in [(65536, 8), (10, 4), (7, 8), (1, 1)]:
        spans = [shard_bounds(E, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == E
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    # 2 ranks, global minibatch 8 -> 4 local rows per rank and step; ranks run out at different steps
    mb_local, n_mb, counts = minibatch_schedule([10, 5], 8)
    assert (mb_local, n_mb, counts) == (4, 3, [8, 5, 2])
    assert sum(counts) == 15
    assert minibatch_schedule([16], 8) == (8, 2, [8, 8])


WORKER = r'''
import os, sys
sys.path[:0] = [%(root)r, %(pkg)r]
import torch, torch.distributed as td
from prl_b200 import dist
comm = dist.init_from_env(backend="gloo")
assert comm is not None and comm.world_size == 2 and dist.active() is comm
g = torch.full((9027,), float(comm.rank + 1))
comm.allreduce_(g)
assert torch.all(g == 3.0)
stats = torch.tensor([1.0 * (comm.rank + 1), 2.0, 10.0 + comm.rank, 0.0], dtype=torch.float64)
comm.allreduce_(stats)
assert stats.tolist() == [3.0, 4.0, 21.0, 0.0]
n_all = comm.allgather_int(100 + 7 * comm.rank)
assert n_all == [100, 107]
lo, hi = dist.shard_bounds(65537, comm.rank, comm.world_size)
sizes = comm.allgather_int(hi - lo)
assert sum(sizes) == 65537
mb_local, n_mb, counts = dist.minibatch_schedule(n_all, 64)
assert mb_local == 32 and n_mb == 4 and counts == [64, 64, 64, 15]
comm.barrier()
td.destroy_process_group()
open(os.path.join(%(out)r, "ok%%d" %% comm.rank), "w").write("ok")
'''


def test_two_rank_gloo_plumbing(tmp_path):
    """world_size-2 run of the data-parallel plumbing on CPU (gloo): gradient allreduce, advantage statistics, schedule."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, pkg=PKG, out=str(tmp_path)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_np_rng_header_equals_numpy_on_the_host(tmp_path):
    """csrc/np_rng.cuh compiled as host C++: PCG64(SeedSequence(seed)) raw outputs and doubles equal numpy's, bit for bit."""
    import ctypes as C

    so = str(tmp_path / "np_rng_check.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-D__host__=", "-D__device__=", "-I", os.path.join(PKG, "csrc"),
                           os.path.join(ROOT, "tests", "host", "np_rng_check.cpp"), "-o", so])
    lib = C.CDLL(so)
    seeds = np.array([0, 1, 42, 2 ** 32, 2 ** 63 + 12345, 2 ** 64 - 1] + list(range(1234, 1334)), dtype=np.uint64)
    n, draws = len(seeds), 12
    raw = np.zeros((n, draws), np.uint64)
    uni = np.zeros((n, draws), np.float64)
    lib.np_rng_check(seeds.ctypes.data_as(C.c_void_p), n, draws, raw.ctypes.data_as(C.c_void_p), uni.ctypes.data_as(C.c_void_p))
    for i, s in enumerate(seeds):
        bg = np.random.PCG64(np.random.SeedSequence(int(s)))
        assert np.array_equal(raw[i], bg.random_raw(draws)), int(s)
        assert np.array_equal(uni[i], np.random.Generator(bg).random(draws)), int(s)


def test_staged_reference_files_are_unmodified():
    """oracle/stage_reference.py copies the reference's PPO/, AsyncTools/ and unittests/ byte for byte (what the GPU box runs as
    "the reference's own unittests" and "the verbatim reference arm" really is the reference)."""
    import pytest

    from oracle import stage_reference as sr

    if not os.path.isdir(sr.REF):
        pytest.skip("/root/reference is not present on this box")
    assert sr.stage() and sr.verify() == []
