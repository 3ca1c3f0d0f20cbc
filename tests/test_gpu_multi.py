"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): in-kernel peer-memory gradient exchange vs NCCL, replica
synchronisation, and the sharded update against the oracle fed the same minibatch permutation.  The log of a
2-GPU run of this file is kept under profiles/ (the driver's GPU test box has one GPU; bench.py --gpus N additionally reports
`replicas_identical` and a peer-exchange-vs-NCCL check on every scaling run)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_two_ranks(tmp_path, port, extra_env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "host", "p2p_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(extra_env or {})))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert (tmp_path / "rank0").read_text() == "ok" and (tmp_path / "rank1").read_text() == "ok"
    return "\n".join(l for l in r.stdout.splitlines() if l.startswith("[parity]"))


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_exchange_matches_nccl(tmp_path):
    print(_run_two_ranks(tmp_path, 29631))


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 1, reason="needs a GPU")
def test_sharded_update_with_both_ranks_on_one_gpu(tmp_path, capsys):
    """The same two-rank worker on a ONE-GPU box (the driver's GPU test box has one): both processes time-slice GPU 0, gloo carries
    the collectives, the gradient exchange inside the step kernel goes through CUDA IPC buffers exactly as between two GPUs.
    Checks what the two-GPU test checks - peer exchange == allreduce path bit for bit (eager and CUDA graph), replicas identical,
    RND run identical, and the sharded learn() against oracle.ppo.learn fed the union-of-k-th-chunks schedule."""
    out = _run_two_ranks(tmp_path, 29633, dict(PRL_SAME_DEVICE_PEERS="1", PRL_TC_TIMEOUT_MS="8000", CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]))
    with capsys.disabled():
        print("\n" + out)
