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


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_exchange_matches_nccl(tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "host", "p2p_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert (tmp_path / "rank0").read_text() == "ok" and (tmp_path / "rank1").read_text() == "ok"
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("[parity]")))
