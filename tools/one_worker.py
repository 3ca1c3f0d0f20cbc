"""A few AsyncPPO.worker() calls at the headline configuration (for ncu: --kernel-name regex:k_rollout)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sys.argv = [sys.argv[0]]
import torch as t  # noqa: E402

args = bench.parse()
dev = t.device("cuda", 0)
t.cuda.set_device(dev)
run = bench.Runner(args.cfg, None, dev, t.empty(1 << 20, dtype=t.uint8, device=dev))
run.step(False)   # one learn(): the policy the later rollouts act with keeps every env alive for most of the horizon
for _ in range(n):
    run.ap.worker()
    run.ppo.memory.clear()
t.cuda.synchronize()
print("done")
