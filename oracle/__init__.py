"""ORACLE package — test infrastructure only (see oracle/README.md).

CPU restatements of the reference's data-parallel PPO hot path, used to CHECK the CUDA product:
  oracle.envs     numpy env physics (gymnasium classic-control restated; parity unpinned vs gymnasium)
  oracle.cref     ctypes view of oracle/c/prl_oracle.c (same algorithms in plain C, fast)
  oracle.vec      EnvVectorizer / VecMemory / utils / AsyncPPO.worker restated in numpy
  oracle.ppo      ActorCritic / RND / PPO.learn restated with torch-CPU fp32 ops
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.
"""
