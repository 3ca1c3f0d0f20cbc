"""Scratch diagnostics (GPU box): correctness and cycles per tcgen05.mma for the operand layouts the update kernel can use."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-reinforcement-learning_b200")]
import numpy as np, torch as t
from prl_b200 import ops

rng = np.random.default_rng(8)
q = lambda *shape: (rng.integers(-16, 17, shape) / 8.0).astype(np.float32)
dev = lambda a: t.from_numpy(np.ascontiguousarray(a)).cuda()
CH = 2048
def idesc(M, N, a_mn, b_mn): return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)
A, W = q(128, 64), q(128, 64)
Z, F, X = q(128, 128), q(128, 64), q(128, 16)
REPS = 64
def run(name, a, b, want, cfg):
    D, st = ops.test_umma(0, dev(a), dev(b), cfg + [1])
    exact = np.array_equal(D.cpu().numpy(), want)
    ops.test_umma(0, dev(a), dev(b), cfg + [REPS])
    cyc = ops.test_umma.last_cycles
    nm = cfg[4] * REPS
    print(f"{name:44s} st={st} exact={exact} cycles/MMA={cyc / nm:7.1f}")
# cfg = [a_cols, b_cols, n_out, idesc, nsteps, a_off, a_step, a_lbo, a_sbo, b_off, b_step, b_lbo, b_sbo, a_layout, b_layout]
run("fwd   N=128 no-swizzle",  A, W, A @ W.T,     [64, 64, 128, idesc(128, 128, 0, 0), 4, 0, 2 * CH, CH, 128, 0, 2 * CH, CH, 128, 0, 0])
run("dgrad N=64  no-swizzle",  A, W, A @ W[64:],  [64, 64, 64, idesc(128, 64, 0, 1), 4, 0, 2 * CH, CH, 128, 1024, 256, 128, CH, 0, 0])
run("wgrad N=64  no-swizzle",  Z, F, Z.T @ F,     [128, 64, 64, idesc(128, 64, 1, 1), 8, 0, 256, 128, CH, 0, 256, 128, CH, 0, 0])
run("wgrad0 N=16 no-swizzle",  Z, X, Z.T @ X,     [128, 16, 16, idesc(128, 16, 1, 1), 8, 0, 256, 128, CH, 0, 256, 128, CH, 0, 0])
# 128-byte swizzle: K-major: SBO = 1024 (8 rows), LBO ignored (16), K-step +32 B; MN-major: LBO = 16384 (next 64-col block), SBO = 1024, K-step +2048
run("fwd   N=128 SW128",       A, W, A @ W.T,     [64, 64, 128, idesc(128, 128, 0, 0), 4, 0, 32, 16, 1024, 0, 32, 16, 1024, 2, 2])
run("dgrad N=64  SW128",       A, W, A @ W[64:],  [64, 64, 64, idesc(128, 64, 0, 1), 4, 0, 32, 16, 1024, 64 * 128, 2048, 16384, 1024, 2, 2])
run("wgrad N=64  SW128",       Z, F, Z.T @ F,     [128, 64, 64, idesc(128, 64, 1, 1), 8, 0, 2048, 16384, 1024, 0, 2048, 16384, 1024, 2, 2])
run("wgrad0 N=16 SW128 x none",Z, X, Z.T @ X,     [128, 16, 16, idesc(128, 16, 1, 1), 8, 0, 2048, 16384, 1024, 0, 256, 128, CH, 2, 0])
run("fwd   N=64  no-swizzle",  A, W, A @ W[:64].T, [64, 64, 64, idesc(128, 64, 0, 0), 4, 0, 2 * CH, CH, 128, 0, 2 * CH, CH, 128, 0, 0])
run("fwd   N=64  SW128",       A, W, A @ W[:64].T, [64, 64, 64, idesc(128, 64, 0, 0), 4, 0, 32, 16, 1024, 0, 32, 16, 1024, 2, 2])
