"""Scratch diagnostics (GPU box): which shared-memory descriptor conventions does tcgen05.mma kind::tf32 accept for MN-major operands?"""
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
def report(name, D, want):
    D = D.cpu().numpy()
    eq = np.array_equal(D, want)
    frac = float(np.mean(D == want))
    print(f"{name:60s} exact={eq} frac_equal={frac:.3f} nan={int(np.isnan(D).sum())} zeros={int((D == 0).sum())} sample={D[0, :4]} want={want[0, :4]}")

A, W = q(128, 64), q(128, 64)
Z, F = q(128, 128), q(128, 64)
for mode, a, b, want in [(0, A, W, A @ W.T), (1, A, W, A @ W[64:]), (2, Z, F, Z.T @ F), (3, Z, q(128, 16), None)]:
    if want is None:
        b = q(128, 16); want = Z.T @ b
    D, st = ops.test_umma(mode, dev(a), dev(b))
    report(f"mode {mode} status={st}", D, want)
# dgrad B (MN-major) variants: swap LBO/SBO
for lbo, sbo, step, off in [(128, CH, 256, 1024), (CH, 128, 256, 1024)]:
    cfg = [64, 64, 64, idesc(128, 64, 0, 1), 4, 0, 2 * CH, CH, 128, off, step, lbo, sbo]
    D, st = ops.test_umma(1, dev(A), dev(W), cfg)
    report(f"dgrad B lbo={lbo} sbo={sbo} step={step} off={off} st={st}", D, A @ W[64:])
for a_lbo, a_sbo, b_lbo, b_sbo in [(128, CH, 128, CH), (CH, 128, CH, 128)]:
    cfg = [128, 64, 64, idesc(128, 64, 1, 1), 8, 0, 256, a_lbo, a_sbo, 0, 256, b_lbo, b_sbo]
    D, st = ops.test_umma(2, dev(Z), dev(F), cfg)
    report(f"wgrad a=({a_lbo},{a_sbo}) b=({b_lbo},{b_sbo}) st={st}", D, Z.T @ F)
