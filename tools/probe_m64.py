"""GPU box: where does a tcgen05.mma with M = 64 (cta_group::1, kind::f16, operands from shared memory) put its 64 accumulator rows in
tensor memory, and what does it cost next to M = 128?  Uses the raw mode of the parity-test hook (libprl_b200_test.so): A[r][0] = r + 1,
B[n][0] = 1, so accumulator row r holds r + 1 in every column; all 128 lanes are read back.  Groundwork for a two-team (2 x 64-row
tiles in flight) layout of the update kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-reinforcement-learning_b200")]
import numpy as np, torch as t
from prl_b200 import ops
CH = 2048
def idesc(M, N, a_mn=0, b_mn=0):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)
A = np.zeros((128, 64), np.float32); A[:, 0] = np.arange(1, 129)
B = np.zeros((128, 64), np.float32); B[:, 0] = 1.0
Ad, Bd = t.from_numpy(A).cuda(), t.from_numpy(B).cuda()
for M, ws, dl in ((128, 1, 0), (64, 1, 0), (64, 1, 16), (64, -1, 0), (64, -1, 64), (32, -1, 0), (128, -1, 0)):
    cfg = [64, 64, 64, idesc(M, 64), 4, 0, 2 * CH, CH, 128, 0, 2 * CH, CH, 128, 0, 0, ws, dl]
    D, st = ops.test_umma(-1, Ad, Bd, cfg=cfg)
    D = D.cpu().numpy()
    print(f"M = {M}{' .ws' if ws < 0 else ''}, accumulator lane offset {dl}: status {st}")
    rows = {}
    for lane in range(128):
        v = D[lane]
        if np.all(v == v[0]) and 1 <= v[0] <= 128 and float(v[0]).is_integer():
            rows[lane] = int(v[0]) - 1
    runs, start = [], None
    for lane in range(129):
        ok = lane in rows and (start is None or rows[lane] == rows[lane - 1] + 1 if lane - 1 in rows else True)
        if lane in rows and start is None:
            start = lane
        if start is not None and (lane not in rows or (lane - 1 in rows and rows[lane] != rows[lane - 1] + 1 and lane != start)):
            runs.append((start, lane - 1, rows[start], rows[lane - 1])); start = lane if lane in rows else None
    print("   lanes -> accumulator rows:", ", ".join(f"lanes {a}-{b} = rows {ra}-{rb}" for a, b, ra, rb in runs) or "none recognised")
    for reps in (1, 512):
        cfg[15] = reps * ws
        _, st = ops.test_umma(-1, Ad, Bd, cfg=cfg)
        print(f"   {reps:4d} x 4 MMAs (N = 64, K = 16 each): {ops.test_umma.last_cycles} cycles from first issue to completion")
    c1 = None

print("\n.ws forms, all 128 lanes x 128 columns read back: where does accumulator row r (value r + 1) land?")
for M in (64, 32):
    cfg = [64, 64, 128, idesc(M, 64), 4, 0, 2 * CH, CH, 128, 0, 2 * CH, CH, 128, 0, 0, -1, 0]
    D, st = ops.test_umma(-1, Ad, Bd, cfg=cfg)
    D = D.cpu().numpy()
    for r in (0, 1, 15, 16, 31, 32, 33, 47, 48, 63)[: 10 if M == 64 else 5]:
        where = np.argwhere(D == float(r + 1))
        lanes = sorted(set(int(x) for x in where[:, 0])); cols = sorted(set(int(x) for x in where[:, 1]))
        print(f"   M = {M} .ws row {r:2d}: lanes {lanes[:4]}{'...' if len(lanes) > 4 else ''} ({len(lanes)}), columns {cols[0] if cols else None}..{cols[-1] if cols else None} ({len(cols)})")
