"""TEST INFRASTRUCTURE (oracle) - integer restatement of the random stream a seeded gymnasium env draws its reset state from.

gymnasium (third-party, pinned gymnasium==1.1.1 in /root/reference/requirements.txt:4, absent here) implements
``env.reset(seed=s)`` as ``np.random.Generator(np.random.PCG64(np.random.SeedSequence(s)))`` followed by
``np_random.uniform(low, high, size)`` (the reference's call site: /root/reference/AsyncTools/AsyncPPO.py:53, per env).
Everything below that line is NUMPY, which IS installed here, so this restatement is pinned against the real
implementation (tests/test_oracle.py::test_numpy_seeded_reset_stream_restatement): SeedSequence's entropy pool and
``generate_state``, PCG64's seeding / 128-bit LCG step / XSL-RR output, and ``Generator.uniform`` = low + (high - low) * u,
u = (next_uint64 >> 11) * 2^-53.  csrc/np_rng.cuh is the device form of these functions.
"""
from __future__ import annotations

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF
M128 = (1 << 128) - 1
INIT_A, MULT_A, INIT_B, MULT_B = 0x43B0D7E5, 0x931E8875, 0x8B51F9DD, 0x58F38DED
MIX_L, MIX_R, XSHIFT, POOL = 0xCA01F9DD, 0x4973F715, 16, 4
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645


def seed_sequence_pool(seed: int) -> list[int]:
    """SeedSequence(seed).pool for 0 <= seed < 2**128 (missing entropy words mix in as zeros)."""
    words = [(seed >> (32 * i)) & M32 for i in range(POOL)]
    assert seed >> 128 == 0
    hc = [INIT_A]

    def hashmix(v):
        v ^= hc[0]
        hc[0] = (hc[0] * MULT_A) & M32
        v = (v * hc[0]) & M32
        return v ^ (v >> XSHIFT)

    def mix(x, y):
        r = (MIX_L * x - MIX_R * y) & M32
        return r ^ (r >> XSHIFT)

    pool = [hashmix(w) for w in words]
    for s in range(POOL):
        for d in range(POOL):
            if s != d:
                pool[d] = mix(pool[d], hashmix(pool[s]))
    return pool


def generate_state_u64(pool: list[int], n: int) -> list[int]:
    """SeedSequence.generate_state(n, np.uint64): 2n 32-bit words, paired little-endian."""
    hc = INIT_B
    out = []
    for i in range(2 * n):
        v = pool[i % POOL] ^ hc
        hc = (hc * MULT_B) & M32
        v = (v * hc) & M32
        out.append(v ^ (v >> XSHIFT))
    return [out[2 * k] | (out[2 * k + 1] << 32) for k in range(n)]


def pcg64_seed(seed: int) -> tuple[int, int]:
    """(state, inc) of PCG64(SeedSequence(seed)) as 128-bit integers."""
    v = generate_state_u64(seed_sequence_pool(seed), 4)
    initstate, initseq = (v[0] << 64) | v[1], (v[2] << 64) | v[3]
    inc = ((initseq << 1) | 1) & M128
    state = inc                      # (0 * MULT + inc)
    state = (state + initstate) & M128
    state = (state * PCG_MULT + inc) & M128
    return state, inc


def pcg64_next(state: int, inc: int) -> tuple[int, int]:
    """-> (new state, 64-bit output): step, then XSL-RR of the NEW state."""
    state = (state * PCG_MULT + inc) & M128
    hi, lo = state >> 64, state & M64
    x, rot = hi ^ lo, hi >> 58
    return state, ((x >> rot) | (x << ((64 - rot) & 63))) & M64


def uniform(state: int, inc: int, low: float, high: float) -> tuple[int, float]:
    """Generator.uniform(low, high) for one value: low + (high - low) * next_double."""
    state, r = pcg64_next(state, inc)
    return state, low + (high - low) * ((r >> 11) * (1.0 / 9007199254740992.0))
