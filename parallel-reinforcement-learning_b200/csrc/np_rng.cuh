// numpy's seeded random stream on the device: SeedSequence(seed) -> PCG64 -> Generator.uniform, bit for bit.
//
// Why: a gymnasium env seeded with env.reset(seed=s) (the per-env call behind /root/reference/AsyncTools/AsyncPPO.py:53)
// draws its start state from np.random.Generator(np.random.PCG64(np.random.SeedSequence(s))).uniform(low, high, size);
// with this generator a seeded EnvVectorizer starts from the same fp64 states as E seeded reference envs, and every
// later reset() continues each env's stream like the reference's per-env generators do.
// Integer restatement pinned against numpy itself: oracle/np_rng.py.
//
// State in HBM: four uint64 planes [4][E] = {state_hi, state_lo, inc_hi, inc_lo} (coalesced across envs).
#pragma once
#include <stdint.h>

namespace prl {

struct Pcg64 {
    uint64_t shi, slo, ihi, ilo;

    // state = state * 0x2360ED051FC65DA4'4385DF649FCCF645 + inc  (mod 2^128)
    __host__ __device__ inline void step() {
        constexpr uint64_t MH = 0x2360ED051FC65DA4ull, ML = 0x4385DF649FCCF645ull;
#ifdef __CUDA_ARCH__
        const uint64_t carry = __umul64hi(slo, ML);
#else
        const uint64_t carry = (uint64_t)(((unsigned __int128)slo * ML) >> 64);
#endif
        uint64_t hi = carry + slo * MH + shi * ML;
        uint64_t lo = slo * ML;
        lo += ilo;
        hi += ihi + (lo < ilo ? 1ull : 0ull);
        shi = hi; slo = lo;
    }
    // pcg64_next64: step, then XSL-RR of the new state
    __host__ __device__ inline uint64_t next64() {
        step();
        const uint64_t x = shi ^ slo;
        const unsigned rot = (unsigned)(shi >> 58);
        return (x >> rot) | (x << ((64u - rot) & 63u));
    }
    // random_standard_uniform: 53 bits in [0, 1)
    __host__ __device__ inline double next_double() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }

    // PCG64(SeedSequence(seed)) for an integer seed below 2^64 (entropy words beyond the seed's are zeros, which is
    // what SeedSequence mixes in for a short entropy array)
    __host__ __device__ static inline Pcg64 from_seed(uint64_t seed) {
        constexpr uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu,
                           MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;
        uint32_t hc = INIT_A;
        auto hashmix = [&](uint32_t v) -> uint32_t {
            v ^= hc;
            hc *= MULT_A;
            v *= hc;
            return v ^ (v >> 16);
        };
        uint32_t pool[4] = {hashmix((uint32_t)seed), hashmix((uint32_t)(seed >> 32)), hashmix(0u), hashmix(0u)};
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int d = 0; d < 4; ++d)
                if (s != d) {
                    const uint32_t y = hashmix(pool[s]);
                    const uint32_t m = MIX_L * pool[d] - MIX_R * y;
                    pool[d] = m ^ (m >> 16);
                }
        // generate_state(4, uint64): eight 32-bit words, paired little-endian
        uint32_t w[8];
        uint32_t hb = INIT_B;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t v = pool[i & 3] ^ hb;
            hb *= MULT_B;
            v *= hb;
            w[i] = v ^ (v >> 16);
        }
        const uint64_t v0 = w[0] | ((uint64_t)w[1] << 32), v1 = w[2] | ((uint64_t)w[3] << 32);
        const uint64_t v2 = w[4] | ((uint64_t)w[5] << 32), v3 = w[6] | ((uint64_t)w[7] << 32);
        // pcg_setseq_128_srandom_r: initstate = {hi v0, lo v1}, initseq = {hi v2, lo v3}
        Pcg64 g;
        g.ihi = (v2 << 1) | (v3 >> 63);
        g.ilo = (v3 << 1) | 1ull;
        g.shi = 0; g.slo = 0;
        g.step();
        g.slo += v1;
        g.shi += v0 + (g.slo < v1 ? 1ull : 0ull);
        g.step();
        return g;
    }
};

}  // namespace prl
