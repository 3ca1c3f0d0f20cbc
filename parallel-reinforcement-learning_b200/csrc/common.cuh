// Shared helpers for the prl_b200 CUDA library: error reporting across the C ABI, Philox4x32-10, warp/block
// reductions.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/prl_b200.h"

namespace prl {

constexpr int HID = 64;     // hidden width of every MLP in the reference (PPO/ActorCritic.py:19-60)
constexpr int GROUPS = 8;   // GroupNorm(64 // 8, 64)
constexpr int GSIZE = HID / GROUPS;
constexpr float GN_EPS = 1e-5f;

void set_error(const char *fmt, ...);
int check_launch(const char *what);

#define PRL_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            prl::set_error(__VA_ARGS__);       \
            return PRL_ERR_INVALID;            \
        }                                      \
    } while (0)

#define PRL_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            prl::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return PRL_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------- Philox
// Philox4x32-10 (Salmon et al. 2011), counter-based: one call = 4 x 32 random bits for (key, counter).
struct Philox {
    uint32_t k0, k1;
    __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
        uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32);
        lo = (uint32_t)p;
    }
    __host__ __device__ inline void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// (0,1) float from 32 bits: (u + 0.5) * 2^-32 is never 0 or 1 after rounding to 24 bits? It can round to 1.0f,
// so use the top 24 bits instead: ((u >> 8) + 0.5) * 2^-24 in (0,1) exactly representable.
__host__ __device__ inline float u01f(uint32_t u) { return ((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// [0,1) double from 64 bits (53 significant), like numpy's random_double
__host__ __device__ inline double u01d(uint32_t hi, uint32_t lo) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

// Random streams: what a Philox counter word c2 selects.
enum : uint32_t { STREAM_RESET = 1, STREAM_ACTION = 2 };

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------- reductions
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result valid in thread 0.  `scratch` must hold 32 T's.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < nw ? scratch[lane] : T(0);
        v = warp_sum(v);
    }
    return v;
}
#endif

}  // namespace prl
