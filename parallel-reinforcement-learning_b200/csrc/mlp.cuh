// ActorCritic / RND network layout and the thread-per-row forward pieces (weights broadcast from shared
// memory, one row's 64 hidden activations in registers).
//
// Reference: /root/reference/PPO/ActorCritic.py:19-60 (trunk `model`, heads `actor` | `mu_head`+`log_std_head`,
// `critic`: Linear(no bias) -> GroupNorm(8, 64) -> SiLU -> Linear(+bias)), PPO/RND.py:25-31.
#pragma once
#include "common.cuh"

namespace prl {

// Offsets (in floats) into the flat parameter buffer = torch `.parameters()` order of the reference module.
struct HeadLayout {
    int w1, gw, gb, w2, b2, out;  // Linear(64,64,no bias), GN weight, GN bias, Linear(64,out) weight [out][64], bias
};
struct PolicyLayout {
    int O, A, cont, n_heads;  // heads: discrete {actor, critic}; continuous {mu, log_std, critic}
    int w0, g0w, g0b;         // trunk: Linear(O,64,no bias) weight [64][O], GN weight, GN bias
    HeadLayout head[3];
    int total;
    __host__ __device__ int critic() const { return n_heads - 1; }
};

__host__ __device__ inline PolicyLayout make_policy_layout(int cont, int O, int A) {
    PolicyLayout L;
    L.O = O; L.A = A; L.cont = cont; L.n_heads = cont ? 3 : 2;
    int off = 0;
    L.w0 = off; off += HID * O;
    L.g0w = off; off += HID;
    L.g0b = off; off += HID;
    for (int h = 0; h < L.n_heads; ++h) {
        const int out = (h == L.n_heads - 1) ? 1 : A;
        L.head[h].out = out;
        L.head[h].w1 = off; off += HID * HID;
        L.head[h].gw = off; off += HID;
        L.head[h].gb = off; off += HID;
        L.head[h].w2 = off; off += out * HID;
        L.head[h].b2 = off; off += out;
    }
    L.total = off;
    return L;
}

// RND net: Linear(I,64)+b -> GN -> SiLU -> Linear(64,Oo)+b ; parameters() order: 0.weight 0.bias 1.weight 1.bias 3.weight 3.bias
struct RndLayout {
    int I, Oo, w0, b0, gw, gb, w2, b2, total;
};
__host__ __device__ inline RndLayout make_rnd_layout(int I, int Oo) {
    RndLayout L;
    L.I = I; L.Oo = Oo;
    int off = 0;
    L.w0 = off; off += HID * I;
    L.b0 = off; off += HID;
    L.gw = off; off += HID;
    L.gb = off; off += HID;
    L.w2 = off; off += Oo * HID;
    L.b2 = off; off += Oo;
    L.total = off;
    return L;
}

#ifdef __CUDACC__
// ---- cooperative staging of weights into shared memory -------------------------------------------------------
// copy n floats global -> shared (whole block)
__device__ __forceinline__ void stage_copy(float *dst, const float *__restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
}
// stage a [rows][cols] row-major matrix transposed: dst[c*rows + r] = src[r*cols + c]
__device__ __forceinline__ void stage_transposed(float *dst, const float *__restrict__ src, int rows, int cols) {
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
        const int r = i / cols, c = i - r * cols;
        dst[c * rows + r] = __ldg(src + i);
    }
}

// ---- per-row math -----------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu(float y) { return y / (1.0f + expf(-y)); }

// GroupNorm(8 groups of 8) + affine + SiLU, in place on one row held in registers.
__device__ __forceinline__ void gn_silu(float (&z)[HID], const float *gw, const float *gb) {
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) m += z[g * GSIZE + i];
        m *= (1.0f / GSIZE);
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) {
            const float d = z[g * GSIZE + i] - m;
            v = fmaf(d, d, v);
        }
        const float rstd = 1.0f / sqrtf(v * (1.0f / GSIZE) + GN_EPS);
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) {
            const int j = g * GSIZE + i;
            z[j] = silu(fmaf((z[j] - m) * rstd, gw[j], gb[j]));
        }
    }
}

// acc[j] += x * Wt[j], Wt = 64 contiguous floats in shared memory (16 broadcast LDS.128)
__device__ __forceinline__ void axpy64(float x, const float *Wt, float (&acc)[HID]) {
    const float4 *w = reinterpret_cast<const float4 *>(Wt);
#pragma unroll
    for (int q = 0; q < HID / 4; ++q) {
        const float4 ww = w[q];
        acc[4 * q + 0] = fmaf(x, ww.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(x, ww.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(x, ww.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(x, ww.w, acc[4 * q + 3]);
    }
}

// dot(h[0..64), w[0..64)) with w in shared memory
__device__ __forceinline__ float dot64(const float (&h)[HID], const float *w) {
    const float4 *w4 = reinterpret_cast<const float4 *>(w);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int q = 0; q < HID / 4; ++q) {
        const float4 ww = w4[q];
        a0 = fmaf(h[4 * q + 0], ww.x, a0);
        a1 = fmaf(h[4 * q + 1], ww.y, a1);
        a2 = fmaf(h[4 * q + 2], ww.z, a2);
        a3 = fmaf(h[4 * q + 3], ww.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

// hidden layer of a head: z = GN_SiLU(W1 . f) with f read from this thread's shared-memory column
// (col[k * stride], k < 64) and W1 staged TRANSPOSED ([in][out]).
__device__ __forceinline__ void head_hidden(const float *col, int stride, const float *W1t, const float *gw,
                                            const float *gb, float (&z)[HID]) {
#pragma unroll
    for (int j = 0; j < HID; ++j) z[j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < HID; ++k) axpy64(col[k * stride], W1t + k * HID, z);
    gn_silu(z, gw, gb);
}

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

#endif  // __CUDACC__
}  // namespace prl
