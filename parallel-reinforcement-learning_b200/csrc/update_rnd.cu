// RND predictor / target networks (reference: /root/reference/PPO/RND.py:71-115): intrinsic reward and the
// predictor's MSE gradient, fp32 FMA path, same tile machinery as update_ppo.cu.
#include "update_common.cuh"

namespace prl {

// =============================================================================================== RND (PPO/RND.py:71-115)
// net(x) = Linear(64,Oo)( SiLU( GroupNorm( Linear(I,64)(x) ) ) ), both Linear layers with bias.
struct RndSmem {
    float *w0t, *b0, *gw, *gb, *w2, *b2;  // [I][64] transposed, [64], [64], [64], [Oo][64], [Oo]
};
__host__ __device__ inline size_t rnd_net_floats(const RndLayout &L) { return (size_t)L.I * HID + 3 * HID + L.Oo * HID + round4(L.Oo); }

__device__ __forceinline__ float *stage_rnd(float *p, const float *__restrict__ params, const RndLayout &L, RndSmem &W) {
    W.w0t = p; p += L.I * HID;
    W.b0 = p; p += HID;
    W.gw = p; p += HID;
    W.gb = p; p += HID;
    W.w2 = p; p += L.Oo * HID;
    W.b2 = p; p += round4(L.Oo);
    stage_transposed(W.w0t, params + L.w0, HID, L.I);
    stage_copy(W.b0, params + L.b0, HID);
    stage_copy(W.gw, params + L.gw, HID);
    stage_copy(W.gb, params + L.gb, HID);
    stage_copy(W.w2, params + L.w2, L.Oo * HID);
    stage_copy(W.b2, params + L.b2, L.Oo);
    return p;
}

template <typename XF>
__device__ __forceinline__ void rnd_hidden_pre(const RndSmem &W, int I, XF xf, float (&z)[HID]) {
#pragma unroll
    for (int j = 0; j < HID; ++j) z[j] = W.b0[j];
    for (int i = 0; i < I; ++i) axpy64(xf(i), W.w0t + i * HID, z);
}

__global__ void __launch_bounds__(128)
k_rnd_intrinsic(const float *__restrict__ tparams, const float *__restrict__ pparams, RndLayout L, const float *__restrict__ states,
                int64_t n, float beta, const float *__restrict__ add_to, float *__restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    RndSmem T, Pn;
    float *p = stage_rnd(smem, tparams, L, T);
    stage_rnd(p, pparams, L, Pn);
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const float *x = states + row * L.I;
    float ht[HID], hp[HID];
    rnd_hidden_pre(T, L.I, [&](int i) { return __ldg(x + i); }, ht);
    gn_silu(ht, T.gw, T.gb);
    rnd_hidden_pre(Pn, L.I, [&](int i) { return __ldg(x + i); }, hp);
    gn_silu(hp, Pn.gw, Pn.gb);
    float ss = 0.f;
    for (int o = 0; o < L.Oo; ++o) {
        const float d = (Pn.b2[o] + dot64(hp, Pn.w2 + o * HID)) - (T.b2[o] + dot64(ht, T.w2 + o * HID));
        ss = fmaf(d, d, ss);
    }
    const float r = sqrtf(ss) * beta;
    out[row] = add_to ? add_to[row] + r : r;
}

__global__ void __launch_bounds__(UP_NT, 1)
k_rnd_grad(const float *__restrict__ tparams, const float *__restrict__ pparams, RndLayout L, const float *__restrict__ states,
           int64_t n, float inv_count, float *__restrict__ partials, double *__restrict__ loss_partials) {
    extern __shared__ __align__(16) float smem[];
    RndSmem T, Pn;
    float *p = stage_rnd(smem, tparams, L, T);
    p = stage_rnd(p, pparams, L, Pn);
    float *Z = p; p += HID * UP_NTP;
    float *D = p; p += round4(L.Oo) * UP_NTP;
    float *X = p; p += L.I * UP_NTP;
    double *red = reinterpret_cast<double *>(p);
    const int P = L.total;
    float *part = partials + (size_t)blockIdx.x * P;
    for (int i = threadIdx.x; i < P; i += UP_NT) part[i] = 0.f;
    __syncthreads();
    float *Zcol = Z + threadIdx.x, *Dcol = D + threadIdx.x, *Xcol = X + threadIdx.x;
    const int64_t ntiles = (n + UP_NT - 1) / UP_NT;
    double sq = 0.0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row = tile * UP_NT + threadIdx.x;
        const bool live = row < n;
        for (int i = 0; i < L.I; ++i) Xcol[i * UP_NTP] = live ? __ldg(states + row * L.I + i) : 0.f;
        float tgt[MAX_OUT], dout[MAX_OUT];
        {
            float ht[HID];
            rnd_hidden_pre(T, L.I, [&](int i) { return Xcol[i * UP_NTP]; }, ht);
            gn_silu(ht, T.gw, T.gb);
            for (int o = 0; o < L.Oo; ++o) tgt[o] = T.b2[o] + dot64(ht, T.w2 + o * HID);
        }
        // predictor forward, everything the backward needs kept in registers (one pass over the hidden features: the SiLU and its
        // derivative are evaluated ONCE per element - this kernel used to re-evaluate expf + division three more times per element
        // for the three staging passes below): zhat, then dy[j] = SiLU'(y_j) until the output gradient is known
        float zhat[HID], dy[HID], rstd[GROUPS];
        rnd_hidden_pre(Pn, L.I, [&](int i) { return Xcol[i * UP_NTP]; }, zhat);
        gn_normalize(zhat, rstd);
        float acc[MAX_OUT];
#pragma unroll
        for (int o = 0; o < MAX_OUT; ++o) acc[o] = 0.f;
#pragma unroll
        for (int j = 0; j < HID; ++j) {
            const float y = fmaf(zhat[j], Pn.gw[j], Pn.gb[j]);
            const float h = silu(y);
            const float sg = 1.0f / (1.0f + expf(-y));
            dy[j] = sg * fmaf(y, 1.0f - sg, 1.0f);
            Zcol[j * UP_NTP] = h;
#pragma unroll
            for (int o = 0; o < MAX_OUT; ++o)
                if (o < L.Oo) acc[o] = fmaf(h, Pn.w2[o * HID + j], acc[o]);
        }
#pragma unroll
        for (int o = 0; o < MAX_OUT; ++o) {
            dout[o] = 0.f;
            if (o < L.Oo) {
                const float d = live ? (Pn.b2[o] + acc[o]) - tgt[o] : 0.f;
                sq += (double)d * d;
                dout[o] = 2.0f * d * inv_count;
                Dcol[o * UP_NTP] = dout[o];
            }
        }
#pragma unroll
        for (int j = 0; j < HID; ++j) {
            float dh = 0.f;
#pragma unroll
            for (int o = 0; o < MAX_OUT; ++o)
                if (o < L.Oo) dh = fmaf(dout[o], Pn.w2[o * HID + j], dh);
            dy[j] *= dh;
        }
        __syncthreads();
        coop_outer_small(D, L.Oo, Z, HID, part + L.w2);
        coop_rowsum(D, L.Oo, part + L.b2);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = dy[j] * zhat[j];
        __syncthreads();
        coop_rowsum(Z, HID, part + L.gw);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = dy[j];
        __syncthreads();
        coop_rowsum(Z, HID, part + L.gb);
        __syncthreads();
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
            float d[GSIZE], m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int i = 0; i < GSIZE; ++i) {
                const int j = g * GSIZE + i;
                d[i] = dy[j] * Pn.gw[j];
                m1 += d[i];
                m2 = fmaf(d[i], zhat[j], m2);
            }
            m1 *= (1.0f / GSIZE); m2 *= (1.0f / GSIZE);
#pragma unroll
            for (int i = 0; i < GSIZE; ++i) {
                const int j = g * GSIZE + i;
                Zcol[j * UP_NTP] = rstd[g] * (d[i] - m1 - zhat[j] * m2);
            }
        }
        __syncthreads();
        coop_outer_small(Z, HID, X, L.I, part + L.w0);
        coop_rowsum(Z, HID, part + L.b0);
        __syncthreads();
    }
    const double bsq = block_sum<double>(sq, red);
    if (threadIdx.x == 0) {
        loss_partials[blockIdx.x * 4 + 0] = bsq;
        loss_partials[blockIdx.x * 4 + 1] = 0.0;
        loss_partials[blockIdx.x * 4 + 2] = 0.0;
        loss_partials[blockIdx.x * 4 + 3] = 0.0;
    }
}

}  // namespace prl

using namespace prl;

extern "C" {

int prl_rnd_intrinsic(const float *target_params, const float *pred_params, int in_features, int out_features, const float *states,
                      int64_t n, float beta, const float *add_to, float *out, void *stream) {
    PRL_REQUIRE(target_params && pred_params && in_features > 0 && out_features > 0 && n >= 0, "prl_rnd_intrinsic: bad arguments");
    if (n == 0) return PRL_OK;
    const RndLayout L = make_rnd_layout(in_features, out_features);
    const size_t smem = 2 * rnd_net_floats(L) * sizeof(float);
    PRL_REQUIRE(smem <= 227 * 1024, "prl_rnd_intrinsic: in_features=%d needs %zu B shared memory", in_features, smem);
    PRL_CUDA(cudaFuncSetAttribute(k_rnd_intrinsic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_rnd_intrinsic<<<cdiv(n, 128), 128, smem, (cudaStream_t)stream>>>(target_params, pred_params, L, states, n, beta, add_to, out);
    return check_launch("k_rnd_intrinsic");
}

int prl_rnd_grad(const float *target_params, const float *pred_params, int in_features, int out_features, const float *states,
                 int64_t n, float *grad, double *loss_out, float *ws, size_t ws_floats, void *stream) {
    PRL_REQUIRE(target_params && pred_params && states && grad && ws && n > 0, "prl_rnd_grad: bad arguments");
    PRL_REQUIRE(out_features <= MAX_OUT, "prl_rnd_grad: out_features=%d > %d not supported", out_features, MAX_OUT);
    const RndLayout L = make_rnd_layout(in_features, out_features);
    const int grid = update_grid(n);
    PRL_REQUIRE(ws_floats >= (size_t)grid * L.total + (size_t)grid * 8 + 8, "prl_rnd_grad: workspace too small");
    const size_t smem = (2 * rnd_net_floats(L) + (size_t)HID * UP_NTP + (size_t)round4(L.Oo) * UP_NTP + (size_t)L.I * UP_NTP + 64) * sizeof(float);
    PRL_REQUIRE(smem <= 227 * 1024, "prl_rnd_grad: in_features=%d needs %zu B shared memory", in_features, smem);
    cudaStream_t st = (cudaStream_t)stream;
    float *partials = ws;
    double *loss_partials = reinterpret_cast<double *>(ws + (((size_t)grid * L.total + 1) & ~(size_t)1));
    PRL_CUDA(cudaFuncSetAttribute(k_rnd_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_rnd_grad<<<grid, UP_NT, smem, st>>>(target_params, pred_params, L, states, n, 1.0f / ((float)n * (float)out_features), partials, loss_partials);
    k_reduce_partials<<<cdiv(L.total, 256), 256, 0, st>>>(partials, grid, L.total, grad, loss_partials, loss_out, (double)n);
    return check_launch("k_rnd_grad");
}

}  // extern "C"
