// Shared pieces of the fp32-FMA update kernels (csrc/update_ppo.cu, csrc/update_rnd.cu): tile constants, the shared-
// memory image of the networks, block-cooperative reductions over a tile's rows, GroupNorm forward / backward on a row
// held in registers, the per-head forward / backward, the fixed-order partial-gradient reduction.
#pragma once
#include "policy.cuh"

namespace prl {

// =============================================================================================== fused minibatch step
constexpr int UP_NT = 256;          // rows per tile = threads per block
constexpr int UP_NTP = UP_NT + 4;   // padded row stride of the [64][rows] staging arrays (keeps float4 alignment)

// shared-memory image for the update kernel: hidden matrices in torch layout [out][in]
struct UpSmem {
    float *w0t, *g0w, *g0b;           // [O][64] transposed, [64], [64]
    float *w1[3], *gw[3], *gb[3];     // [64 out][64 in]
    float *w2[3], *b2[3];             // [out][64], [out]
    float *F, *Z;                     // [64][UP_NTP] staging arrays: trunk activations / transient
    float *D;                         // [max(A,1)][UP_NTP] head-output gradients
    float *X;                         // [O][UP_NTP] inputs
    double *red;                      // 32 doubles
};

__host__ __device__ inline size_t up_smem_floats(const PolicyLayout &L) {
    size_t w = (size_t)L.O * HID + 2 * HID;
    for (int h = 0; h < L.n_heads; ++h) w += HID * HID + 2 * HID + L.head[h].out * HID + round4(L.head[h].out);
    return w + 2 * (size_t)HID * UP_NTP + (size_t)round4(L.A) * UP_NTP + (size_t)L.O * UP_NTP + 64 /* red */;
}

__device__ __forceinline__ UpSmem stage_update_weights(float *smem, const float *__restrict__ params, const PolicyLayout &L) {
    UpSmem W;
    float *p = smem;
    W.w0t = p; p += L.O * HID;
    W.g0w = p; p += HID;
    W.g0b = p; p += HID;
    stage_transposed(W.w0t, params + L.w0, HID, L.O);
    stage_copy(W.g0w, params + L.g0w, HID);
    stage_copy(W.g0b, params + L.g0b, HID);
    for (int h = 0; h < L.n_heads; ++h) {
        const HeadLayout &H = L.head[h];
        W.w1[h] = p; p += HID * HID;
        W.gw[h] = p; p += HID;
        W.gb[h] = p; p += HID;
        W.w2[h] = p; p += H.out * HID;
        W.b2[h] = p; p += round4(H.out);
        stage_copy(W.w1[h], params + H.w1, HID * HID);
        stage_copy(W.gw[h], params + H.gw, HID);
        stage_copy(W.gb[h], params + H.gb, HID);
        stage_copy(W.w2[h], params + H.w2, H.out * HID);
        stage_copy(W.b2[h], params + H.b2, H.out);
    }
    W.F = p; p += HID * UP_NTP;
    W.Z = p; p += HID * UP_NTP;
    W.D = p; p += round4(L.A) * UP_NTP;
    W.X = p; p += L.O * UP_NTP;
    W.red = reinterpret_cast<double *>(p);
    return W;
}

// ---- block-cooperative reductions over the tile's rows ----------------------------------------------------------
// part[j*64 + k] += sum_s Zr[j][s] * Fr[k][s]   (64 x 64 outputs, 4 x 4 per thread, float4 along s)
__device__ __forceinline__ void coop_outer64(const float *__restrict__ Zr, const float *__restrict__ Fr, float *__restrict__ part) {
    const int tj = threadIdx.x >> 4, tk = threadIdx.x & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[i][m] = 0.f;
#pragma unroll 2
    for (int s = 0; s < UP_NT; s += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(Zr + (tj + 16 * i) * UP_NTP + s);
#pragma unroll
        for (int m = 0; m < 4; ++m) b[m] = *reinterpret_cast<const float4 *>(Fr + (tk + 16 * m) * UP_NTP + s);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                acc[i][m] = fmaf(a[i].x, b[m].x, acc[i][m]);
                acc[i][m] = fmaf(a[i].y, b[m].y, acc[i][m]);
                acc[i][m] = fmaf(a[i].z, b[m].z, acc[i][m]);
                acc[i][m] = fmaf(a[i].w, b[m].w, acc[i][m]);
            }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) part[(tj + 16 * i) * HID + tk + 16 * m] += acc[i][m];
}

// part[r*nb + c] += sum_s A_[r][s] * B_[c][s]  for a small (na x nb) output; one output per thread (strided)
__device__ __forceinline__ void coop_outer_small(const float *__restrict__ A_, int na, const float *__restrict__ B_, int nb,
                                                 float *__restrict__ part) {
    for (int idx = threadIdx.x; idx < na * nb; idx += UP_NT) {
        const int r = idx / nb, c = idx - r * nb;
        const float4 *a = reinterpret_cast<const float4 *>(A_ + r * UP_NTP), *b = reinterpret_cast<const float4 *>(B_ + c * UP_NTP);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
        for (int s = 0; s < UP_NT / 4; ++s) {
            const float4 av = a[s], bv = b[s];
            s0 = fmaf(av.x, bv.x, s0); s1 = fmaf(av.y, bv.y, s1); s2 = fmaf(av.z, bv.z, s2); s3 = fmaf(av.w, bv.w, s3);
        }
        part[idx] += (s0 + s1) + (s2 + s3);
    }
}

// part[r] += sum_s A_[r][s]  for nr rows; 4 threads per row
__device__ __forceinline__ void coop_rowsum(const float *__restrict__ A_, int nr, float *__restrict__ part) {
    for (int base = 0; base < nr; base += UP_NT / 4) {
        const int r = base + (threadIdx.x >> 2), q = threadIdx.x & 3;
        float s0 = 0.f, s1 = 0.f;
        if (r < nr) {
            const float4 *a = reinterpret_cast<const float4 *>(A_ + r * UP_NTP + q * (UP_NT / 4));
#pragma unroll 4
            for (int s = 0; s < UP_NT / 16; ++s) {
                const float4 v = a[s];
                s0 += v.x + v.y; s1 += v.z + v.w;
            }
        }
        float t = s0 + s1;
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        if (r < nr && q == 0) part[r] += t;
    }
}

// GroupNorm forward on a row in registers: z -> zhat (in place), returns per-group rstd
__device__ __forceinline__ void gn_normalize(float (&z)[HID], float (&rstd)[GROUPS]) {
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) m += z[g * GSIZE + i];
        m *= (1.0f / GSIZE);
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) { const float d = z[g * GSIZE + i] - m; v = fmaf(d, d, v); }
        const float r = 1.0f / sqrtf(v * (1.0f / GSIZE) + GN_EPS);
        rstd[g] = r;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) z[g * GSIZE + i] = (z[g * GSIZE + i] - m) * r;
    }
}

// GroupNorm backward: dzhat (in d) and zhat -> dz (in d)
__device__ __forceinline__ void gn_backward(float (&d)[HID], const float (&zhat)[HID], const float (&rstd)[GROUPS]) {
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) { m1 += d[g * GSIZE + i]; m2 = fmaf(d[g * GSIZE + i], zhat[g * GSIZE + i], m2); }
        m1 *= (1.0f / GSIZE); m2 *= (1.0f / GSIZE);
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) d[g * GSIZE + i] = rstd[g] * (d[g * GSIZE + i] - m1 - zhat[g * GSIZE + i] * m2);
    }
}

// One head, one row: forward from F column, output-gradient callback, backward; accumulates this block's partial
// gradients (part) cooperatively and this row's df (gradient wrt trunk activations) in registers.
// `loss_grad(out, dout)` maps the head's outputs to their gradients (both arrays of H.out floats in registers/local).
struct HeadCtx {
    const float *w1, *gw, *gb, *w2, *b2;
    int out;
    float *p_w1, *p_gw, *p_gb, *p_w2, *p_b2;  // this block's partial-gradient rows
};

constexpr int MAX_OUT = 8;  // action_dim supported by the fused update kernel (register-resident head outputs)

// forward of one head for this thread's row: returns outputs, keeps zhat/rstd.
// NOT inlined (nor is head_backward_row) since round 2: the fp32-FMA update kernel calls them up to seven times (continuous policy:
// mu, log_std, mu again, critic) and with everything inlined it was 2.7 MB of SASS that took ptxas 6.3 minutes - most of a clean
// build().  As calls: 0.9 MB, 35 s; with dy evaluated once per element (head_backward_row) the path costs 807 us per 65 536-row
// launch against 730 us fully inlined, and since round 2 it
// is the second implementation the parity tests run next to the tensor-core path and the fallback for policies wider than
// observ_dim 16 / action_dim 8, not the path any BASELINE configuration takes.
static __device__ __noinline__ void head_forward_row(const HeadCtx &H, const float *Fcol, float *Zcol, float (&zhat)[HID],
                                                 float (&rstd)[GROUPS], float (&out)[MAX_OUT]) {
    // dot form: z[j] = <f, W1[j][:]> with f in registers; results parked in the Z column
    {
        float f[HID];
#pragma unroll
        for (int k = 0; k < HID; ++k) f[k] = Fcol[k * UP_NTP];
#pragma unroll 2
        for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = dot64(f, H.w1 + j * HID);
    }
#pragma unroll
    for (int j = 0; j < HID; ++j) zhat[j] = Zcol[j * UP_NTP];
    gn_normalize(zhat, rstd);
#pragma unroll
    for (int a = 0; a < MAX_OUT; ++a) out[a] = 0.f;
    for (int a = 0; a < H.out; ++a) {
        float acc0 = 0.f, acc1 = 0.f;
        const float4 *w4 = reinterpret_cast<const float4 *>(H.w2 + a * HID);
#pragma unroll
        for (int q = 0; q < HID / 4; ++q) {
            const float4 w = w4[q];
            acc0 = fmaf(silu(fmaf(zhat[4 * q + 0], H.gw[4 * q + 0], H.gb[4 * q + 0])), w.x, acc0);
            acc1 = fmaf(silu(fmaf(zhat[4 * q + 1], H.gw[4 * q + 1], H.gb[4 * q + 1])), w.y, acc1);
            acc0 = fmaf(silu(fmaf(zhat[4 * q + 2], H.gw[4 * q + 2], H.gb[4 * q + 2])), w.z, acc0);
            acc1 = fmaf(silu(fmaf(zhat[4 * q + 3], H.gw[4 * q + 3], H.gb[4 * q + 3])), w.w, acc1);
        }
        out[a] = H.b2[a] + (acc0 + acc1);
    }
}

// backward of one head.  On entry zhat/rstd hold the forward state of this row, dout its output gradients (zeros for
// padding rows).  Adds the row's contribution to df (registers) and the tile's contribution to the block partials.
// dy (dL/dy through Linear(64,out) and SiLU) is evaluated ONCE per element into a local array and read by the three staging passes
// (it used to be recomputed - expf, a division, the dot with dout - in each of them).
static __device__ __noinline__ void head_backward_row(const HeadCtx &H, const UpSmem &W, const float (&zhat)[HID],
                                                  const float (&rstd)[GROUPS], const float (&dout)[MAX_OUT], float (&df)[HID]) {
    float *Zcol = W.Z + threadIdx.x, *Dcol = W.D + threadIdx.x;
    float dy[HID];
    // (1) stage h and dout -> dW2, db2
#pragma unroll
    for (int j = 0; j < HID; ++j) {
        float dh = 0.f;
        for (int a = 0; a < H.out; ++a) dh = fmaf(dout[a], H.w2[a * HID + j], dh);
        const float y = fmaf(zhat[j], H.gw[j], H.gb[j]);
        const float sg = 1.0f / (1.0f + expf(-y));
        dy[j] = dh * sg * fmaf(y, 1.0f - sg, 1.0f);
        Zcol[j * UP_NTP] = silu(y);
    }
    for (int a = 0; a < H.out; ++a) Dcol[a * UP_NTP] = dout[a];
    __syncthreads();
    coop_outer_small(W.D, H.out, W.Z, HID, H.p_w2);
    coop_rowsum(W.D, H.out, H.p_b2);
    __syncthreads();
    // (2) stage dy * zhat -> dgamma
#pragma unroll
    for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = dy[j] * zhat[j];
    __syncthreads();
    coop_rowsum(W.Z, HID, H.p_gw);
    __syncthreads();
    // (3) stage dy -> dbeta
#pragma unroll
    for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = dy[j];
    __syncthreads();
    coop_rowsum(W.Z, HID, H.p_gb);
    __syncthreads();
    // (4) dz1 = GroupNorm backward of (dy * gamma), group by group; stage -> dW1 (with F), df += W1^T dz1
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
        float d[GSIZE], m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) {
            const int j = g * GSIZE + i;
            d[i] = dy[j] * H.gw[j];
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
    coop_outer64(W.Z, W.F, H.p_w1);
#pragma unroll 4
    for (int j = 0; j < HID; ++j) axpy64(Zcol[j * UP_NTP], H.w1 + j * HID, df);
    __syncthreads();
}

// grad[i] = sum over blocks of partials[b][i], fixed order; loss_out += block loss partials
static __global__ void k_reduce_partials(const float *__restrict__ partials, int nblocks, int P, float *__restrict__ grad,
                                  const double *__restrict__ loss_partials, double *__restrict__ loss_out, double rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) {
        float s = 0.f;
        for (int bl = 0; bl < nblocks; ++bl) s += partials[(size_t)bl * P + i];
        grad[i] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x < 3 && loss_out) {
        double s = 0.0;
        for (int bl = 0; bl < nblocks; ++bl) s += loss_partials[bl * 4 + threadIdx.x];
        loss_out[threadIdx.x] += s;
        if (threadIdx.x == 0) loss_out[3] += rows;
    }
}

static inline int update_grid(int64_t b) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ntiles = (b + UP_NT - 1) / UP_NT;
    return (int)(ntiles < sms ? (ntiles > 0 ? ntiles : 1) : sms);
}

}  // namespace prl
