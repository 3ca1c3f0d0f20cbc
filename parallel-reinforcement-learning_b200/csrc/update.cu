// PPO.learn numerics: ActorCritic.get_evaluate forward, the fused clipped-surrogate minibatch step
// (forward + loss + backward -> flat gradient), and clip_grad_norm_ + AdamW.
//
// Reference: /root/reference/PPO/ActorCritic.py:118-146 (get_evaluate), PPO/PPO.py:134-154 (old-policy pass),
// :219-252 (k_epochs x minibatch loop: ratio clamp +-20, clipped surrogate, 0.5*SmoothL1(mean), detached entropy,
// loss.mean().backward(), clip_grad_norm_(2.0), AdamW.step()).
//
// Kernel shape (fp32 FMA parity path): one minibatch row per thread for everything that is per-row (GEMV against
// weights broadcast from shared memory, GroupNorm, SiLU, loss), and block-cooperative reductions over the tile's
// rows for everything that sums over rows (weight / affine gradients), staged through two [64][rows] shared arrays.
// Blocks are persistent; each accumulates its tiles into its own row of a global partial-gradient workspace, and a
// second kernel adds the rows in a fixed order -> the gradient is bit-reproducible run to run.
#include "policy.cuh"

namespace prl {

constexpr float LOG_2PI = 1.8378770664093453f;

// =============================================================================================== evaluate (forward)
constexpr int EV_TPB = 128;

__global__ void __launch_bounds__(EV_TPB)
k_policy_evaluate(const float *__restrict__ params, PolicyLayout L, const float *__restrict__ states,
                  const float *__restrict__ actions, int64_t n, float *__restrict__ logp_out, float *__restrict__ value_out,
                  double *__restrict__ entropy_sum) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double red[32];
    const ActSmem W = stage_act_weights(smem, params, L, /*with_critic=*/true);
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * EV_TPB + threadIdx.x;
    float ent = 0.f;
    if (row < n) {
        float *col = W.scratch + threadIdx.x;
        const float *x = states + row * L.O;
        policy_forward(W, L.O, [&](int i) { return __ldg(x + i); }, col, EV_TPB);
        const int A = L.A;
        float logp;
        if (!L.cont) {
            float *lg = col + W.row[0] * EV_TPB;
            float m = lg[0];
            for (int a = 1; a < A; ++a) m = fmaxf(m, lg[a * EV_TPB]);
            float S = 0.f;
            for (int a = 0; a < A; ++a) { const float e = expf(lg[a * EV_TPB] - m); lg[a * EV_TPB] = e; S += e; }
            float P = 0.f;
            for (int a = 0; a < A; ++a) { const float p = lg[a * EV_TPB] / S; lg[a * EV_TPB] = p; P += p; }
            const int act = (int)actions[row];
            logp = 0.f;
            for (int a = 0; a < A; ++a) {
                const float p = lg[a * EV_TPB] / P;
                const float l = logf(fminf(fmaxf(p, F32_EPS), 1.0f - F32_EPS));
                ent -= l * p;
                if (a == act) logp = l;
            }
        } else {
            float q = 0.f, hld = 0.f;
            for (int a = 0; a < A; ++a) {
                const float mu = col[(W.row[0] + a) * EV_TPB];
                const float ls = col[(W.row[1] + a) * EV_TPB];
                const float sd = softplus_t(fminf(fmaxf(ls, -2.f), 2.f));
                const float tril = sqrtf(sd * sd);
                const float zt = (actions[row * A + a] - mu) / tril;
                q = fmaf(zt, zt, q);
                hld += logf(tril);
            }
            logp = -0.5f * (A * LOG_2PI + q) - hld;
            ent = 0.5f * A * (1.0f + LOG_2PI) + hld;
        }
        logp_out[row] = logp;
        value_out[row] = col[W.row[L.n_heads - 1] * EV_TPB];
    }
    const double bs = block_sum<double>((double)ent, red);
    if (threadIdx.x == 0 && entropy_sum) atomicAdd(entropy_sum, bs);
}

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

// forward of one head for this thread's row: returns outputs, keeps zhat/rstd
__device__ __forceinline__ void head_forward_row(const HeadCtx &H, const float *Fcol, float *Zcol, float (&zhat)[HID],
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
// Register budget: only zhat[64] and df[64] stay live; dy is recomputed from (zhat, dout) in each staging pass.
__device__ __forceinline__ void head_backward_row(const HeadCtx &H, const UpSmem &W, const float (&zhat)[HID],
                                                  const float (&rstd)[GROUPS], const float (&dout)[MAX_OUT], float (&df)[HID]) {
    float *Zcol = W.Z + threadIdx.x, *Dcol = W.D + threadIdx.x;
    auto dy_of = [&](int j) -> float {   // dL/dy_j through Linear(64,out) and SiLU
        float dh = 0.f;
        for (int a = 0; a < H.out; ++a) dh = fmaf(dout[a], H.w2[a * HID + j], dh);
        const float y = fmaf(zhat[j], H.gw[j], H.gb[j]);
        const float sg = 1.0f / (1.0f + expf(-y));
        return dh * sg * fmaf(y, 1.0f - sg, 1.0f);
    };
    // (1) stage h and dout -> dW2, db2
#pragma unroll
    for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = silu(fmaf(zhat[j], H.gw[j], H.gb[j]));
    for (int a = 0; a < H.out; ++a) Dcol[a * UP_NTP] = dout[a];
    __syncthreads();
    coop_outer_small(W.D, H.out, W.Z, HID, H.p_w2);
    coop_rowsum(W.D, H.out, H.p_b2);
    __syncthreads();
    // (2) stage dy * zhat -> dgamma
#pragma unroll
    for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = dy_of(j) * zhat[j];
    __syncthreads();
    coop_rowsum(W.Z, HID, H.p_gw);
    __syncthreads();
    // (3) stage dy -> dbeta
#pragma unroll
    for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = dy_of(j);
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
            d[i] = dy_of(j) * H.gw[j];
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

__global__ void __launch_bounds__(UP_NT, 1)
k_ppo_grad(const float *__restrict__ params, PolicyLayout L, const float *__restrict__ states, const float *__restrict__ actions,
           const float *__restrict__ old_logp, const float *__restrict__ adv, const float *__restrict__ returns, int64_t b,
           float clip, float inv_count, float *__restrict__ partials, double *__restrict__ loss_partials) {
    extern __shared__ __align__(16) float smem[];
    UpSmem W = stage_update_weights(smem, params, L);
    const int P = L.total;
    float *part = partials + (size_t)blockIdx.x * P;
    for (int i = threadIdx.x; i < P; i += UP_NT) part[i] = 0.f;
    __syncthreads();
    const int A = L.A, O = L.O;
    const int64_t ntiles = (b + UP_NT - 1) / UP_NT;
    double l_pol = 0.0, l_val = 0.0, l_ent = 0.0;
    float *Fcol = W.F + threadIdx.x, *Zcol = W.Z + threadIdx.x, *Xcol = W.X + threadIdx.x;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row = tile * UP_NT + threadIdx.x;
        const bool live = row < b;
        // ---- trunk forward
        float zhat[HID], rstd[GROUPS];
        {
            for (int i = 0; i < O; ++i) Xcol[i * UP_NTP] = live ? __ldg(states + row * O + i) : 0.f;
#pragma unroll
            for (int j = 0; j < HID; ++j) zhat[j] = 0.f;
            for (int i = 0; i < O; ++i) axpy64(Xcol[i * UP_NTP], W.w0t + i * HID, zhat);
            gn_normalize(zhat, rstd);
#pragma unroll
            for (int j = 0; j < HID; ++j) Fcol[j * UP_NTP] = silu(fmaf(zhat[j], W.g0w[j], W.g0b[j]));
        }
        float df[HID];
#pragma unroll
        for (int j = 0; j < HID; ++j) df[j] = 0.f;
        float out[MAX_OUT], dout[MAX_OUT];
        HeadCtx H;
        auto ctx = [&](int h) {
            const HeadLayout &hl = L.head[h];
            H.w1 = W.w1[h]; H.gw = W.gw[h]; H.gb = W.gb[h]; H.w2 = W.w2[h]; H.b2 = W.b2[h]; H.out = hl.out;
            H.p_w1 = part + hl.w1; H.p_gw = part + hl.gw; H.p_gb = part + hl.gb; H.p_w2 = part + hl.w2; H.p_b2 = part + hl.b2;
        };
        const float adv_i = live ? adv[row] : 0.f, old_i = live ? old_logp[row] : 0.f;
        // dL/dlogp of -min(r*A, clamp(r)*A) * inv_count, and the row's policy loss term
        auto surrogate = [&](float logp, float &pol_term) -> float {
            const float dl = logp - old_i;
            const float dlc = fminf(fmaxf(dl, -20.f), 20.f);
            const float r = expf(dlc);
            const float s1 = r * adv_i;
            const float rc = fminf(fmaxf(r, 1.0f - clip), 1.0f + clip);
            const float s2 = rc * adv_i;
            pol_term = -fminf(s1, s2);
            const float g1 = s1 < s2 ? 1.f : (s1 > s2 ? 0.f : 0.5f);  // torch.min splits ties evenly
            const float g2 = 1.f - g1;
            const float in_clip = (r >= 1.0f - clip && r <= 1.0f + clip) ? 1.f : 0.f;
            const float dr = -inv_count * adv_i * (g1 + g2 * in_clip);
            const float in20 = (dl >= -20.f && dl <= 20.f) ? 1.f : 0.f;
            return dr * r * in20;
        };
        if (!L.cont) {
            // ---- actor head
            ctx(0);
            head_forward_row(H, Fcol, Zcol, zhat, rstd, out);
#pragma unroll
            for (int a = 0; a < MAX_OUT; ++a) dout[a] = 0.f;
            if (live) {
                float m = out[0];
                for (int a = 1; a < A; ++a) m = fmaxf(m, out[a]);
                float S = 0.f, p[MAX_OUT];
                for (int a = 0; a < A; ++a) { p[a] = expf(out[a] - m); S += p[a]; }
                float Psum = 0.f;
                for (int a = 0; a < A; ++a) { p[a] = p[a] / S; Psum += p[a]; }
                const int act = (int)actions[row];
                float pa = 0.f, ent = 0.f;
                for (int a = 0; a < A; ++a) {
                    p[a] = p[a] / Psum;
                    const float l = logf(fminf(fmaxf(p[a], F32_EPS), 1.0f - F32_EPS));
                    ent -= l * p[a];
                    if (a == act) pa = p[a];
                }
                const float logp = logf(fminf(fmaxf(pa, F32_EPS), 1.0f - F32_EPS));
                float pol;
                float dlogp = surrogate(logp, pol);
                if (!(pa >= F32_EPS && pa <= 1.0f - F32_EPS)) dlogp = 0.f;  // clamp in probs_to_logits blocks the gradient
                for (int a = 0; a < A; ++a) dout[a] = dlogp * ((a == act ? 1.f : 0.f) - p[a]);
                l_pol += pol; l_ent += ent;
            }
            head_backward_row(H, W, zhat, rstd, dout, df);
        } else {
            // ---- mu head forward only (outputs kept), then log_std head fwd+bwd, then mu head fwd(recompute)+bwd
            float mu[MAX_OUT], dmu[MAX_OUT];
            ctx(0);
            head_forward_row(H, Fcol, Zcol, zhat, rstd, mu);
            ctx(1);
            head_forward_row(H, Fcol, Zcol, zhat, rstd, out);
#pragma unroll
            for (int a = 0; a < MAX_OUT; ++a) { dout[a] = 0.f; dmu[a] = 0.f; }
            if (live) {
                float q = 0.f, hld = 0.f, zt[MAX_OUT], tril[MAX_OUT];
                for (int a = 0; a < A; ++a) {
                    const float lc = fminf(fmaxf(out[a], -2.f), 2.f);
                    const float sd = softplus_t(lc);
                    tril[a] = sqrtf(sd * sd);
                    zt[a] = (actions[row * A + a] - mu[a]) / tril[a];
                    q = fmaf(zt[a], zt[a], q);
                    hld += logf(tril[a]);
                }
                const float logp = -0.5f * (A * LOG_2PI + q) - hld;
                float pol;
                const float dlogp = surrogate(logp, pol);
                for (int a = 0; a < A; ++a) {
                    dmu[a] = dlogp * zt[a] / tril[a];
                    const float lc = fminf(fmaxf(out[a], -2.f), 2.f);
                    const float in2 = (out[a] >= -2.f && out[a] <= 2.f) ? 1.f : 0.f;
                    const float dsd = dlogp * (zt[a] * zt[a] - 1.0f) / tril[a];
                    dout[a] = dsd * (1.0f / (1.0f + expf(-lc))) * in2;
                }
                l_pol += pol; l_ent += 0.5f * A * (1.0f + LOG_2PI) + hld;
            }
            head_backward_row(H, W, zhat, rstd, dout, df);
            ctx(0);
            head_forward_row(H, Fcol, Zcol, zhat, rstd, mu);
            head_backward_row(H, W, zhat, rstd, dmu, df);
        }
        // ---- critic head
        {
            ctx(L.n_heads - 1);
            head_forward_row(H, Fcol, Zcol, zhat, rstd, out);
#pragma unroll
            for (int a = 0; a < MAX_OUT; ++a) dout[a] = 0.f;
            if (live) {
                const float dv = out[0] - returns[row];
                const float ad = fabsf(dv);
                l_val += ad < 1.f ? 0.5f * dv * dv : ad - 0.5f;
                dout[0] = 0.5f * inv_count * (ad < 1.f ? dv : (dv > 0.f ? 1.f : -1.f));
            }
            head_backward_row(H, W, zhat, rstd, dout, df);
        }
        // ---- trunk backward (recompute the trunk's normalised pre-activations from the staged inputs)
        {
#pragma unroll
            for (int j = 0; j < HID; ++j) zhat[j] = 0.f;
            for (int i = 0; i < O; ++i) axpy64(Xcol[i * UP_NTP], W.w0t + i * HID, zhat);
            gn_normalize(zhat, rstd);
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                const float y = fmaf(zhat[j], W.g0w[j], W.g0b[j]);
                const float sg = 1.0f / (1.0f + expf(-y));
                df[j] = df[j] * sg * fmaf(y, 1.0f - sg, 1.0f);   // dy
                Zcol[j * UP_NTP] = df[j] * zhat[j];
            }
            __syncthreads();
            coop_rowsum(W.Z, HID, part + L.g0w);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = df[j];
            __syncthreads();
            coop_rowsum(W.Z, HID, part + L.g0b);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < HID; ++j) df[j] *= W.g0w[j];
            gn_backward(df, zhat, rstd);
#pragma unroll
            for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = df[j];
            __syncthreads();
            coop_outer_small(W.Z, HID, W.X, O, part + L.w0);   // dW0[j][i] = sum_s dz0[j][s] x[i][s]
            __syncthreads();
        }
    }
    // ---- loss partials (reporting only)
    const double bp = block_sum<double>(l_pol, W.red);
    const double bv = block_sum<double>(l_val, W.red);
    const double be = block_sum<double>(l_ent, W.red);
    if (threadIdx.x == 0) {
        loss_partials[blockIdx.x * 4 + 0] = bp;
        loss_partials[blockIdx.x * 4 + 1] = bv;
        loss_partials[blockIdx.x * 4 + 2] = be;
        loss_partials[blockIdx.x * 4 + 3] = 0.0;
    }
}

// grad[i] = sum over blocks of partials[b][i], fixed order; loss_out += block loss partials
__global__ void k_reduce_partials(const float *__restrict__ partials, int nblocks, int P, float *__restrict__ grad,
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

// =============================================================================================== clip + AdamW
__global__ void __launch_bounds__(1024)
k_adamw(float *__restrict__ params, const float *__restrict__ grad, float *__restrict__ m, float *__restrict__ v, int64_t n,
        int64_t step, float lr, float wd, float max_norm, double *__restrict__ norm_out) {
    __shared__ double red[32];
    __shared__ float coef_s;
    double ss = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const double g = grad[i]; ss += g * g; }
    ss = block_sum<double>(ss, red);
    if (threadIdx.x == 0) {
        const float total = (float)sqrt(ss);
        float c = 1.f;
        if (max_norm > 0.f) c = fminf(max_norm / (total + 1e-6f), 1.0f);
        coef_s = c;
        if (norm_out) *norm_out = (double)total;
    }
    __syncthreads();
    const float coef = coef_s;
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const double bc1 = 1.0 - pow(0.9, (double)step), bc2 = 1.0 - pow(0.999, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float decay = 1.0f - lr * wd;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float g = grad[i] * coef;
        float p = params[i] * decay;
        const float mi = m[i] + (1.0f - b1) * (g - m[i]);          // lerp
        const float vi = fmaf(v[i], b2, (1.0f - b2) * g * g);
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p = p - step_size * (mi / denom);
        params[i] = p; m[i] = mi; v[i] = vi;
    }
}

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
        float zhat[HID], rstd[GROUPS];
        rnd_hidden_pre(Pn, L.I, [&](int i) { return Xcol[i * UP_NTP]; }, zhat);
        gn_normalize(zhat, rstd);
        for (int o = 0; o < L.Oo; ++o) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < HID; ++j) acc = fmaf(silu(fmaf(zhat[j], Pn.gw[j], Pn.gb[j])), Pn.w2[o * HID + j], acc);
            const float d = live ? (Pn.b2[o] + acc) - tgt[o] : 0.f;
            sq += (double)d * d;
            dout[o] = 2.0f * d * inv_count;
        }
        auto dy_of = [&](int j) -> float {
            float dh = 0.f;
            for (int o = 0; o < L.Oo; ++o) dh = fmaf(dout[o], Pn.w2[o * HID + j], dh);
            const float y = fmaf(zhat[j], Pn.gw[j], Pn.gb[j]);
            const float sg = 1.0f / (1.0f + expf(-y));
            return dh * sg * fmaf(y, 1.0f - sg, 1.0f);
        };
#pragma unroll
        for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = silu(fmaf(zhat[j], Pn.gw[j], Pn.gb[j]));
        for (int o = 0; o < L.Oo; ++o) Dcol[o * UP_NTP] = dout[o];
        __syncthreads();
        coop_outer_small(D, L.Oo, Z, HID, part + L.w2);
        coop_rowsum(D, L.Oo, part + L.b2);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = dy_of(j) * zhat[j];
        __syncthreads();
        coop_rowsum(Z, HID, part + L.gw);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < HID; ++j) Zcol[j * UP_NTP] = dy_of(j);
        __syncthreads();
        coop_rowsum(Z, HID, part + L.gb);
        __syncthreads();
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
            float d[GSIZE], m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int i = 0; i < GSIZE; ++i) {
                const int j = g * GSIZE + i;
                d[i] = dy_of(j) * Pn.gw[j];
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

static int update_grid(int64_t b) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ntiles = (b + UP_NT - 1) / UP_NT;
    return (int)(ntiles < sms ? (ntiles > 0 ? ntiles : 1) : sms);
}

}  // namespace prl

using namespace prl;

extern "C" {

int prl_policy_evaluate(const float *params, int is_continuous, int obs_dim, int action_dim, const float *states,
                        const float *actions, int64_t n, float *logp, float *value, double *entropy_sum, void *stream) {
    PRL_REQUIRE(params && obs_dim > 0 && action_dim > 0 && n >= 0, "prl_policy_evaluate: bad arguments");
    if (n == 0) return PRL_OK;
    PRL_REQUIRE(states && actions && logp && value, "prl_policy_evaluate: null pointer");
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    const size_t smem = act_smem_floats(L, EV_TPB, true) * sizeof(float);
    PRL_REQUIRE(smem <= 227 * 1024, "prl_policy_evaluate: observ_dim=%d action_dim=%d needs %zu B shared memory", obs_dim, action_dim, smem);
    PRL_CUDA(cudaFuncSetAttribute(k_policy_evaluate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_policy_evaluate<<<cdiv(n, EV_TPB), EV_TPB, smem, (cudaStream_t)stream>>>(params, L, states, actions, n, logp, value, entropy_sum);
    return check_launch("k_policy_evaluate");
}

size_t prl_update_ws_floats(int is_continuous, int obs_dim, int action_dim, int64_t batch) {
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    const int grid = update_grid(batch);
    return (size_t)grid * L.total + (size_t)grid * 8 /* loss partials as doubles */ + 8;
}

int prl_ppo_grad(const float *params, int is_continuous, int obs_dim, int action_dim, const float *states, const float *actions,
                 const float *old_logp, const float *adv, const float *returns, int64_t b, float policy_clip, float inv_count,
                 float *grad, double *loss_out, float *ws, size_t ws_floats, void *stream) {
    PRL_REQUIRE(params && states && actions && old_logp && adv && returns && grad && ws && b > 0, "prl_ppo_grad: bad arguments");
    PRL_REQUIRE(action_dim <= MAX_OUT, "prl_ppo_grad: action_dim=%d > %d not supported by the fused update kernel", action_dim, MAX_OUT);
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    const int grid = update_grid(b);
    PRL_REQUIRE(ws_floats >= (size_t)grid * L.total + (size_t)grid * 8 + 8, "prl_ppo_grad: workspace too small");
    const size_t smem = up_smem_floats(L) * sizeof(float);
    PRL_REQUIRE(smem <= 227 * 1024, "prl_ppo_grad: observ_dim=%d needs %zu B shared memory (> 227 KB)", obs_dim, smem);
    cudaStream_t st = (cudaStream_t)stream;
    float *partials = ws;
    double *loss_partials = reinterpret_cast<double *>(ws + (((size_t)grid * L.total + 1) & ~(size_t)1));
    PRL_CUDA(cudaFuncSetAttribute(k_ppo_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ppo_grad<<<grid, UP_NT, smem, st>>>(params, L, states, actions, old_logp, adv, returns, b, policy_clip, inv_count, partials, loss_partials);
    k_reduce_partials<<<cdiv(L.total, 256), 256, 0, st>>>(partials, grid, L.total, grad, loss_partials, loss_out, (double)b);
    return check_launch("k_ppo_grad");
}

int prl_adamw_step(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, int64_t step, float lr,
                   float weight_decay, float max_norm, double *grad_norm_out, void *stream) {
    PRL_REQUIRE(params && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "prl_adamw_step: bad arguments");
    k_adamw<<<1, 1024, 0, (cudaStream_t)stream>>>(params, grad, exp_avg, exp_avg_sq, n, step, lr, weight_decay, max_norm, grad_norm_out);
    return check_launch("k_adamw");
}

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
