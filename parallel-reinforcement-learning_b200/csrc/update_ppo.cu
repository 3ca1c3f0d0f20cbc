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
#include "update_common.cuh"

namespace prl {

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

}  // namespace prl

using namespace prl;

extern "C" {

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

}  // extern "C"
