// PPO.learn numerics, forward-only and optimiser parts: ActorCritic.get_evaluate forward and clip_grad_norm_ + AdamW.
//
// Reference: /root/reference/PPO/ActorCritic.py:118-146 (get_evaluate), PPO/PPO.py:134-154 (old-policy pass),
// :250-252 (clip_grad_norm_(2.0), AdamW.step()).  The fused minibatch gradient kernels live in update_ppo.cu (fp32 FMA
// path) and update_rnd.cu.
#include "update_common.cuh"

namespace prl {

// =============================================================================================== evaluate (forward)
// Register-tiled fp32 forward: one CTA = 256 threads = one tile of 256 rows at a time, persistent over tiles.  Thread
// (rg, fg) = (tid >> 3, tid & 7) owns rows 8 rg .. 8 rg + 7 and hidden features 8 fg .. 8 fg + 7 (exactly one GroupNorm
// group, so the normalisation needs no exchange): 64 accumulators, each k-step feeds 64 FMAs from 2 + 8/4 128-bit
// shared-memory loads.  The trunk output F lives in shared memory as [row][64]; a head's hidden layer stays in
// registers and its small output layer is a per-thread partial dot + a butterfly over the 8 feature lanes.
//
// Feature j is kept at "physical" position phys(j) in every shared-memory vector so that the 8 feature lanes of a
// quarter-warp read / write 128 contiguous bytes (no bank conflicts): thread fg's features sit at [4 fg, 4 fg + 4) and
// [32 + 4 fg, 32 + 4 fg + 4).
constexpr int EV_THREADS = 256, EV_ROWS = 256, EV_RPT = 8;
__host__ __device__ __forceinline__ int ev_phys(int j) { return ((j & 4) ? 32 : 0) + 4 * (j >> 3) + (j & 3); }

struct EvSmem {
    int x, f, w0, g0w, g0b, w1[3], gw[3], gb[3], w2[3], b2[3], col[3], out, n_out, so, total;   // offsets in floats
};
__host__ __device__ inline EvSmem ev_layout(const PolicyLayout &L) {
    EvSmem S;
    int off = 0;
    S.f = off; off += EV_ROWS * HID;
    S.x = off; off += round4(EV_ROWS * L.O);
    S.w0 = off; off += L.O * HID;
    S.g0w = off; off += HID;
    S.g0b = off; off += HID;
    S.n_out = 0;
    for (int h = 0; h < L.n_heads; ++h) {
        S.w1[h] = off; off += HID * HID;
        S.gw[h] = off; off += HID;
        S.gb[h] = off; off += HID;
        S.w2[h] = off; off += L.head[h].out * HID;
        S.b2[h] = off; off += round4(L.head[h].out);
        S.col[h] = S.n_out; S.n_out += L.head[h].out;
    }
    S.so = S.n_out | 1;   // odd row stride of the output scratch
    S.out = off; off += round4(EV_ROWS * S.so);
    S.total = off;
    return S;
}

// GroupNorm (this thread's 8 features = one group) + affine + SiLU on 8 rows; same operation order as gn_silu (mlp.cuh)
__device__ __forceinline__ void ev_gn_silu(float (&z)[EV_RPT][8], const float *gw_p, const float *gb_p, int fg) {
    const float4 ga = *reinterpret_cast<const float4 *>(gw_p + 4 * fg), gb_ = *reinterpret_cast<const float4 *>(gw_p + 32 + 4 * fg);
    const float4 ba = *reinterpret_cast<const float4 *>(gb_p + 4 * fg), bb = *reinterpret_cast<const float4 *>(gb_p + 32 + 4 * fg);
    const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb_.x, gb_.y, gb_.z, gb_.w}, bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int r = 0; r < EV_RPT; ++r) {
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) m += z[r][i];
        m *= (1.0f / GSIZE);
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = z[r][i] - m; v = fmaf(d, d, v); }
        // IEEE sqrt and divisions, as in gn_silu: the post-update weight parity (AdamW's normalised steps amplify 1-ulp
        // differences in old_logp / values) does not survive the approximate reciprocal forms - measured
        const float rstd = 1.0f / sqrtf(v * (1.0f / GSIZE) + GN_EPS);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[r][i] = silu(fmaf((z[r][i] - m) * rstd, g[i], bt[i]));
    }
}

__global__ void __launch_bounds__(EV_THREADS, 2)
k_policy_evaluate(const float *__restrict__ params, PolicyLayout L, const float *__restrict__ states,
                  const float *__restrict__ actions, int64_t n, float *__restrict__ logp_out, float *__restrict__ value_out,
                  double *__restrict__ entropy_sum) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double red[32];
    const EvSmem S = ev_layout(L);
    const int tid = threadIdx.x, fg = tid & 7, rg = tid >> 3, O = L.O, A = L.A;
    // ---- weights -> shared memory, features permuted to their physical positions (rows of W1 = contraction index too)
    for (int i = tid; i < HID * O; i += EV_THREADS) { const int j = i / O, o = i - j * O; smem[S.w0 + o * HID + ev_phys(j)] = __ldg(params + L.w0 + i); }
    for (int i = tid; i < HID; i += EV_THREADS) {
        smem[S.g0w + ev_phys(i)] = __ldg(params + L.g0w + i);
        smem[S.g0b + ev_phys(i)] = __ldg(params + L.g0b + i);
    }
    for (int h = 0; h < L.n_heads; ++h) {
        const HeadLayout &H = L.head[h];
        for (int i = tid; i < HID * HID; i += EV_THREADS) { const int j = i >> 6, k = i & 63; smem[S.w1[h] + ev_phys(k) * HID + ev_phys(j)] = __ldg(params + H.w1 + i); }
        for (int i = tid; i < HID; i += EV_THREADS) {
            smem[S.gw[h] + ev_phys(i)] = __ldg(params + H.gw + i);
            smem[S.gb[h] + ev_phys(i)] = __ldg(params + H.gb + i);
        }
        for (int i = tid; i < H.out * HID; i += EV_THREADS) { const int a = i >> 6, j = i & 63; smem[S.w2[h] + a * HID + ev_phys(j)] = __ldg(params + H.w2 + i); }
        for (int i = tid; i < H.out; i += EV_THREADS) smem[S.b2[h] + i] = __ldg(params + H.b2 + i);
    }
    float *sF = smem + S.f, *sX = smem + S.x, *sO = smem + S.out;
    const int64_t ntiles = (n + EV_ROWS - 1) / EV_ROWS;
    double ent_acc = 0.0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * EV_ROWS;
        const int rows = (int)min((int64_t)EV_ROWS, n - row0);
        __syncthreads();   // weights staged (first pass) / previous tile's readers of sX and sO done
        for (int i = tid; i < EV_ROWS * O; i += EV_THREADS) sX[i] = i < rows * O ? __ldg(states + row0 * O + i) : 0.f;
        __syncthreads();
        float z[EV_RPT][8];
        // ---- trunk: Linear(O, 64, no bias) -> GN -> SiLU -> sF
#pragma unroll
        for (int r = 0; r < EV_RPT; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) z[r][i] = 0.f;
        for (int o = 0; o < O; ++o) {
            const float4 wa = *reinterpret_cast<const float4 *>(smem + S.w0 + o * HID + 4 * fg);
            const float4 wb = *reinterpret_cast<const float4 *>(smem + S.w0 + o * HID + 32 + 4 * fg);
#pragma unroll
            for (int r = 0; r < EV_RPT; ++r) {
                const float xv = sX[(EV_RPT * rg + r) * O + o];
                z[r][0] = fmaf(xv, wa.x, z[r][0]); z[r][1] = fmaf(xv, wa.y, z[r][1]); z[r][2] = fmaf(xv, wa.z, z[r][2]); z[r][3] = fmaf(xv, wa.w, z[r][3]);
                z[r][4] = fmaf(xv, wb.x, z[r][4]); z[r][5] = fmaf(xv, wb.y, z[r][5]); z[r][6] = fmaf(xv, wb.z, z[r][6]); z[r][7] = fmaf(xv, wb.w, z[r][7]);
            }
        }
        ev_gn_silu(z, smem + S.g0w, smem + S.g0b, fg);
#pragma unroll
        for (int r = 0; r < EV_RPT; ++r) {
            float *dst = sF + (EV_RPT * rg + r) * HID + 4 * fg;
            *reinterpret_cast<float4 *>(dst) = make_float4(z[r][0], z[r][1], z[r][2], z[r][3]);
            *reinterpret_cast<float4 *>(dst + 32) = make_float4(z[r][4], z[r][5], z[r][6], z[r][7]);
        }
        __syncthreads();
        // ---- heads: Linear(64, 64, no bias) -> GN -> SiLU -> Linear(64, out) + bias
        for (int h = 0; h < L.n_heads; ++h) {
#pragma unroll
            for (int r = 0; r < EV_RPT; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) z[r][i] = 0.f;
            const float *Wp = smem + S.w1[h] + 4 * fg;
            const float *Fp = sF + (EV_RPT * rg) * HID;
#pragma unroll 2
            for (int k4 = 0; k4 < HID / 4; ++k4) {
                float4 a[EV_RPT];
#pragma unroll
                for (int r = 0; r < EV_RPT; ++r) a[r] = *reinterpret_cast<const float4 *>(Fp + r * HID + 4 * k4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 wa = *reinterpret_cast<const float4 *>(Wp + (4 * k4 + kk) * HID);
                    const float4 wb = *reinterpret_cast<const float4 *>(Wp + (4 * k4 + kk) * HID + 32);
#pragma unroll
                    for (int r = 0; r < EV_RPT; ++r) {
                        const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
                        z[r][0] = fmaf(av, wa.x, z[r][0]); z[r][1] = fmaf(av, wa.y, z[r][1]); z[r][2] = fmaf(av, wa.z, z[r][2]); z[r][3] = fmaf(av, wa.w, z[r][3]);
                        z[r][4] = fmaf(av, wb.x, z[r][4]); z[r][5] = fmaf(av, wb.y, z[r][5]); z[r][6] = fmaf(av, wb.z, z[r][6]); z[r][7] = fmaf(av, wb.w, z[r][7]);
                    }
                }
            }
            ev_gn_silu(z, smem + S.gw[h], smem + S.gb[h], fg);
            const int outs = L.head[h].out;
            for (int a = 0; a < outs; ++a) {
                const float4 wa = *reinterpret_cast<const float4 *>(smem + S.w2[h] + a * HID + 4 * fg);
                const float4 wb = *reinterpret_cast<const float4 *>(smem + S.w2[h] + a * HID + 32 + 4 * fg);
                float mine = 0.f;
#pragma unroll
                for (int r = 0; r < EV_RPT; ++r) {
                    float p = ((z[r][0] * wa.x + z[r][1] * wa.y) + (z[r][2] * wa.z + z[r][3] * wa.w)) +
                              ((z[r][4] * wb.x + z[r][5] * wb.y) + (z[r][6] * wb.z + z[r][7] * wb.w));
                    p += __shfl_xor_sync(0xffffffffu, p, 1);
                    p += __shfl_xor_sync(0xffffffffu, p, 2);
                    p += __shfl_xor_sync(0xffffffffu, p, 4);
                    mine = fg == r ? p : mine;
                }
                sO[(EV_RPT * rg + fg) * S.so + S.col[h] + a] = mine + smem[S.b2[h] + a];   // lane fg keeps row 8 rg + fg
            }
        }
        __syncthreads();
        // ---- per-row epilogue: thread t finishes row t of the tile (coalesced stores)
        if (tid < rows) {
            const int64_t row = row0 + tid;
            const float *o = sO + tid * S.so;
            float logp, ent = 0.f;
            if (!L.cont) {
                float m = o[0];
                for (int a = 1; a < A; ++a) m = fmaxf(m, o[a]);
                float e[16], Ssum = 0.f;   // A <= 16 enforced by the entry point for the discrete case
                for (int a = 0; a < A; ++a) { e[a & 15] = expf(o[a] - m); Ssum += e[a & 15]; }
                float P = 0.f;
                for (int a = 0; a < A; ++a) { e[a & 15] = e[a & 15] / Ssum; P += e[a & 15]; }
                const int act = (int)actions[row];
                logp = 0.f;
                for (int a = 0; a < A; ++a) {
                    const float p = e[a & 15] / P;
                    const float l = logf(fminf(fmaxf(p, F32_EPS), 1.0f - F32_EPS));
                    ent -= l * p;
                    if (a == act) logp = l;
                }
            } else {
                float q = 0.f, hld = 0.f;
                for (int a = 0; a < A; ++a) {
                    const float mu = o[S.col[0] + a], ls = o[S.col[1] + a];
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
            value_out[row] = o[S.col[L.n_heads - 1]];
            ent_acc += (double)ent;
        }
    }
    const double bs = block_sum<double>(ent_acc, red);
    if (tid == 0 && entropy_sum) atomicAdd(entropy_sum, bs);
}

// =============================================================================================== clip + AdamW
__global__ void __launch_bounds__(1024)
k_adamw(float *__restrict__ params, const float *__restrict__ grad, float *__restrict__ m, float *__restrict__ v, int64_t n,
        int64_t step, int64_t *__restrict__ step_dev, float lr, float wd, float max_norm, double *__restrict__ norm_out) {
    __shared__ double red[32];
    __shared__ float coef_s, step_size_s, bc2_sqrt_s;
    // device-resident optimiser clock {step, beta1^step, beta2^step}: the launch is replayable (CUDA graphs) and the bias
    // corrections cost two multiplications instead of two double-precision pow()
    double *pows = reinterpret_cast<double *>(step_dev) + 1;
    if (step_dev) step = *step_dev + 1;
    // gradient norm: float loads, double accumulation (9 027 values for CartPole: ~9 per thread)
    double ss = 0.0;
    const int64_t n4 = n >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(grad);
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 g = g4[i];
        ss += (double)g.x * g.x + (double)g.y * g.y + (double)g.z * g.z + (double)g.w * g.w;
    }
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) ss += (double)grad[i] * grad[i];
    ss = block_sum<double>(ss, red);
    if (threadIdx.x == 0) {
        const float total = (float)sqrt(ss);
        coef_s = max_norm > 0.f ? fminf(max_norm / (total + 1e-6f), 1.0f) : 1.f;
        if (norm_out) *norm_out = (double)total;
        // bias corrections once per launch (double pow is expensive; every thread used to evaluate it)
        double p1, p2;
        if (step_dev && step > 1 && pows[0] > 0.0) {
            p1 = pows[0] * 0.9; p2 = pows[1] * 0.999;
        } else {
            p1 = pow(0.9, (double)step); p2 = pow(0.999, (double)step);
        }
        step_size_s = (float)((double)lr / (1.0 - p1));
        bc2_sqrt_s = (float)sqrt(1.0 - p2);
        if (step_dev) { *step_dev = step; pows[0] = p1; pows[1] = p2; }
    }
    __syncthreads();
    const float coef = coef_s, step_size = step_size_s, bc2_sqrt = bc2_sqrt_s;
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
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

}  // namespace prl

using namespace prl;

extern "C" {

int prl_policy_evaluate(const float *params, int is_continuous, int obs_dim, int action_dim, const float *states,
                        const float *actions, int64_t n, float *logp, float *value, double *entropy_sum, void *stream) {
    PRL_REQUIRE(params && obs_dim > 0 && action_dim > 0 && n >= 0, "prl_policy_evaluate: bad arguments");
    if (n == 0) return PRL_OK;
    PRL_REQUIRE(states && actions && logp && value, "prl_policy_evaluate: null pointer");
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    PRL_REQUIRE(is_continuous || action_dim <= 16, "prl_policy_evaluate: discrete action_dim=%d > 16", action_dim);
    const size_t smem = (size_t)ev_layout(L).total * sizeof(float);
    PRL_REQUIRE(smem <= 227 * 1024, "prl_policy_evaluate: observ_dim=%d action_dim=%d needs %zu B shared memory", obs_dim, action_dim, smem);
    PRL_CUDA(cudaFuncSetAttribute(k_policy_evaluate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ntiles = cdiv(n, EV_ROWS);
    const int grid = (int)(ntiles < 2 * sms ? ntiles : 2 * sms);   // two resident CTAs per SM, persistent over tiles
    k_policy_evaluate<<<grid, EV_THREADS, smem, (cudaStream_t)stream>>>(params, L, states, actions, n, logp, value, entropy_sum);
    return check_launch("k_policy_evaluate");
}

int prl_adamw_step(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, int64_t step, float lr,
                   float weight_decay, float max_norm, double *grad_norm_out, void *stream) {
    PRL_REQUIRE(params && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "prl_adamw_step: bad arguments");
    k_adamw<<<1, 1024, 0, (cudaStream_t)stream>>>(params, grad, exp_avg, exp_avg_sq, n, step, nullptr, lr, weight_decay, max_norm, grad_norm_out);
    return check_launch("k_adamw");
}

int prl_adamw_step_dev(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, int64_t *step_counter, float lr,
                       float weight_decay, float max_norm, double *grad_norm_out, void *stream) {
    PRL_REQUIRE(params && grad && exp_avg && exp_avg_sq && n > 0 && step_counter, "prl_adamw_step_dev: bad arguments");
    k_adamw<<<1, 1024, 0, (cudaStream_t)stream>>>(params, grad, exp_avg, exp_avg_sq, n, 0, step_counter, lr, weight_decay, max_norm, grad_norm_out);
    return check_launch("k_adamw");
}

}  // extern "C"
