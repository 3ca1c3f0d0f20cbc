// PPO.learn numerics, forward-only and optimiser parts: ActorCritic.get_evaluate forward and clip_grad_norm_ + AdamW.
//
// Reference: /root/reference/PPO/ActorCritic.py:118-146 (get_evaluate), PPO/PPO.py:134-154 (old-policy pass),
// :250-252 (clip_grad_norm_(2.0), AdamW.step()).  The fused minibatch gradient kernels live in update_ppo.cu (fp32 FMA
// path) and update_rnd.cu.
#include "update_common.cuh"
#include "tiled_mlp.cuh"

namespace prl {

// =============================================================================================== evaluate (forward)
// register-tiled forward (tiled_mlp.cuh), persistent over 256-row tiles; thread t then finishes row t of the tile
__global__ void __launch_bounds__(EV_THREADS, 2)
k_policy_evaluate(const float *__restrict__ params, PolicyLayout L, const float *__restrict__ states,
                  const float *__restrict__ actions, int64_t n, float *__restrict__ logp_out, float *__restrict__ value_out,
                  double *__restrict__ entropy_sum) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double red[32];
    const EvSmem S = ev_layout(L, L.n_heads);
    const int tid = threadIdx.x, O = L.O, A = L.A;
    ev_stage_weights(smem, S, params, L);
    float *sX = smem + S.x;
    const float *sO = smem + S.out;
    const int64_t ntiles = (n + EV_ROWS - 1) / EV_ROWS;
    double ent_acc = 0.0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * EV_ROWS;
        const int rows = (int)min((int64_t)EV_ROWS, n - row0);
        __syncthreads();   // weights staged (first pass) / previous tile's readers of sX and sO done
        for (int i = tid; i < EV_ROWS * O; i += EV_THREADS) sX[i] = i < rows * O ? __ldg(states + row0 * O + i) : 0.f;
        __syncthreads();
        ev_forward_tile(smem, S, L);
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
    const size_t smem = (size_t)ev_layout(L, L.n_heads).total * sizeof(float);
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
