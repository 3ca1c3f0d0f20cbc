// Thread-per-row policy forward + sampling shared by prl_policy_act (generic O/A, rows from global memory) and
// the fused rollout kernel (env-specific O/A, observation in registers).
//
// Reference: PPO.get_action (/root/reference/PPO/PPO.py:82-96) = ActorCritic.get_dist (ActorCritic.py:85-110)
// + dist.sample() + tanh * action_scaling for continuous policies.
#pragma once
#include "mlp.cuh"

namespace prl {

constexpr float F32_EPS = 1.1920928955078125e-07f;  // torch.finfo(float32).eps (probs_to_logits clamp)
constexpr float LOG_2PI = 1.8378770664093453f;

// shared-memory image of the trunk and the first n_heads heads (acting needs the policy heads only; evaluation adds
// the critic), hidden-layer matrices staged TRANSPOSED ([in][out]) for the axpy-form forward.
struct ActSmem {
    const float *w0t, *g0w, *g0b;          // [O][64], [64], [64]
    const float *w1t[3], *gw[3], *gb[3];   // [64][64] transposed, GN affine
    const float *w2[3], *b2[3];            // [out][64], [out]
    int out[3], row[3];                    // output width of head h and its first scratch row (after the 64 hidden rows)
    float *scratch;                        // (64 + sum(out)) rows x blockDim.x floats, one column per thread
    int n_heads, n_out;
};

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }
// n_heads = 0 -> policy heads only (1 discrete / 2 continuous); otherwise the first n_heads heads of the layout
__host__ __device__ inline int act_heads(const PolicyLayout &L, bool with_critic) { return with_critic ? L.n_heads : L.n_heads - 1; }
__host__ __device__ inline size_t act_smem_floats(const PolicyLayout &L, int nthreads, bool with_critic = false) {
    const int nh = act_heads(L, with_critic);
    size_t w = (size_t)L.O * HID + 2 * HID;
    int outs = 0;
    for (int h = 0; h < nh; ++h) {
        w += HID * HID + 2 * HID + L.head[h].out * HID + round4(L.head[h].out);
        outs += L.head[h].out;
    }
    return w + (size_t)(HID + outs) * nthreads;
}

#ifdef __CUDACC__
__device__ __forceinline__ ActSmem stage_act_weights(float *smem, const float *__restrict__ params, const PolicyLayout &L,
                                                     bool with_critic = false) {
    ActSmem W;
    float *p = smem;
    float *w0t = p; p += L.O * HID;
    float *g0w = p; p += HID;
    float *g0b = p; p += HID;
    stage_transposed(w0t, params + L.w0, HID, L.O);
    stage_copy(g0w, params + L.g0w, HID);
    stage_copy(g0b, params + L.g0b, HID);
    W.w0t = w0t; W.g0w = g0w; W.g0b = g0b;
    W.n_heads = act_heads(L, with_critic);
    W.n_out = 0;
    for (int h = 0; h < W.n_heads; ++h) {
        const HeadLayout &H = L.head[h];
        float *w1t = p; p += HID * HID;
        float *gw = p; p += HID;
        float *gb = p; p += HID;
        float *w2 = p; p += H.out * HID;
        float *b2 = p; p += round4(H.out);
        stage_transposed(w1t, params + H.w1, HID, HID);
        stage_copy(gw, params + H.gw, HID);
        stage_copy(gb, params + H.gb, HID);
        stage_copy(w2, params + H.w2, H.out * HID);
        stage_copy(b2, params + H.b2, H.out);
        W.w1t[h] = w1t; W.gw[h] = gw; W.gb[h] = gb; W.w2[h] = w2; W.b2[h] = b2;
        W.out[h] = H.out; W.row[h] = HID + W.n_out;
        W.n_out += H.out;
    }
    W.scratch = p;
    return W;
}

// Runs trunk + heads for one row.  `xf(i)` returns input feature i.  Leaves head h's outputs in this thread's scratch
// column at rows W.row[h] .. W.row[h] + W.out[h] - 1 (discrete: logits, value; continuous: mu, log_std pre-activation, value).
template <typename XF>
__device__ __forceinline__ void policy_forward(const ActSmem &W, int O, XF xf, float *col, int stride) {
    float z[HID];
#pragma unroll
    for (int j = 0; j < HID; ++j) z[j] = 0.f;
    for (int i = 0; i < O; ++i) axpy64(xf(i), W.w0t + i * HID, z);
    gn_silu(z, W.g0w, W.g0b);
#pragma unroll
    for (int k = 0; k < HID; ++k) col[k * stride] = z[k];
    for (int h = 0; h < W.n_heads; ++h) {
        head_hidden(col, stride, W.w1t[h], W.gw[h], W.gb[h], z);
        for (int a = 0; a < W.out[h]; ++a) col[(W.row[h] + a) * stride] = W.b2[h][a] + dot64(z, W.w2[h] + a * HID);
    }
}

// softmax -> Categorical(probs) (renormalised) -> inverse-CDF sample with one uniform.  Overwrites the logits in the
// scratch column with the normalised probabilities; optionally copies them out.
__device__ __forceinline__ int sample_categorical(float *col, int stride, int A, float u, float *probs_out) {
    float *lg = col + HID * stride;
    float m = lg[0];
    for (int a = 1; a < A; ++a) m = fmaxf(m, lg[a * stride]);
    float S = 0.f;
    for (int a = 0; a < A; ++a) {
        const float e = expf(lg[a * stride] - m);
        lg[a * stride] = e;
        S += e;
    }
    float P = 0.f;  // torch Categorical(probs) divides by probs.sum(-1) once more
    for (int a = 0; a < A; ++a) {
        const float p = lg[a * stride] / S;
        lg[a * stride] = p;
        P += p;
    }
    float cum = 0.f;
    int pick = A - 1;
    bool found = false;
    for (int a = 0; a < A; ++a) {
        const float p = lg[a * stride] / P;
        if (probs_out) probs_out[a] = p;
        cum += p;
        if (!found && u < cum) { pick = a; found = true; }
    }
    return pick;
}

// standard normals from Philox via Box-Muller: fills n4 <= 4 values from one Philox block
__device__ __forceinline__ void normals4(const uint32_t r[4], float out[4]) {
    const float u0 = u01f(r[0]), u1 = u01f(r[1]), u2 = u01f(r[2]), u3 = u01f(r[3]);
    const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
    float s, c;
    sincospif(2.0f * u1, &s, &c);
    out[0] = ra * c; out[1] = ra * s;
    sincospif(2.0f * u3, &s, &c);
    out[2] = rb * c; out[3] = rb * s;
}

__device__ __forceinline__ uint32_t action_stream_word(int block) { return (STREAM_ACTION << 24) | (uint32_t)block; }
#endif

}  // namespace prl
