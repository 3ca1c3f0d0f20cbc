// Register-tiled fp32 forward of the ActorCritic MLP, shared by the old-policy evaluation (update.cu), PPO.get_action and
// the fused rollout (rollout.cu) - the three produce bit-identical network outputs for the same row and weights.
//
// Reference: /root/reference/PPO/ActorCritic.py:19-60 (trunk + heads), :85-110 (get_dist), :118-146 (get_evaluate).
//
// One CTA = 256 threads = one tile of 256 rows.  Thread (rg, fg) = (tid >> 3, tid & 7) owns rows 8 rg .. 8 rg + 7 and
// hidden features 8 fg .. 8 fg + 7 (exactly one GroupNorm group, so the normalisation needs no exchange): 64
// accumulators, each k-step feeds 64 FMAs from 2 + 8/4 128-bit shared-memory loads.  The trunk output F lives in
// shared memory as [row][64]; a head's hidden layer stays in registers and its small output layer is a per-thread
// partial dot + a butterfly over the 8 feature lanes (commutative, so every lane and every row slot gets the same bits).
//
// Feature j is kept at "physical" position ev_phys(j) in every shared-memory vector so that the 8 feature lanes of a
// quarter-warp read / write 128 contiguous bytes (no bank conflicts): thread fg's features sit at [4 fg, 4 fg + 4) and
// [32 + 4 fg, 32 + 4 fg + 4).
#pragma once
#include "policy.cuh"

namespace prl {

constexpr int EV_THREADS = 256, EV_ROWS = 256, EV_RPT = 8;
__host__ __device__ __forceinline__ int ev_phys(int j) { return ((j & 4) ? 32 : 0) + 4 * (j >> 3) + (j & 3); }

struct EvSmem {
    int x, f, w0, g0w, g0b, w1[3], gw[3], gb[3], w2[3], b2[3], col[3], out, n_out, so, n_heads, total;   // offsets in floats
};
// n_heads: how many heads of the layout take part (all of them for get_evaluate; the policy heads only for get_action)
__host__ __device__ inline EvSmem ev_layout(const PolicyLayout &L, int n_heads) {
    EvSmem S;
    int off = 0;
    S.n_heads = n_heads;
    S.f = off; off += EV_ROWS * HID;
    S.x = off; off += round4(EV_ROWS * L.O);
    S.w0 = off; off += L.O * HID;
    S.g0w = off; off += HID;
    S.g0b = off; off += HID;
    S.n_out = 0;
    for (int h = 0; h < n_heads; ++h) {
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

#ifdef __CUDACC__
// weights -> shared memory, features permuted to their physical positions (rows of W1 = contraction index too).
// Whole CTA; the caller synchronises.
__device__ __forceinline__ void ev_stage_weights(float *smem, const EvSmem &S, const float *__restrict__ params, const PolicyLayout &L) {
    const int tid = threadIdx.x, O = L.O;
    for (int i = tid; i < HID * O; i += EV_THREADS) { const int j = i / O, o = i - j * O; smem[S.w0 + o * HID + ev_phys(j)] = __ldg(params + L.w0 + i); }
    for (int i = tid; i < HID; i += EV_THREADS) {
        smem[S.g0w + ev_phys(i)] = __ldg(params + L.g0w + i);
        smem[S.g0b + ev_phys(i)] = __ldg(params + L.g0b + i);
    }
    for (int h = 0; h < S.n_heads; ++h) {
        const HeadLayout &H = L.head[h];
        for (int i = tid; i < HID * HID; i += EV_THREADS) { const int j = i >> 6, k = i & 63; smem[S.w1[h] + ev_phys(k) * HID + ev_phys(j)] = __ldg(params + H.w1 + i); }
        for (int i = tid; i < HID; i += EV_THREADS) {
            smem[S.gw[h] + ev_phys(i)] = __ldg(params + H.gw + i);
            smem[S.gb[h] + ev_phys(i)] = __ldg(params + H.gb + i);
        }
        for (int i = tid; i < H.out * HID; i += EV_THREADS) { const int a = i >> 6, j = i & 63; smem[S.w2[h] + a * HID + ev_phys(j)] = __ldg(params + H.w2 + i); }
        for (int i = tid; i < H.out; i += EV_THREADS) smem[S.b2[h] + i] = __ldg(params + H.b2 + i);
    }
}

// z[0..7] += a * {wa, wb}: eight independent IEEE float32 FMAs, issued as four packed FFMA2 - the same results bit for bit (two
// independent roundings per instruction) in half the issue slots; the forward is issue-bound (2.6 warp instructions per cycle and
// SM, `not_selected` the top stall).  A/B on the B200 (PRL_EV_SCALAR_FMA = the scalar form): CartPole rollout 5.34 -> 5.13 ms,
// Pendulum rollout 62.6 -> 57.2 ms, all bit-identity tests unchanged.
__device__ __forceinline__ void ev_fma8(float (&z)[8], float a, const float4 &wa, const float4 &wb) {
#ifndef PRL_EV_SCALAR_FMA
    const float2 aa = make_float2(a, a);
    float2 t;
    t = __ffma2_rn(aa, make_float2(wa.x, wa.y), make_float2(z[0], z[1])); z[0] = t.x; z[1] = t.y;
    t = __ffma2_rn(aa, make_float2(wa.z, wa.w), make_float2(z[2], z[3])); z[2] = t.x; z[3] = t.y;
    t = __ffma2_rn(aa, make_float2(wb.x, wb.y), make_float2(z[4], z[5])); z[4] = t.x; z[5] = t.y;
    t = __ffma2_rn(aa, make_float2(wb.z, wb.w), make_float2(z[6], z[7])); z[6] = t.x; z[7] = t.y;
#else
    z[0] = fmaf(a, wa.x, z[0]); z[1] = fmaf(a, wa.y, z[1]); z[2] = fmaf(a, wa.z, z[2]); z[3] = fmaf(a, wa.w, z[3]);
    z[4] = fmaf(a, wb.x, z[4]); z[5] = fmaf(a, wb.y, z[5]); z[6] = fmaf(a, wb.z, z[6]); z[7] = fmaf(a, wb.w, z[7]);
#endif
}

// GroupNorm (this thread's 8 features = one group) + affine + SiLU on 8 rows; same operation order as gn_silu (mlp.cuh)
template <int RPT>
__device__ __forceinline__ void ev_gn_silu(float (&z)[RPT][8], const float *gw_p, const float *gb_p, int fg) {
    const float4 ga = *reinterpret_cast<const float4 *>(gw_p + 4 * fg), gb_ = *reinterpret_cast<const float4 *>(gw_p + 32 + 4 * fg);
    const float4 ba = *reinterpret_cast<const float4 *>(gb_p + 4 * fg), bb = *reinterpret_cast<const float4 *>(gb_p + 32 + 4 * fg);
    const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb_.x, gb_.y, gb_.z, gb_.w}, bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
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

// Forward of one 256-row tile.  In: sX = smem + S.x holds the tile's inputs [row][O] (written and synchronised by the
// caller).  Out: smem + S.out holds [row][S.so] head outputs (head h at columns S.col[h] ..), valid after the caller's
// next __syncthreads().  Contains one __syncthreads(); every thread of the CTA must call it.
// RPT = rows per thread: the tile is 32 RPT rows, thread (rg, fg) owns rows RPT rg .. RPT rg + RPT - 1.  8 is the throughput form
// (64 FMAs per 4 shared-memory loads); the fused rollout of a FEW envs (configs[0]: 32) runs RPT = 1, one row per 8 threads, so
// that all 8 warps share the 32 rows instead of one warp walking them alone - the step is latency-bound there.  A row's
// arithmetic (operation order, contraction) does not depend on RPT: the same bits.
template <int RPT = EV_RPT>
__device__ __forceinline__ void ev_forward_tile(float *smem, const EvSmem &S, const PolicyLayout &L) {
    const int tid = threadIdx.x, fg = tid & 7, rg = tid >> 3, O = L.O;
    float *sF = smem + S.f, *sO = smem + S.out;
    const float *sX = smem + S.x;
    float z[RPT][8];
    // ---- trunk: Linear(O, 64, no bias) -> GN -> SiLU -> sF
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) z[r][i] = 0.f;
    for (int o = 0; o < O; ++o) {
        const float4 wa = *reinterpret_cast<const float4 *>(smem + S.w0 + o * HID + 4 * fg);
        const float4 wb = *reinterpret_cast<const float4 *>(smem + S.w0 + o * HID + 32 + 4 * fg);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const float xv = sX[(RPT * rg + r) * O + o];
            ev_fma8(z[r], xv, wa, wb);
        }
    }
    ev_gn_silu<RPT>(z, smem + S.g0w, smem + S.g0b, fg);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        float *dst = sF + (RPT * rg + r) * HID + 4 * fg;
        *reinterpret_cast<float4 *>(dst) = make_float4(z[r][0], z[r][1], z[r][2], z[r][3]);
        *reinterpret_cast<float4 *>(dst + 32) = make_float4(z[r][4], z[r][5], z[r][6], z[r][7]);
    }
    __syncthreads();
    // ---- heads: Linear(64, 64, no bias) -> GN -> SiLU -> Linear(64, out) + bias
    for (int h = 0; h < S.n_heads; ++h) {
#pragma unroll
        for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) z[r][i] = 0.f;
        const float *Wp = smem + S.w1[h] + 4 * fg;
        const float *Fp = sF + (RPT * rg) * HID;
#pragma unroll 2
        for (int k4 = 0; k4 < HID / 4; ++k4) {
            float4 a[RPT];
#pragma unroll
            for (int r = 0; r < RPT; ++r) a[r] = *reinterpret_cast<const float4 *>(Fp + r * HID + 4 * k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 wa = *reinterpret_cast<const float4 *>(Wp + (4 * k4 + kk) * HID);
                const float4 wb = *reinterpret_cast<const float4 *>(Wp + (4 * k4 + kk) * HID + 32);
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
                    ev_fma8(z[r], av, wa, wb);
                }
            }
        }
        ev_gn_silu<RPT>(z, smem + S.gw[h], smem + S.gb[h], fg);
        const int outs = L.head[h].out;
        for (int a = 0; a < outs; ++a) {
            const float4 wa = *reinterpret_cast<const float4 *>(smem + S.w2[h] + a * HID + 4 * fg);
            const float4 wb = *reinterpret_cast<const float4 *>(smem + S.w2[h] + a * HID + 32 + 4 * fg);
            float mine = 0.f;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                float p = ((z[r][0] * wa.x + z[r][1] * wa.y) + (z[r][2] * wa.z + z[r][3] * wa.w)) +
                          ((z[r][4] * wb.x + z[r][5] * wb.y) + (z[r][6] * wb.z + z[r][7] * wb.w));
                p += __shfl_xor_sync(0xffffffffu, p, 1);
                p += __shfl_xor_sync(0xffffffffu, p, 2);
                p += __shfl_xor_sync(0xffffffffu, p, 4);
                mine = fg == r ? p : mine;
            }
            if (RPT == 8 || fg < RPT) sO[(RPT * rg + fg) * S.so + S.col[h] + a] = mine + smem[S.b2[h] + a];   // lane fg keeps row RPT rg + fg
        }
    }
}

// softmax -> Categorical(probs) (renormalised) -> inverse-CDF sample with one uniform, on A contiguous logits
// (overwritten with the probabilities); same arithmetic as sample_categorical (policy.cuh)
__device__ __forceinline__ int ev_sample_categorical(float *lg, int A, float u, float *probs_out) {
    float m = lg[0];
    for (int a = 1; a < A; ++a) m = fmaxf(m, lg[a]);
    float S = 0.f;
    for (int a = 0; a < A; ++a) { const float e = expf(lg[a] - m); lg[a] = e; S += e; }
    float P = 0.f;  // torch Categorical(probs) divides by probs.sum(-1) once more
    for (int a = 0; a < A; ++a) { const float p = lg[a] / S; lg[a] = p; P += p; }
    float cum = 0.f;
    int pick = A - 1;
    bool found = false;
    for (int a = 0; a < A; ++a) {
        const float p = lg[a] / P;
        if (probs_out) probs_out[a] = p;
        cum += p;
        if (!found && u < cum) { pick = a; found = true; }
    }
    return pick;
}
#endif

}  // namespace prl
