// PPO minibatch step on the 5th-generation tensor cores: fused forward + clipped-surrogate / SmoothL1 loss + backward
// for discrete policies (trunk + actor + critic), the throughput form of csrc/update_ppo.cu (same math, same flat
// gradient layout, same C-ABI contract).
//
// Reference: /root/reference/PPO/PPO.py:219-252 and PPO/ActorCritic.py:118-146.
//
// Work split.  One CTA = 128 threads = one 128-row tile at a time, persistent over tiles; thread r owns row r for all
// per-row math (GroupNorm, SiLU, softmax / loss, their backward), held in registers.  Everything that contracts over
// features or over rows runs as tcgen05.mma (kind::f16, bf16x3 split operands, fp32 accumulators in tensor memory):
//     forward   Z[r][(h,j)]  = sum_k F[r][k] W1_h[j][k]          M=128 N=128 K=64    (both heads in one GEMM)
//     dgrad     DF[r][k]    += sum_j DZ_h[r][j] W1_h[j][k]       M=128 N=64  K=64    (B = MN-major view of the same W1 bytes)
//     wgrad     DW_h[j][k]  += sum_r DZ_h[r][j] F[r][k]          M=128 N=64  K=128   (A, B = MN-major views of the same DZ / F bytes)
//     wgrad0    DW0[j][i]   += sum_r DZ0[r][j] X[r][i]           M=128 N=16  K=128
// The weight-gradient accumulators stay in tensor memory across ALL tiles of the CTA and are read out once.
// What remains on the CUDA cores are the row-wise nonlinearities and the narrow column sums (GroupNorm affine and
// output-layer gradients), done as warp butterfly reductions into per-warp register accumulators.
#include "policy.cuh"
#include "umma.cuh"

namespace prl {
using namespace umma;

constexpr int TC_THREADS = 128;
constexpr int PIECE = 8 * CHUNK;     // one bf16 piece of a [128][64] matrix: 8 chunks x 2048 B = 16 KB
constexpr int XPIECE = 2 * CHUNK;    // one bf16 piece of the [128][16] input matrix
constexpr int TC_MAX_O = 16, TC_MAX_A = 8;

// tensor-memory columns
constexpr uint32_t TM_Z = 0, TM_DF = 128, TM_DW = 192 /* + 64 h */, TM_DW0 = 320, TM_COLS = 512;

struct TcSmem {
    unsigned char *W, *F, *DZ, *X;   // bf16 piece buffers: W 3 x PIECE (n = h*64 + j), F 3 x PIECE, DZ 4 x PIECE (p0 p1 p2 zero), X 3 x XPIECE
    float *w0t, *g0w, *g0b;          // [O][64], [64], [64]
    float *gw[2], *gb[2], *w2[2], *b2[2];
    float *red;                      // [4 warps][NQ][64] final cross-warp combine
};

__host__ __device__ inline size_t tc_small_floats(const PolicyLayout &L) {
    size_t n = (size_t)L.O * HID + 2 * HID;
    for (int h = 0; h < 2; ++h) n += 2 * HID + (size_t)L.head[h].out * HID + round4(L.head[h].out);
    return n;
}
__host__ __device__ inline int tc_num_q(const PolicyLayout &L) { return 2 + 2 + L.head[0].out + 2 + L.head[1].out; }
__host__ __device__ inline size_t tc_smem_bytes(const PolicyLayout &L) {
    return 1024 + 3 * PIECE + 3 * PIECE + 4 * PIECE + 3 * XPIECE + tc_small_floats(L) * 4 + (size_t)4 * tc_num_q(L) * HID * 4 + 256;
}

// write one row's 64 fp32 values as three bf16 pieces into a [128][64] piece-buffer triple (row-per-thread layout)
__device__ __forceinline__ void store_row_pieces(unsigned char *base, int r, const float (&v)[HID]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t q0[4], q1[4], q2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_bf16x3(v[8 * c + 2 * i], v[8 * c + 2 * i + 1], q0[i], q1[i], q2[i]);
        unsigned char *p = base + c * CHUNK + r * 16;
        *reinterpret_cast<uint4 *>(p) = make_uint4(q0[0], q0[1], q0[2], q0[3]);
        *reinterpret_cast<uint4 *>(p + PIECE) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
        *reinterpret_cast<uint4 *>(p + 2 * PIECE) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
    }
}

// column sums over the warp's 32 rows: on return lane l holds the sums of features 2l and 2l+1 in v[0], v[1].
// 5 exchange steps, 62 shuffles; v is consumed.
__device__ __forceinline__ void warp_colsum64(float (&v)[HID], float &s0, float &s1) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int step = 0; step < 5; ++step) {
        const int off = 16 >> step, n = 32 >> step;   // partner distance, surviving length
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float keep = up ? v[i + n] : v[i];
            const float send = up ? v[i] : v[i + n];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    s0 = v[0];
    s1 = v[1];
}

// GroupNorm statistics of a row in registers: z -> zhat in place, rstd per group
__device__ __forceinline__ void tc_gn_normalize(float (&z)[HID], float (&rstd)[GROUPS]) {
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
// dy -> dz in place: d = dy * gamma, dz = rstd * (d - mean(d) - zhat * mean(d * zhat)) per group
__device__ __forceinline__ void tc_gn_backward(float (&d)[HID], const float (&zhat)[HID], const float (&rstd)[GROUPS], const float *gamma) {
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) {
            const int j = g * GSIZE + i;
            d[j] *= gamma[j];
            m1 += d[j];
            m2 = fmaf(d[j], zhat[j], m2);
        }
        m1 *= (1.0f / GSIZE); m2 *= (1.0f / GSIZE);
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) {
            const int j = g * GSIZE + i;
            d[j] = rstd[g] * (d[j] - m1 - zhat[j] * m2);
        }
    }
}

__device__ __forceinline__ void tc_sync_for_mma() {
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
}

// ---- MMA issue (one thread) ---------------------------------------------------------------------------------------
struct TcAddr {
    uint32_t W, F, DZ, X, tmem;
};
// six-term product of two bf16x3 operands: (piece of A, piece of B)
__device__ __constant__ const int TERM_A[6] = {0, 0, 1, 1, 0, 2};
__device__ __constant__ const int TERM_B[6] = {0, 1, 0, 1, 2, 0};

__device__ __forceinline__ void issue_forward(const TcAddr &a) {
    constexpr uint32_t id = idesc_bf16(128, 128, 0, 0);
#pragma unroll 1
    for (int t = 0; t < 6; ++t)
#pragma unroll
        for (int s = 0; s < 4; ++s)
            mma_bf16_ss(a.tmem + TM_Z, smem_desc(a.F + TERM_A[t] * PIECE + s * 2 * CHUNK, CHUNK, 128),
                        smem_desc(a.W + TERM_B[t] * PIECE + s * 2 * CHUNK, CHUNK, 128), id, (t | s) != 0);
}
__device__ __forceinline__ void issue_head_backward(const TcAddr &a, int h, bool first_tile) {
    constexpr uint32_t id_d = idesc_bf16(128, 64, 0, 1), id_w = idesc_bf16(128, 64, 1, 1);
    // dgrad: DF (+)= DZ_h . W1_h   (B: MN-major view, MN = k groups CHUNK apart, K = j groups 128 B apart, head h at + h*64*16)
#pragma unroll 1
    for (int t = 0; t < 6; ++t)
#pragma unroll
        for (int s = 0; s < 4; ++s)
            mma_bf16_ss(a.tmem + TM_DF, smem_desc(a.DZ + TERM_A[t] * PIECE + s * 2 * CHUNK, CHUNK, 128),
                        smem_desc(a.W + TERM_B[t] * PIECE + h * 64 * 16 + s * 256, 128, CHUNK), id_d, (h | t | s) != 0);
    // wgrad: DW_h[(piece window, j)][k] += sum_r DZ[r][.] F[r][k]; windows [p0|p1] x f0, f1, f2 and [p2|0] x f0
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
        const uint32_t win = (t == 3) ? 2 * PIECE : 0, fp = (t == 3) ? 0 : t * PIECE;
#pragma unroll
        for (int s = 0; s < 8; ++s)
            mma_bf16_ss(a.tmem + TM_DW + 64 * h, smem_desc(a.DZ + win + s * 256, 128, CHUNK), smem_desc(a.F + fp + s * 256, 128, CHUNK), id_w,
                        !(first_tile && t == 0 && s == 0));
    }
}
__device__ __forceinline__ void issue_trunk_wgrad(const TcAddr &a, bool first_tile) {
    constexpr uint32_t id = idesc_bf16(128, 16, 1, 1);
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
        const uint32_t win = (t == 3) ? 2 * PIECE : 0, xp = (t == 3) ? 0 : t * XPIECE;
#pragma unroll
        for (int s = 0; s < 8; ++s)
            mma_bf16_ss(a.tmem + TM_DW0, smem_desc(a.DZ + win + s * 256, 128, CHUNK), smem_desc(a.X + xp + s * 256, 128, CHUNK), id,
                        !(first_tile && t == 0 && s == 0));
    }
}

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[HID]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float t[16];
        tmem_ld16(taddr + 16 * c, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[16 * c + i] = t[i];
    }
    tmem_ld_wait();
}

// ===================================================================================================== the kernel
// NA = compile-time bound of the output widths (action_dim rounded up to 2, 4 or 8): the per-output loops unroll over it
template <int NA>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_ppo_grad_tc(const float *__restrict__ params, PolicyLayout L, const float *__restrict__ states, const float *__restrict__ actions,
              const float *__restrict__ old_logp, const float *__restrict__ adv, const float *__restrict__ returns, int64_t b,
              float clip, float inv_count, float *__restrict__ partials, double *__restrict__ loss_partials, int *__restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bars[4];          // forward done, actor backward done, critic backward done, trunk wgrad done
    __shared__ uint32_t tmem_slot;
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int O = L.O, A = L.A;

    // ---- carve shared memory
    TcSmem S;
    {
        unsigned char *p = smem_raw;
        S.W = p; p += 3 * PIECE;
        S.F = p; p += 3 * PIECE;
        S.DZ = p; p += 4 * PIECE;
        S.X = p; p += 3 * XPIECE;
        float *f = reinterpret_cast<float *>(p);
        S.w0t = f; f += O * HID;
        S.g0w = f; f += HID;
        S.g0b = f; f += HID;
        for (int h = 0; h < 2; ++h) {
            S.gw[h] = f; f += HID;
            S.gb[h] = f; f += HID;
            S.w2[h] = f; f += L.head[h].out * HID;
            S.b2[h] = f; f += round4(L.head[h].out);
        }
        S.red = f;
    }
    // ---- stage parameters: fp32 small ones, W1 of both heads as bf16x3 in the operand layout (row n = h*64 + j)
    stage_transposed(S.w0t, params + L.w0, HID, O);
    stage_copy(S.g0w, params + L.g0w, HID);
    stage_copy(S.g0b, params + L.g0b, HID);
    for (int h = 0; h < 2; ++h) {
        stage_copy(S.gw[h], params + L.head[h].gw, HID);
        stage_copy(S.gb[h], params + L.head[h].gb, HID);
        stage_copy(S.w2[h], params + L.head[h].w2, L.head[h].out * HID);
        stage_copy(S.b2[h], params + L.head[h].b2, L.head[h].out);
    }
    {
        const float *wrow = params + L.head[tid >> 6].w1 + (tid & 63) * HID;   // thread n stages row n of the stacked [128][64] W1
        float v[HID];
#pragma unroll
        for (int k = 0; k < HID; ++k) v[k] = __ldg(wrow + k);
        store_row_pieces(S.W, tid, v);
        // zero slot behind the three DZ pieces, and the X pieces (columns >= O stay zero for the whole kernel)
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4 *>(S.DZ + 3 * PIECE + c * CHUNK + tid * 16) = make_uint4(0, 0, 0, 0);
        for (int c = 0; c < 6; ++c) *reinterpret_cast<uint4 *>(S.X + c * CHUNK + tid * 16) = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, TM_COLS);
    tc_sync_for_mma();
    fence_after_sync();
    TcAddr addr{smem_u32(S.W), smem_u32(S.F), smem_u32(S.DZ), smem_u32(S.X), tmem_slot};
    const uint32_t lane_base = addr.tmem + ((uint32_t)(warp * 32) << 16);
    bool mma_ok = true;

    // per-warp column-sum accumulators, features 2*lane and 2*lane+1: trunk (dgamma, dbeta), head h (dgamma, dbeta, dW2[a])
    float q_g0[2] = {0.f, 0.f}, q_b0[2] = {0.f, 0.f};
    float q_g[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, q_b[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    float q_w2[2][NA][2];
    float q_b2[2][NA];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int a = 0; a < NA; ++a) { q_w2[h][a][0] = q_w2[h][a][1] = 0.f; q_b2[h][a] = 0.f; }
    double l_pol = 0.0, l_val = 0.0, l_ent = 0.0;

    const int64_t ntiles = (b + TC_THREADS - 1) / TC_THREADS;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t parity = it & 1;
        const int64_t row = tile * TC_THREADS + tid;
        const bool live = row < b;
        const float *xrow = states + (live ? row : 0) * O;   // dead rows read row 0 and are masked below
        const float xmask = live ? 1.f : 0.f;

        // ================= trunk forward (CUDA cores): z0 = W0 x, GroupNorm, SiLU -> F pieces, X pieces
        {
            float z[HID], rstd[GROUPS];
#pragma unroll
            for (int j = 0; j < HID; ++j) z[j] = 0.f;
            for (int i = 0; i < O; ++i) axpy64(xmask * __ldg(xrow + i), S.w0t + i * HID, z);
            tc_gn_normalize(z, rstd);
#pragma unroll
            for (int j = 0; j < HID; ++j) z[j] = silu(fmaf(z[j], S.g0w[j], S.g0b[j]));
            // the previous tile's trunk-wgrad MMAs read X and DZ; its head MMAs (already waited for) read F
            if (it > 0) mma_ok &= mbar_wait(&bars[3], parity ^ 1);
            store_row_pieces(S.F, tid, z);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t q0[4], q1[4], q2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = 8 * c + 2 * i;
                    const float xa = e < O ? xmask * __ldg(xrow + e) : 0.f, xb = e + 1 < O ? xmask * __ldg(xrow + e + 1) : 0.f;
                    split_bf16x3(xa, xb, q0[i], q1[i], q2[i]);
                }
                unsigned char *p = S.X + c * CHUNK + tid * 16;
                *reinterpret_cast<uint4 *>(p) = make_uint4(q0[0], q0[1], q0[2], q0[3]);
                *reinterpret_cast<uint4 *>(p + XPIECE) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
                *reinterpret_cast<uint4 *>(p + 2 * XPIECE) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
            }
        }
        tc_sync_for_mma();
        if (tid == 0) {
            fence_after_sync();
            issue_forward(addr);
            mma_commit(&bars[0]);
        }
        const float adv_i = live ? adv[row] : 0.f, old_i = live ? old_logp[row] : 0.f;
        const float ret_i = live ? returns[row] : 0.f;
        const int act = live ? (int)actions[row] : 0;
        mma_ok &= mbar_wait(&bars[0], parity);
        fence_after_sync();

        // ================= heads: forward epilogue, loss, backward epilogue -> DZ pieces, tensor-core dgrad + wgrad
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int nout = L.head[h].out;
            const float *gw = S.gw[h], *gb = S.gb[h], *w2 = S.w2[h];
            float zhat[HID], rstd[GROUPS], sg[HID];
            tmem_ld64(lane_base + TM_Z + 64 * h, zhat);
            tc_gn_normalize(zhat, rstd);
            float out[NA];
#pragma unroll
            for (int a = 0; a < NA; ++a) out[a] = (a < nout) ? S.b2[h][a] : 0.f;
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                const float y = fmaf(zhat[j], gw[j], gb[j]);
                const float s = 1.0f / (1.0f + expf(-y));
                sg[j] = s;
                const float hj = y * s;
#pragma unroll
                for (int a = 0; a < NA; ++a)
                    if (a < nout) out[a] = fmaf(hj, w2[a * HID + j], out[a]);
            }
            // ---- loss and output gradients
            float dout[NA];
#pragma unroll
            for (int a = 0; a < NA; ++a) dout[a] = 0.f;
            if (h == 0) {
                if (live) {
                    float m = out[0];
#pragma unroll
                    for (int a = 1; a < NA; ++a)
                        if (a < A) m = fmaxf(m, out[a]);
                    float p[NA], Ssum = 0.f, Psum = 0.f;
#pragma unroll
                    for (int a = 0; a < NA; ++a) { p[a] = (a < A) ? expf(out[a] - m) : 0.f; Ssum += p[a]; }
#pragma unroll
                    for (int a = 0; a < NA; ++a) { p[a] = p[a] / Ssum; Psum += p[a]; }
                    float pa = 0.f, ent = 0.f;
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
                        if (a < A) {
                            p[a] = p[a] / Psum;
                            const float l = logf(fminf(fmaxf(p[a], F32_EPS), 1.0f - F32_EPS));
                            ent -= l * p[a];
                            if (a == act) pa = p[a];
                        }
                    }
                    const float logp = logf(fminf(fmaxf(pa, F32_EPS), 1.0f - F32_EPS));
                    const float dl = logp - old_i;
                    const float r = expf(fminf(fmaxf(dl, -20.f), 20.f));
                    const float s1 = r * adv_i;
                    const float s2 = fminf(fmaxf(r, 1.0f - clip), 1.0f + clip) * adv_i;
                    const float g1 = s1 < s2 ? 1.f : (s1 > s2 ? 0.f : 0.5f);   // torch.min splits ties evenly
                    const float in_clip = (r >= 1.0f - clip && r <= 1.0f + clip) ? 1.f : 0.f;
                    const float in20 = (dl >= -20.f && dl <= 20.f) ? 1.f : 0.f;
                    float dlogp = -inv_count * adv_i * (g1 + (1.f - g1) * in_clip) * r * in20;
                    if (!(pa >= F32_EPS && pa <= 1.0f - F32_EPS)) dlogp = 0.f;   // clamp in probs_to_logits blocks the gradient
#pragma unroll
                    for (int a = 0; a < NA; ++a)
                        if (a < A) dout[a] = dlogp * ((a == act ? 1.f : 0.f) - p[a]);
                    l_pol += -fminf(s1, s2);
                    l_ent += ent;
                }
            } else if (live) {
                const float dv = out[0] - ret_i, ad = fabsf(dv);
                l_val += ad < 1.f ? 0.5f * dv * dv : ad - 0.5f;
                dout[0] = 0.5f * inv_count * (ad < 1.f ? dv : (dv > 0.f ? 1.f : -1.f));
            }
            // ---- output-layer gradients: dW2[a][j] = sum_r dout[a] h_j, db2[a] = sum_r dout[a]
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                if (a < nout) {
                    float t[HID];
#pragma unroll
                    for (int j = 0; j < HID; ++j) t[j] = dout[a] * (fmaf(zhat[j], gw[j], gb[j]) * sg[j]);
                    float s0, s1;
                    warp_colsum64(t, s0, s1);
                    q_w2[h][a][0] += s0; q_w2[h][a][1] += s1;
                    q_b2[h][a] += warp_sum(dout[a]);
                }
            }
            // ---- dy (in place of sg), GroupNorm-affine gradients
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                float dh = 0.f;
#pragma unroll
                for (int a = 0; a < NA; ++a)
                    if (a < nout) dh = fmaf(dout[a], w2[a * HID + j], dh);
                const float y = fmaf(zhat[j], gw[j], gb[j]);
                sg[j] = dh * sg[j] * fmaf(y, 1.0f - sg[j], 1.0f);
            }
            {
                float t[HID], s0, s1;
#pragma unroll
                for (int j = 0; j < HID; ++j) t[j] = sg[j] * zhat[j];
                warp_colsum64(t, s0, s1);
                q_g[h][0] += s0; q_g[h][1] += s1;
#pragma unroll
                for (int j = 0; j < HID; ++j) t[j] = sg[j];
                warp_colsum64(t, s0, s1);
                q_b[h][0] += s0; q_b[h][1] += s1;
            }
            tc_gn_backward(sg, zhat, rstd, gw);   // sg now holds dz
            // the actor's MMAs read DZ: they must have completed before the critic overwrites it
            if (h == 1) mma_ok &= mbar_wait(&bars[1], parity);
            store_row_pieces(S.DZ, tid, sg);
            tc_sync_for_mma();
            if (tid == 0) {
                fence_after_sync();
                issue_head_backward(addr, h, it == 0);
                mma_commit(&bars[1 + h]);
            }
        }

        // ================= trunk backward: DF -> dy0 -> GroupNorm backward -> DZ pieces, tensor-core wgrad against X
        {
            float zhat[HID], rstd[GROUPS], df[HID];
#pragma unroll
            for (int j = 0; j < HID; ++j) zhat[j] = 0.f;
            for (int i = 0; i < O; ++i) axpy64(xmask * __ldg(xrow + i), S.w0t + i * HID, zhat);
            tc_gn_normalize(zhat, rstd);
            mma_ok &= mbar_wait(&bars[2], parity);    // critic dgrad complete -> DF final; DZ free again
            fence_after_sync();
            tmem_ld64(lane_base + TM_DF, df);
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                const float y = fmaf(zhat[j], S.g0w[j], S.g0b[j]);
                const float s = 1.0f / (1.0f + expf(-y));
                df[j] = df[j] * s * fmaf(y, 1.0f - s, 1.0f);
            }
            {
                float t[HID], s0, s1;
#pragma unroll
                for (int j = 0; j < HID; ++j) t[j] = df[j] * zhat[j];
                warp_colsum64(t, s0, s1);
                q_g0[0] += s0; q_g0[1] += s1;
#pragma unroll
                for (int j = 0; j < HID; ++j) t[j] = df[j];
                warp_colsum64(t, s0, s1);
                q_b0[0] += s0; q_b0[1] += s1;
            }
            tc_gn_backward(df, zhat, rstd, S.g0w);
            store_row_pieces(S.DZ, tid, df);
            tc_sync_for_mma();
            if (tid == 0) {
                fence_after_sync();
                issue_trunk_wgrad(addr, it == 0);
                mma_commit(&bars[3]);
            }
        }
    }

    // ================= read the accumulators out: tensor memory -> this block's partial-gradient row
    const int P = L.total;
    float *part = partials + (size_t)blockIdx.x * P;
    if (it > 0) mma_ok &= mbar_wait(&bars[3], (it - 1) & 1);
    fence_after_sync();
    float *scratch = reinterpret_cast<float *>(S.F);   // 64 x 64 fp32 = 16 KB, free now
    for (int h = 0; h < 2; ++h) {
        float v[HID];
        tmem_ld64(lane_base + TM_DW + 64 * h, v);     // lanes 0..63: first piece of the window, lanes 64..127: second
        __syncthreads();
        if (tid >= 64) {
#pragma unroll
            for (int k = 0; k < HID; ++k) scratch[(tid - 64) * (HID + 1) + k] = v[k];
        }
        __syncthreads();
        if (tid < 64) {
            float *dst = part + L.head[h].w1 + tid * HID;
#pragma unroll
            for (int k = 0; k < HID; ++k) dst[k] = (it > 0) ? v[k] + scratch[tid * (HID + 1) + k] : 0.f;
        }
    }
    {
        float v[16];
        tmem_ld16(lane_base + TM_DW0, v);
        tmem_ld_wait();
        __syncthreads();
        if (tid >= 64) {
#pragma unroll
            for (int i = 0; i < 16; ++i) scratch[(tid - 64) * 17 + i] = v[i];
        }
        __syncthreads();
        if (tid < 64) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < O) part[L.w0 + tid * O + i] = (it > 0) ? v[i] + scratch[tid * 17 + i] : 0.f;
        }
    }
    // column-sum accumulators: combine the four warps through shared memory, features 2*lane, 2*lane+1
    {
        __syncthreads();
        const int NQ = tc_num_q(L);
        float *r4 = S.red + (size_t)warp * NQ * HID;
        int q = 0;
        auto put = [&](const float (&s)[2]) { r4[q * HID + 2 * lane] = s[0]; r4[q * HID + 2 * lane + 1] = s[1]; ++q; };
        put(q_g0); put(q_b0);
        for (int h = 0; h < 2; ++h) {
            put(q_g[h]); put(q_b[h]);
#pragma unroll
            for (int a = 0; a < NA; ++a)
                if (a < L.head[h].out) put(q_w2[h][a]);
        }
        __syncthreads();
        // destination offsets of the NQ vectors in the flat gradient
        for (int idx = tid; idx < NQ * HID; idx += TC_THREADS) {
            const int qq = idx / HID, j = idx - qq * HID;
            const float s = (S.red[(0 * NQ + qq) * HID + j] + S.red[(1 * NQ + qq) * HID + j]) + (S.red[(2 * NQ + qq) * HID + j] + S.red[(3 * NQ + qq) * HID + j]);
            int off;
            if (qq == 0) off = L.g0w;
            else if (qq == 1) off = L.g0b;
            else {
                int k = qq - 2, h = 0;
                if (k >= 2 + L.head[0].out) { k -= 2 + L.head[0].out; h = 1; }
                off = (k == 0) ? L.head[h].gw : (k == 1) ? L.head[h].gb : L.head[h].w2 + (k - 2) * HID;
            }
            part[off + j] = s;
        }
        // db2: every lane of a warp holds the warp total; combine warps
        __shared__ float b2s[4][2][NA];
        if (lane == 0)
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int a = 0; a < NA; ++a) b2s[warp][h][a] = q_b2[h][a];
        __syncthreads();
        if (tid < 2 * NA) {
            const int h = tid / NA, a = tid % NA;
            if (a < L.head[h].out) part[L.head[h].b2 + a] = (b2s[0][h][a] + b2s[1][h][a]) + (b2s[2][h][a] + b2s[3][h][a]);
        }
    }
    const double bp = block_sum<double>(l_pol, red);
    const double bv = block_sum<double>(l_val, red);
    const double be = block_sum<double>(l_ent, red);
    if (tid == 0) {
        loss_partials[blockIdx.x * 4 + 0] = bp;
        loss_partials[blockIdx.x * 4 + 1] = bv;
        loss_partials[blockIdx.x * 4 + 2] = be;
        loss_partials[blockIdx.x * 4 + 3] = 0.0;
        if (!mma_ok) atomicExch(status, 1);
    }
    if (!mma_ok && lane == 0 && tid != 0) atomicExch(status, 1);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(addr.tmem, TM_COLS);
}

// grad[i] = sum over blocks of partials[b][i], fixed order (bit-reproducible); loss_out += block loss partials
__global__ void k_reduce_partials_tc(const float *__restrict__ partials, int nblocks, int P, float *__restrict__ grad,
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

static int tc_grid(int64_t b) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ntiles = (b + TC_THREADS - 1) / TC_THREADS;
    return (int)(ntiles < sms ? (ntiles > 0 ? ntiles : 1) : sms);
}

}  // namespace prl

using namespace prl;

extern "C" {

int prl_ppo_grad_tc_supported(int is_continuous, int obs_dim, int action_dim) {
    return !is_continuous && obs_dim >= 1 && obs_dim <= TC_MAX_O && action_dim >= 1 && action_dim <= TC_MAX_A;
}

size_t prl_update_tc_ws_floats(int is_continuous, int obs_dim, int action_dim, int64_t batch) {
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    const int grid = tc_grid(batch);
    return (size_t)grid * L.total + (size_t)grid * 8 + 16;
}

int prl_ppo_grad_tc(const float *params, int is_continuous, int obs_dim, int action_dim, const float *states, const float *actions,
                    const float *old_logp, const float *adv, const float *returns, int64_t b, float policy_clip, float inv_count,
                    float *grad, double *loss_out, float *ws, size_t ws_floats, void *stream) {
    PRL_REQUIRE(params && states && actions && old_logp && adv && returns && grad && ws && b > 0, "prl_ppo_grad_tc: bad arguments");
    PRL_REQUIRE(prl_ppo_grad_tc_supported(is_continuous, obs_dim, action_dim),
                "prl_ppo_grad_tc: only discrete policies with observ_dim <= %d and action_dim <= %d (got continuous=%d O=%d A=%d)", TC_MAX_O,
                TC_MAX_A, is_continuous, obs_dim, action_dim);
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    const int grid = tc_grid(b);
    PRL_REQUIRE(ws_floats >= (size_t)grid * L.total + (size_t)grid * 8 + 16, "prl_ppo_grad_tc: workspace too small");
    const size_t smem = tc_smem_bytes(L);
    PRL_REQUIRE(smem <= 227 * 1024, "prl_ppo_grad_tc: needs %zu B shared memory (> 227 KB)", smem);
    cudaStream_t st = (cudaStream_t)stream;
    float *partials = ws;
    double *loss_partials = reinterpret_cast<double *>(ws + (((size_t)grid * L.total + 1) & ~(size_t)1));
    int *status = reinterpret_cast<int *>(loss_partials + (size_t)grid * 4);
    PRL_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
    auto launch = [&](auto kernel) -> int {
        PRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<grid, TC_THREADS, smem, st>>>(params, L, states, actions, old_logp, adv, returns, b, policy_clip, inv_count, partials,
                                             loss_partials, status);
        return PRL_OK;
    };
    const int rc = action_dim <= 2 ? launch(k_ppo_grad_tc<2>) : action_dim <= 4 ? launch(k_ppo_grad_tc<4>) : launch(k_ppo_grad_tc<8>);
    if (rc != PRL_OK) return rc;
    k_reduce_partials_tc<<<cdiv(L.total, 256), 256, 0, st>>>(partials, grid, L.total, grad, loss_partials, loss_out, (double)b);
    return check_launch("k_ppo_grad_tc");
}

/* 0 = every tensor-core phase completed; 1 = an mbarrier wait timed out (results invalid).  Host-synchronising. */
int prl_ppo_grad_tc_status(const float *ws, int is_continuous, int obs_dim, int action_dim, int64_t batch, int *status_host, void *stream) {
    PRL_REQUIRE(ws && status_host, "prl_ppo_grad_tc_status: bad arguments");
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    const int grid = tc_grid(batch);
    const double *loss_partials = reinterpret_cast<const double *>(ws + (((size_t)grid * L.total + 1) & ~(size_t)1));
    PRL_CUDA(cudaMemcpyAsync(status_host, loss_partials + (size_t)grid * 4, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PRL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PRL_OK;
}

}  // extern "C"
