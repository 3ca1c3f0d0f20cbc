// PPO minibatch step on the 5th-generation tensor cores: fused forward + clipped-surrogate / SmoothL1 loss + backward
// for discrete policies (trunk + actor + critic), the throughput form of csrc/update_ppo.cu (same math, same flat
// gradient layout, same C-ABI contract).
//
// Reference: /root/reference/PPO/PPO.py:219-252 and PPO/ActorCritic.py:118-146.
//
// Work split.  One CTA = 16 compute warps + 1 MMA-issue warp = one 128-row tile at a time, persistent over tiles.  Row r
// of the tile is owned by FOUR compute threads (warps w, w+4, w+8, w+12 share a 32-row quarter; thread q of a row holds
// features 16q..16q+15 = two GroupNorm groups), so every per-row array is 16 wide and lives in registers.
// Everything that contracts over features or over rows runs as tcgen05.mma (kind::f16 on bf16x3 split operands, fp32
// accumulators in tensor memory), issued by one elected lane of warp 16, which does nothing else: the compute warps hand
// it staged operands through named barriers (arrive, no wait) and only ever wait for MMA completion (mbarriers):
//     forward   Z[r][(h,j)]  = sum_k F[r][k] W1_h[j][k]          M=128 N=128 K=64    (both heads in one GEMM)
//     dgrad     DF[r][k]    += sum_j DZ_h[r][j] W1_h[j][k]       M=128 N=64  K=64    (B = MN-major view of the same W1 bytes)
//     wgrad     DW_h[j][k]  += sum_r DZ_h[r][j] F[r][k]          M=128 N=64  K=128   (A, B = MN-major views of the same DZ / F bytes)
//     wgrad0    DW0[j][i]   += sum_r DZ0[r][j] X[r][i]           M=128 N=16  K=128
// The weight-gradient accumulators stay in tensor memory across ALL tiles of the CTA and are read out once.
// What remains on the CUDA cores are the row-wise nonlinearities (GroupNorm, SiLU, softmax / loss and their backward)
// and the narrow column sums (GroupNorm affine and output-layer gradients), done as warp butterfly reductions into
// per-warp register accumulators that are combined once at the end.
#include <stdlib.h>
#include <string.h>

#include "policy.cuh"
#include "umma.cuh"

namespace prl {
using namespace umma;

// 16 compute warps + 1 MMA-issue warp; rows per tile; features per thread
constexpr int TC_COMPUTE = 512, TC_THREADS = TC_COMPUTE + 32, TC_ROWS = 128, TC_W = 16;
constexpr int PIECE = 8 * CHUNK;     // one bf16 piece of a [128][64] matrix: 8 chunks x 2048 B = 16 KB
constexpr int XPIECE = 2 * CHUNK;    // one bf16 piece of the [128][16] input matrix
constexpr int TC_MAX_O = 16, TC_MAX_A = 8;

// tensor-memory columns
constexpr uint32_t TM_Z = 0, TM_DF = 128, TM_DW = 192 /* + 64 h */, TM_DW0 = 320, TM_COLS = 512;

__host__ __device__ inline int tc_small_floats(const PolicyLayout &L) {
    int n = L.O * HID + 2 * HID;
    for (int h = 0; h < 2; ++h) n += 2 * HID + L.head[h].out * HID + round4(L.head[h].out);
    return n;
}
__host__ __device__ inline int tc_num_q(const PolicyLayout &L) { return 2 + 2 + L.head[0].out + 2 + L.head[1].out; }
__host__ __device__ inline size_t tc_smem_bytes(const PolicyLayout &L, int NA) {
    return 1024 + 3 * PIECE + 3 * PIECE + 4 * PIECE + 3 * XPIECE + (size_t)tc_small_floats(L) * 4 + (size_t)4 * TC_ROWS * NA * 4 +
           (size_t)4 * tc_num_q(L) * HID * 4 + 256;
}

// ---- small math ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_sigmoid(float y) { return __fdividef(1.0f, 1.0f + __expf(-y)); }

// write 16 fp32 values (features 16q..16q+15 of row r) as three bf16 pieces: chunks 2q, 2q+1 of a [128][64] piece triple
__device__ __forceinline__ void store_pieces16(unsigned char *base, int r, int q, const float (&v)[TC_W]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t q0[4], q1[4], q2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_bf16x3(v[8 * c + 2 * i], v[8 * c + 2 * i + 1], q0[i], q1[i], q2[i]);
        unsigned char *p = base + (2 * q + c) * CHUNK + r * 16;
        *reinterpret_cast<uint4 *>(p) = make_uint4(q0[0], q0[1], q0[2], q0[3]);
        *reinterpret_cast<uint4 *>(p + PIECE) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
        *reinterpret_cast<uint4 *>(p + 2 * PIECE) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
    }
}

// column sum of 8 per-row values over the warp's 32 rows: on return every lane holds the total of feature
// f(lane) = 4*bit4 + 2*bit3 + bit2 of its lane index (4 lanes hold each feature).  9 shuffles.
__device__ __forceinline__ float colsum8(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7) {
    const int lane = threadIdx.x & 31;
    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
    const float a0 = (u16 ? v4 : v0) + __shfl_xor_sync(0xffffffffu, u16 ? v0 : v4, 16);
    const float a1 = (u16 ? v5 : v1) + __shfl_xor_sync(0xffffffffu, u16 ? v1 : v5, 16);
    const float a2 = (u16 ? v6 : v2) + __shfl_xor_sync(0xffffffffu, u16 ? v2 : v6, 16);
    const float a3 = (u16 ? v7 : v3) + __shfl_xor_sync(0xffffffffu, u16 ? v3 : v7, 16);
    const float b0 = (u8 ? a2 : a0) + __shfl_xor_sync(0xffffffffu, u8 ? a0 : a2, 8);
    const float b1 = (u8 ? a3 : a1) + __shfl_xor_sync(0xffffffffu, u8 ? a1 : a3, 8);
    float c0 = (u4 ? b1 : b0) + __shfl_xor_sync(0xffffffffu, u4 ? b0 : b1, 4);
    c0 += __shfl_xor_sync(0xffffffffu, c0, 2);
    c0 += __shfl_xor_sync(0xffffffffu, c0, 1);
    return c0;
}
// column sums of the thread's two groups, accumulated into the warp's 16-float slot of the shared accumulator array
// (slot[8 g + f(lane)] += total; one of the four lanes holding a feature does the update)
__device__ __forceinline__ void colsum16(const float (&v)[TC_W], float *slot) {
    const float c0 = colsum8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    const float c1 = colsum8(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
    const int lane = threadIdx.x & 31;
    if ((lane & 3) == 0) {
        const int f = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        slot[f] += c0;
        slot[8 + f] += c1;
    }
}

// GroupNorm on the thread's two groups: z -> zhat in place, rstd per group
__device__ __forceinline__ void gn_normalize16(float (&z)[TC_W], float (&rstd)[2]) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) m += z[g * GSIZE + i];
        m *= (1.0f / GSIZE);
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) { const float d = z[g * GSIZE + i] - m; v = fmaf(d, d, v); }
        const float r = rsqrtf(v * (1.0f / GSIZE) + GN_EPS);
        rstd[g] = r;
#pragma unroll
        for (int i = 0; i < GSIZE; ++i) z[g * GSIZE + i] = (z[g * GSIZE + i] - m) * r;
    }
}
// dy -> dz in place: d = dy * gamma, dz = rstd * (d - mean(d) - zhat * mean(d * zhat)) per group
__device__ __forceinline__ void gn_backward16(float (&d)[TC_W], const float (&zhat)[TC_W], const float (&rstd)[2], const float (&gamma)[TC_W]) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
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
__device__ __forceinline__ void load16(const float *src, float (&v)[TC_W]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = reinterpret_cast<const float4 *>(src)[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}

// named barrier 1 = the 512 compute threads among themselves (the MMA warp never joins it)
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// compute side of an "operands staged" hand-off to the MMA warp: make the shared-memory stores visible to the tensor core
// (async proxy), then arrive on the hand-off mbarrier (count 512) without waiting
__device__ __forceinline__ void staged(uint64_t *bar) {
    fence_async_smem();
    fence_before_sync();
    mbar_arrive(bar);
}

// ---- MMA issue: all lanes of the MMA warp call these (warp-uniform), one elected lane issues --------------------------------
struct TcDesc {          // base descriptors, K-major and MN-major views, built once per kernel
    uint64_t F_k, W_k, DZ_k, W_mn, DZ_mn, F_mn, X_mn;
    uint32_t tmem;
};
__device__ __forceinline__ uint64_t dadd(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// forward of head h: Z_h = F . W1_h^T  (W1 rows 64 h .. 64 h + 63 of the stacked [128][64] weight tile, N = 64), committed per
// head so that the actor's epilogue starts while the critic's MMAs still run
__device__ __forceinline__ void issue_forward(const TcDesc &D, int h) {
    constexpr uint32_t id = idesc_bf16(128, 64, 0, 0);
    constexpr int TA[6] = {0, 0, 1, 1, 0, 2}, TB[6] = {0, 1, 0, 1, 2, 0};
    const uint64_t wh = dadd(D.W_k, h * 64 * 16);
#pragma unroll
    for (int t = 0; t < 6; ++t)
#pragma unroll
        for (int s = 0; s < 4; ++s)
            mma_bf16_ss(D.tmem + TM_Z + 64 * h, dadd(D.F_k, TA[t] * PIECE + s * 2 * CHUNK), dadd(wh, TB[t] * PIECE + s * 2 * CHUNK), id, (t | s) != 0);
}
__device__ __forceinline__ void issue_head_dgrad(const TcDesc &D, int h) {
    constexpr uint32_t id_d = idesc_bf16(128, 64, 0, 1);
    constexpr int TA[6] = {0, 0, 1, 1, 0, 2}, TB[6] = {0, 1, 0, 1, 2, 0};
    // dgrad: DF (+)= DZ_h . W1_h   (B: MN-major view of W1, head h starts 64 rows = 1024 B in; 16 j's per step = 256 B)
    const uint64_t wh = dadd(D.W_mn, h * 64 * 16);
#pragma unroll
    for (int t = 0; t < 6; ++t)
#pragma unroll
        for (int s = 0; s < 4; ++s)
            mma_bf16_ss(D.tmem + TM_DF, dadd(D.DZ_k, TA[t] * PIECE + s * 2 * CHUNK), dadd(wh, TB[t] * PIECE + s * 256), id_d, (h | t | s) != 0);
}
__device__ __forceinline__ void issue_head_wgrad(const TcDesc &D, int h, bool first_tile) {
    constexpr uint32_t id_w = idesc_bf16(128, 64, 1, 1);
    // wgrad: DW_h[(piece window, j)][k] += sum_r DZ[r][.] F[r][k]; windows [p0|p1] x f0, f1, f2 and [p2|0] x f0
    const uint32_t dw = D.tmem + TM_DW + 64 * h;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const uint32_t win = (t == 3) ? 2 * PIECE : 0, fp = (t == 3) ? 0 : t * PIECE;
#pragma unroll
        for (int s = 0; s < 8; ++s) mma_bf16_ss(dw, dadd(D.DZ_mn, win + s * 256), dadd(D.F_mn, fp + s * 256), id_w, !(first_tile && t == 0 && s == 0));
    }
}
__device__ __forceinline__ void issue_trunk_wgrad(const TcDesc &D, bool first_tile) {
    constexpr uint32_t id = idesc_bf16(128, 16, 1, 1);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const uint32_t win = (t == 3) ? 2 * PIECE : 0, xp = (t == 3) ? 0 : t * XPIECE;
#pragma unroll
        for (int s = 0; s < 8; ++s)
            mma_bf16_ss(D.tmem + TM_DW0, dadd(D.DZ_mn, win + s * 256), dadd(D.X_mn, xp + s * 256), id, !(first_tile && t == 0 && s == 0));
    }
}

__device__ __forceinline__ void tmem_ld16w(uint32_t taddr, float (&v)[TC_W]) {
    tmem_ld16(taddr, v);
    tmem_ld_wait();
}

// Fused optimiser tail (single-GPU path): after a grid-wide barrier every CTA reduces its slice of the parameters over
// all CTAs' partial gradients (fixed order), a second barrier makes the squared-norm partials visible, then every CTA
// applies clip_grad_norm_ + AdamW to its slice.  One launch per optimiser step instead of three.
struct TcOptimizer {
    float *params_rw, *grad, *m, *v;          // params_rw == nullptr: gradient only (the reduction runs as a separate kernel)
    int64_t *clock;                           // {int64 step, double beta1^step, double beta2^step}
    unsigned int *sync;                       // {arrival count, generation} of the grid barrier
    double *sumsq;                            // [grid] squared-norm partials (unused by the fused step: tagged words instead)
    unsigned long long *tagw;                 // [grid][8] tagged words: loss sums (3 doubles = 6 words), squared norm (2 words)
    double *norm_out;
    float lr, wd, max_norm;
    double *loss_out;                         // 4 doubles, accumulated
    double rows;
    // sharded runs: the gradient exchange happens inside this kernel over NVLink peer memory instead of an NCCL allreduce.
    // peers[r] = base of rank r's exchange buffer: inbox[2 (step parity)][world (sender)][gstride] words {float bits, step number}
    float *const *peers;
    int rank, world, gstride;
};

// one-shot cross-GPU exchange helpers (system-scope release / acquire on the flag words in peer memory)
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// tagged words inside one GPU: {32 payload bits, step number}; a naturally aligned 64-bit access is single-copy atomic, so a
// reader that sees this launch's step number also sees the payload - no fence, no barrier between producer and consumer
__device__ __forceinline__ void st_tag(unsigned long long *p, unsigned int payload, unsigned int epoch) {
    const unsigned long long v = ((unsigned long long)epoch << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_tag(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_tag_double(unsigned long long *p2, double x, unsigned int epoch) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    st_tag(p2, (unsigned int)(b >> 32), epoch);
    st_tag(p2 + 1, (unsigned int)b, epoch);
}
// polls until both halves carry `epoch` (bounded); ok is cleared on a time-out
__device__ __forceinline__ double ld_tag_double(const unsigned long long *p2, unsigned int epoch, bool &ok) {
    unsigned long long hi = 0, lo = 0;
    bool got = false;
    for (int it = 0; it < (1 << 22) && !got; ++it) {
        hi = ld_tag(p2); lo = ld_tag(p2 + 1);
        got = (unsigned int)(hi >> 32) == epoch && (unsigned int)(lo >> 32) == epoch;
    }
    ok = ok && got;
    return __longlong_as_double((long long)(((hi & 0xffffffffull) << 32) | (lo & 0xffffffffull)));
}
__device__ __forceinline__ float ld_relaxed_sys(const float *p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// grid-wide barrier for a grid whose CTAs are all resident (cooperative launch): bounded spin, false on time-out.
// bar = {arrival count, generation}, both zero once (workspace allocation); the last CTA to arrive resets the count and
// advances the generation, so the same pair serves every barrier of every launch with no host-side reset in between.
__device__ __forceinline__ bool grid_barrier(unsigned int *bar, unsigned int nblocks) {
    __syncthreads();
    __shared__ int ok_s;
    if (threadIdx.x == 0) {
        unsigned int gen, cur;
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");   // cannot advance before this CTA arrives
        __threadfence();
        int ok = 1;
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            atomicExch(bar, 0u);
            __threadfence();
            asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1) : "memory");
        } else {
            ok = 0;
            for (int it = 0; it < (1 << 24); ++it) {
                asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(cur) : "l"(bar + 1) : "memory");
                if (cur != gen) { ok = 1; break; }
            }
        }
        ok_s = ok;
    }
    __syncthreads();
    return ok_s != 0;
}

// sum over the CTAs' partial rows of parameter i, slice sl of RED_SL (blocks sl, sl + RED_SL, ... in ascending order)
constexpr int RED_SL = 8, RED_MAX = 19;   // 8 x 19 >= 148 CTAs: every word of a slice is in flight at once (one L2 round trip, one pass of the threads)
__device__ __forceinline__ float reduce_slice(const float *__restrict__ partials, int nblocks, int stride, int i, int sl) {
    float s = 0.f;
    for (int base = sl; base < nblocks; base += RED_SL * RED_MAX) {
        float v[RED_MAX];
#pragma unroll
        for (int u = 0; u < RED_MAX; ++u) {
            const int bl = base + u * RED_SL;
            v[u] = bl < nblocks ? __ldcg(partials + (size_t)bl * stride + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < RED_MAX; ++u) s += v[u];
    }
    return s;
}
// the same over TAGGED partial rows (fused step): every word is polled until it carries this launch's step number, so the
// reduction needs no grid barrier behind the producers; the summation order is the one above (bit-identical result)
__device__ __forceinline__ float reduce_slice_tagged(const unsigned long long *__restrict__ partials, int nblocks, int stride, int i, int sl,
                                                     unsigned int epoch, bool &ok) {
    float s = 0.f;
    for (int base = sl; base < nblocks; base += RED_SL * RED_MAX) {
        unsigned long long v[RED_MAX];
        bool all = false;
        for (int it = 0; it < (1 << 22) && !all; ++it) {
            if (it) __nanosleep(400);   // the poll of a whole grid is ~11 MB of L2 reads per round: leave the producers their bandwidth
            all = true;
#pragma unroll
            for (int u = 0; u < RED_MAX; ++u) {
                const int bl = base + u * RED_SL;
                v[u] = bl < nblocks ? ld_tag(partials + (size_t)bl * stride + i) : ((unsigned long long)epoch << 32);
                all = all && (unsigned int)(v[u] >> 32) == epoch;
            }
        }
        ok = ok && all;
#pragma unroll
        for (int u = 0; u < RED_MAX; ++u) s += __uint_as_float((unsigned int)v[u]);
    }
    return s;
}
// loss_out[0..2] += sum over the CTAs' loss partials (lane l adds blocks l, l + 32, ... in ascending order, then a fixed
// shuffle tree: bit-reproducible), loss_out[3] += rows.  One full warp.
__device__ __forceinline__ void add_loss_sums(const double *loss_partials, int nblocks, double *loss_out, double rows, int lane) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int bl = lane; bl < nblocks; bl += 32) {
        a0 += __ldcg(loss_partials + bl * 4); a1 += __ldcg(loss_partials + bl * 4 + 1); a2 += __ldcg(loss_partials + bl * 4 + 2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) { loss_out[0] += a0; loss_out[1] += a1; loss_out[2] += a2; loss_out[3] += rows; }
}
__device__ __forceinline__ float reduce_tree(float (&t)[RED_SL]) {
#pragma unroll
    for (int w = RED_SL / 2; w > 0; w >>= 1)
#pragma unroll
        for (int u = 0; u < w; ++u) t[u] += t[u + w];
    return t[0];
}

// optional phase timestamps of CTA 0 (PRL_TC_TIMING=1 in the environment prints them after the launch; debugging aid)
__device__ long long g_tc_clock[32];
#define TC_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_tc_clock[i] = clock64(); } while (0)
// per-CTA wall-clock marks (ns): [cta][0] first instruction, [1] tiles done, [2] partials written, [3] last instruction
__device__ unsigned long long g_tc_span[160 * 4 + 8];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
#define TC_SPAN0(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_tc_span[640 + (i)] = globaltimer_ns(); } while (0)
#define TC_SPAN(i) do { if (threadIdx.x == 0 && blockIdx.x < 160) g_tc_span[blockIdx.x * 4 + (i)] = globaltimer_ns(); } while (0)

// ===================================================================================================== the kernel
// NA = compile-time bound of the output widths (action_dim rounded up to 2, 4 or 8): the per-output loops unroll over it
template <int NA>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_ppo_grad_tc(const float *__restrict__ params, PolicyLayout L, const float *__restrict__ states, const float *__restrict__ actions,
              const float *__restrict__ old_logp, const float *__restrict__ adv, const float *__restrict__ returns, int64_t b,
              float clip, float inv_count, float *__restrict__ partials, int part_stride, double *__restrict__ loss_partials,
              int *__restrict__ status, TcOptimizer opt, int qpc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bars[6];          // MMA completion: actor forward, actor backward, critic dgrad, trunk wgrad, critic wgrad, critic forward
    __shared__ uint64_t sbar[4];          // operands staged (512 arrivals): F + X, actor DZ, critic DZ, trunk DZ
    __shared__ uint32_t tmem_slot;
    __shared__ double red[32];
    __shared__ float b2s[4][2][TC_MAX_A];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_mma_warp = warp == TC_COMPUTE / 32;
    const int rq = warp & 3, q = (warp >> 2) & 3;   // row quarter (= tensor-memory lane quarter), feature quarter
    const int r = rq * 32 + lane, j0 = TC_W * q;    // row of the tile, first feature of this thread
    const int O = L.O, A = L.A;
    TC_STAMP(0);
    TC_SPAN(0);

    // ---- carve shared memory (all pointers derive from smem_raw so they stay in the shared state space)
    unsigned char *sW = smem_raw, *sF = sW + 3 * PIECE, *sDZ = sF + 3 * PIECE, *sX = sDZ + 4 * PIECE;
    float *sSmall = reinterpret_cast<float *>(sX + 3 * XPIECE);
    float *s_w0t = sSmall, *s_g0w = s_w0t + O * HID, *s_g0b = s_g0w + HID;
    float *s_head0 = s_g0b + HID;                                   // per head: gw[64] gb[64] w2[out][64] b2[round4(out)]
    const int head0_floats = 2 * HID + L.head[0].out * HID + round4(L.head[0].out);
    float *s_head1 = s_head0 + head0_floats;
    float *s_po = s_head1 + 2 * HID + L.head[1].out * HID + round4(L.head[1].out);   // [4 q][128 r][NA] partial head outputs
    float *s_red = s_po + 4 * TC_ROWS * NA;                                            // [4 rq][NQ][64] final combine

    // ---- stage parameters: fp32 small ones; W1 of both heads as bf16x3 in the operand layout (row n = h*64 + j).
    // Every global load is issued before the first dependent shared-memory store (one exposed memory latency, not ten).
    {
        float v[TC_W];
        const int n_small0 = HID * O, n_small1 = n_small0 + 2 * HID, n_h0 = 2 * HID + L.head[0].out * HID + L.head[0].out,
                  n_h1 = 2 * HID + L.head[1].out * HID + L.head[1].out, n_small = n_small1 + n_h0 + n_h1;
        if (!is_mma_warp) {
            // thread (r, q) stages features 16q..16q+15 of row n = r of the stacked [128][64] W1 (n < 64: actor, else critic)
            const float4 *wrow = reinterpret_cast<const float4 *>(params + L.head[r >> 6].w1 + (r & 63) * HID + j0);
            if ((L.head[r >> 6].w1 & 3) == 0) {   // warp-uniform (a warp's rows belong to one head)
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float4 t = __ldg(wrow + k); v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w; }
            } else {
                const float *ws = reinterpret_cast<const float *>(wrow);
#pragma unroll
                for (int k = 0; k < TC_W; ++k) v[k] = __ldg(ws + k);
            }
        }
        // small parameters: w0 (stored transposed), then three blocks that are contiguous both in `params` and in shared
        // memory: {g0w, g0b}, head 0 {gw, gb, w2, b2}, head 1 {gw, gb, w2, b2}
        for (int base = 0; base < n_small; base += 2 * TC_THREADS) {
            float sv[2];
            float *dst[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = base + u * TC_THREADS + tid;
                dst[u] = nullptr;
                if (i < n_small0) { const int j = i / O, o = i - j * O; dst[u] = s_w0t + o * HID + j; sv[u] = __ldg(params + L.w0 + i); }
                else if (i < n_small1) { dst[u] = s_g0w + (i - n_small0); sv[u] = __ldg(params + L.g0w + (i - n_small0)); }
                else if (i < n_small1 + n_h0) { dst[u] = s_head0 + (i - n_small1); sv[u] = __ldg(params + L.head[0].gw + (i - n_small1)); }
                else if (i < n_small) { dst[u] = s_head1 + (i - n_small1 - n_h0); sv[u] = __ldg(params + L.head[1].gw + (i - n_small1 - n_h0)); }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (dst[u]) *dst[u] = sv[u];
        }
        if (!is_mma_warp) {
            store_pieces16(sW, r, q, v);
            // zero slot behind the three DZ pieces; X pieces (columns >= O stay zero for the whole kernel)
#pragma unroll
            for (int c = 0; c < 2; ++c) *reinterpret_cast<uint4 *>(sDZ + 3 * PIECE + (2 * q + c) * CHUNK + r * 16) = make_uint4(0, 0, 0, 0);
            if (q == 0)
                for (int c = 0; c < 6; ++c) *reinterpret_cast<uint4 *>(sX + c * CHUNK + r * 16) = make_uint4(0, 0, 0, 0);
        }
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&sbar[i], TC_COMPUTE);
        for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    if (is_mma_warp) tmem_alloc(&tmem_slot, TM_COLS);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    TC_STAMP(1);
    TcDesc D;
    D.tmem = tmem_slot;
    D.F_k = smem_desc(smem_u32(sF), CHUNK, 128);
    D.W_k = smem_desc(smem_u32(sW), CHUNK, 128);
    D.DZ_k = smem_desc(smem_u32(sDZ), CHUNK, 128);
    D.W_mn = smem_desc(smem_u32(sW), 128, CHUNK);
    D.DZ_mn = smem_desc(smem_u32(sDZ), 128, CHUNK);
    D.F_mn = smem_desc(smem_u32(sF), 128, CHUNK);
    D.X_mn = smem_desc(smem_u32(sX), 128, CHUNK);
    const uint32_t lane_base = D.tmem + ((uint32_t)(rq * 32) << 16);
    bool mma_ok = true;
    // this CTA's rows: quarters [blockIdx.x * qpc, + myq), tile t = quarters 4t .. 4t+3 of them
    const int64_t row_base = (int64_t)blockIdx.x * qpc * 32;
    const int myq = (int)max((int64_t)0, min((int64_t)qpc, (b + 31) / 32 - (int64_t)blockIdx.x * qpc));
    const int ntiles = (myq + 3) >> 2;
    uint32_t it = 0;
    double l_pol = 0.0, l_val = 0.0, l_ent = 0.0;
    const int P = L.total;
    float *part = partials + (size_t)blockIdx.x * part_stride;
    // fused optimiser step: this CTA's partial row is written as tagged 64-bit words (see st_tag)
    const bool fused = opt.params_rw != nullptr;
    unsigned long long *part64 = reinterpret_cast<unsigned long long *>(partials) + (size_t)blockIdx.x * part_stride;
    // optimiser clock, read before anybody can advance it (CTA 0 does, after the second grid barrier)
    int64_t opt_step = 0;
    double opt_p1 = 0.0, opt_p2 = 0.0;
    if (opt.params_rw) {
        opt_step = opt.clock[0] + 1;
        const double *pw = reinterpret_cast<const double *>(opt.clock) + 1;
        const bool have = opt_step > 1 && pw[0] > 0.0;
        opt_p1 = have ? pw[0] * 0.9 : pow(0.9, (double)opt_step);
        opt_p2 = have ? pw[1] * 0.999 : pow(0.999, (double)opt_step);
    }
    // the tag of every word this launch publishes inside the GPU: a launch counter kept in the workspace header (word 3),
    // read by everybody here and advanced by CTA 0 at the very end - monotonic per workspace, whatever optimiser uses it
    const unsigned int epoch = fused ? __ldcg(reinterpret_cast<const unsigned int *>(status) + 3) + 1u : 0u;
    // one partial-gradient entry: plain float (separate reduction kernel) or tagged word (fused step)
    auto put = [&](int idx, float v) {
        if (fused) st_tag(part64 + idx, __float_as_uint(v), epoch);
        else part[idx] = v;
    };

    if (is_mma_warp) {
        // =============================================================================== MMA-issue warp
        for (int tile = 0; tile < ntiles; ++tile, ++it) {
            const uint32_t parity = it & 1;
            mma_ok &= mbar_wait(&sbar[0], parity);
            fence_after_sync();
            if (elect_one()) { issue_forward(D, 0); mma_commit(&bars[0]); issue_forward(D, 1); mma_commit(&bars[5]); }
            __syncwarp();
            mma_ok &= mbar_wait(&sbar[1], parity);
            fence_after_sync();
            if (elect_one()) { issue_head_dgrad(D, 0); issue_head_wgrad(D, 0, it == 0); mma_commit(&bars[1]); }
            __syncwarp();
            mma_ok &= mbar_wait(&sbar[2], parity);
            fence_after_sync();
            // the trunk backward needs DF (both dgrads) only: the critic's wgrad gets its own completion barrier
            if (elect_one()) { issue_head_dgrad(D, 1); mma_commit(&bars[2]); issue_head_wgrad(D, 1, it == 0); mma_commit(&bars[4]); }
            __syncwarp();
            mma_ok &= mbar_wait(&sbar[3], parity);
            fence_after_sync();
            if (elect_one()) { issue_trunk_wgrad(D, it == 0); mma_commit(&bars[3]); }
            __syncwarp();
        }
    } else {
        // =============================================================================== compute warps
        // column-sum accumulators live in shared memory: s_red[rq][quantity][64 features]; this warp owns features j0..j0+15
        // of row quarter rq.  Quantities: 0 dgamma0, 1 dbeta0, then per head: dgamma, dbeta, dW2[a] (a < out).
        const int NQ = tc_num_q(L);
        float *acc = s_red + (size_t)rq * NQ * HID + j0;
        const int qh[2] = {2, 2 + 2 + L.head[0].out};
        for (int i = lane; i < NQ * TC_W; i += 32) acc[(i / TC_W) * HID + (i % TC_W)] = 0.f;
        if (q == 0 && lane < 2 * TC_MAX_A) (&b2s[rq][0][0])[lane] = 0.f;
        __syncwarp();

        // inputs of the first tile; inside the loop the next tile's are prefetched while the current one computes
        constexpr int XR = 8;   // observation values kept in registers (observ_dim > 8 reads the rest on demand)
        float xn[XR];
        {
            const int64_t row0 = row_base + r;
#pragma unroll
            for (int i = 0; i < XR; ++i) xn[i] = (rq < myq && row0 < b && i < O) ? __ldg(states + row0 * O + i) : 0.f;
        }
        for (int tile = 0; tile < ntiles; ++tile, ++it) {
            const uint32_t parity = it & 1;
            const int64_t row = row_base + (int64_t)tile * TC_ROWS + r;
            if (4 * tile + rq >= myq) {
                // this warp's quarter lies past the CTA's rows (last tile only): zero operand rows, keep every hand-off
                if (it > 0) mma_ok &= mbar_wait(&bars[3], parity ^ 1);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int pc = 0; pc < 3; ++pc) {
                        *reinterpret_cast<uint4 *>(sF + pc * PIECE + (2 * q + c) * CHUNK + r * 16) = make_uint4(0, 0, 0, 0);
                        *reinterpret_cast<uint4 *>(sDZ + pc * PIECE + (2 * q + c) * CHUNK + r * 16) = make_uint4(0, 0, 0, 0);
                    }
                if (q == 0)
                    for (int c = 0; c < 6; ++c) *reinterpret_cast<uint4 *>(sX + c * CHUNK + r * 16) = make_uint4(0, 0, 0, 0);
                staged(&sbar[0]);
                bar_compute();       // the live warps' partial-output exchange, actor
                staged(&sbar[1]);
                bar_compute();       // critic
                staged(&sbar[2]);
                staged(&sbar[3]);
                continue;
            }
            const bool live = row < b;
            const float *xrow = states + (live ? row : 0) * O;   // dead rows read row 0 and are masked
            const float xmask = live ? 1.f : 0.f;
            float x[XR];
#pragma unroll
            for (int i = 0; i < XR; ++i) x[i] = xn[i];
            {
                const int64_t rown = row + TC_ROWS;
                const bool nlive = 4 * (tile + 1) + rq < myq && rown < b;
#pragma unroll
                for (int i = 0; i < XR; ++i) xn[i] = (nlive && i < O) ? __ldg(states + rown * O + i) : 0.f;
            }
            const float adv_i = live ? __ldg(adv + row) : 0.f, old_i = live ? __ldg(old_logp + row) : 0.f;
            const float ret_i = live ? __ldg(returns + row) : 0.f;
            const int act = live ? (int)__ldg(actions + row) : 0;

            // ================= trunk forward (CUDA cores): z0 = W0 x, GroupNorm, SiLU -> F pieces, X pieces
            auto trunk_pre = [&](float (&zh0)[TC_W], float (&rs0)[2]) {
#pragma unroll
                for (int j = 0; j < TC_W; ++j) zh0[j] = 0.f;
#pragma unroll
                for (int i = 0; i < XR; ++i) {
                    if (i < O) {
                        float w[TC_W];
                        load16(s_w0t + i * HID + j0, w);
#pragma unroll
                        for (int j = 0; j < TC_W; ++j) zh0[j] = fmaf(x[i], w[j], zh0[j]);
                    }
                }
                for (int i = XR; i < O; ++i) {
                    const float xi = xmask * __ldg(xrow + i);
                    float w[TC_W];
                    load16(s_w0t + i * HID + j0, w);
#pragma unroll
                    for (int j = 0; j < TC_W; ++j) zh0[j] = fmaf(xi, w[j], zh0[j]);
                }
                gn_normalize16(zh0, rs0);
            };
            {
                float zh0[TC_W], rs0[2];
                trunk_pre(zh0, rs0);
                float f[TC_W], g0w[TC_W], g0b[TC_W];
                load16(s_g0w + j0, g0w);
                load16(s_g0b + j0, g0b);
#pragma unroll
                for (int j = 0; j < TC_W; ++j) {
                    const float y = fmaf(zh0[j], g0w[j], g0b[j]);
                    f[j] = y * fast_sigmoid(y);
                }
                // the previous tile's trunk-wgrad MMAs read X and DZ; its head MMAs (already waited for) read F
                if (it > 0) mma_ok &= mbar_wait(&bars[3], parity ^ 1);
                store_pieces16(sF, r, q, f);
                if (q == 0) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t q0[4], q1[4], q2[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int e = 8 * c + 2 * i;
                            float xa, xb;
                            if (c == 0) { xa = x[e & 7]; xb = x[(e + 1) & 7]; }
                            else { xa = e < O ? xmask * __ldg(xrow + e) : 0.f; xb = e + 1 < O ? xmask * __ldg(xrow + e + 1) : 0.f; }
                            split_bf16x3(xa, xb, q0[i], q1[i], q2[i]);
                        }
                        unsigned char *p = sX + c * CHUNK + r * 16;
                        *reinterpret_cast<uint4 *>(p) = make_uint4(q0[0], q0[1], q0[2], q0[3]);
                        *reinterpret_cast<uint4 *>(p + XPIECE) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
                        *reinterpret_cast<uint4 *>(p + 2 * XPIECE) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
                    }
                }
            }
            staged(&sbar[0]);
            if (it == 0) TC_STAMP(2);
            mma_ok &= mbar_wait(&bars[0], parity);
            fence_after_sync();
            if (it == 0) TC_STAMP(4);

            // ================= heads: forward epilogue, loss, backward epilogue -> DZ pieces, tensor-core dgrad + wgrad
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int nout = L.head[h].out;
                const float *sh = h ? s_head1 : s_head0;
                const float *w2 = sh + 2 * HID;
                float gw[TC_W], gb[TC_W];
                load16(sh + j0, gw);
                load16(sh + HID + j0, gb);
                float zhat[TC_W], rstd[2], sg[TC_W];
                if (h == 1) { mma_ok &= mbar_wait(&bars[5], parity); fence_after_sync(); }   // critic forward complete
                tmem_ld16w(lane_base + TM_Z + 64 * h + j0, zhat);
                gn_normalize16(zhat, rstd);
                float po[NA];
#pragma unroll
                for (int a = 0; a < NA; ++a) po[a] = 0.f;
#pragma unroll
                for (int j = 0; j < TC_W; ++j) {
                    const float y = fmaf(zhat[j], gw[j], gb[j]);
                    sg[j] = fast_sigmoid(y);
                    const float hj = y * sg[j];
#pragma unroll
                    for (int a = 0; a < NA; ++a)
                        if (a < nout) po[a] = fmaf(hj, w2[a * HID + j0 + j], po[a]);
                }
#pragma unroll
                for (int a = 0; a < NA; ++a) s_po[(q * TC_ROWS + r) * NA + a] = po[a];
                bar_compute();
                float out[NA];
#pragma unroll
                for (int a = 0; a < NA; ++a)
                    out[a] = (a < nout) ? w2[nout * HID + a] + ((s_po[(0 * TC_ROWS + r) * NA + a] + s_po[(1 * TC_ROWS + r) * NA + a]) +
                                                              (s_po[(2 * TC_ROWS + r) * NA + a] + s_po[(3 * TC_ROWS + r) * NA + a]))
                                        : 0.f;
                // ---- loss and output gradients (the four threads of a row compute them redundantly; q == 0 keeps the sums)
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
                        const float rr = expf(fminf(fmaxf(dl, -20.f), 20.f));
                        const float s1 = rr * adv_i;
                        const float s2 = fminf(fmaxf(rr, 1.0f - clip), 1.0f + clip) * adv_i;
                        const float g1 = s1 < s2 ? 1.f : (s1 > s2 ? 0.f : 0.5f);   // torch.min splits ties evenly
                        const float in_clip = (rr >= 1.0f - clip && rr <= 1.0f + clip) ? 1.f : 0.f;
                        const float in20 = (dl >= -20.f && dl <= 20.f) ? 1.f : 0.f;
                        float dlogp = -inv_count * adv_i * (g1 + (1.f - g1) * in_clip) * rr * in20;
                        if (!(pa >= F32_EPS && pa <= 1.0f - F32_EPS)) dlogp = 0.f;   // clamp in probs_to_logits blocks the gradient
#pragma unroll
                        for (int a = 0; a < NA; ++a)
                            if (a < A) dout[a] = dlogp * ((a == act ? 1.f : 0.f) - p[a]);
                        if (q == 0) { l_pol += -fminf(s1, s2); l_ent += ent; }
                    }
                } else if (live) {
                    const float dv = out[0] - ret_i, ad = fabsf(dv);
                    if (q == 0) l_val += ad < 1.f ? 0.5f * dv * dv : ad - 0.5f;
                    dout[0] = 0.5f * inv_count * (ad < 1.f ? dv : (dv > 0.f ? 1.f : -1.f));
                }
                // ---- output-layer gradients: dW2[a][j] = sum_r dout[a] h_j, db2[a] = sum_r dout[a]
#pragma unroll
                for (int a = 0; a < NA; ++a) {
                    if (a < nout) {
                        float t[TC_W];
#pragma unroll
                        for (int j = 0; j < TC_W; ++j) t[j] = dout[a] * (fmaf(zhat[j], gw[j], gb[j]) * sg[j]);
                        colsum16(t, acc + (qh[h] + 2 + a) * HID);
                        if (q == 0) {
                            const float sb = warp_sum(dout[a]);
                            if (lane == 0) b2s[rq][h][a] += sb;
                        }
                    }
                }
                // ---- dy (in place of sg), GroupNorm-affine gradients, GroupNorm backward -> dz
#pragma unroll
                for (int j = 0; j < TC_W; ++j) {
                    float dh = 0.f;
#pragma unroll
                    for (int a = 0; a < NA; ++a)
                        if (a < nout) dh = fmaf(dout[a], w2[a * HID + j0 + j], dh);
                    const float y = fmaf(zhat[j], gw[j], gb[j]);
                    sg[j] = dh * sg[j] * fmaf(y, 1.0f - sg[j], 1.0f);
                }
                {
                    float t[TC_W];
#pragma unroll
                    for (int j = 0; j < TC_W; ++j) t[j] = sg[j] * zhat[j];
                    colsum16(t, acc + qh[h] * HID);
                    colsum16(sg, acc + (qh[h] + 1) * HID);
                }
                gn_backward16(sg, zhat, rstd, gw);   // sg now holds dz
                if (it == 0) TC_STAMP(5 + 3 * h);
                // the actor's MMAs read DZ: they must have completed before the critic overwrites it
                if (h == 1) mma_ok &= mbar_wait(&bars[1], parity);
                store_pieces16(sDZ, r, q, sg);
                staged(&sbar[1 + h]);
                if (it == 0) TC_STAMP(6 + 3 * h);
            }

            // ================= trunk backward: DF -> dy0 -> GroupNorm backward -> DZ pieces, tensor-core wgrad against X
            {
                float df[TC_W], g0w[TC_W], g0b[TC_W], zh0[TC_W], rs0[2];
                trunk_pre(zh0, rs0);   // recomputed (cheap) rather than kept in registers across the head phases
                load16(s_g0w + j0, g0w);
                load16(s_g0b + j0, g0b);
                mma_ok &= mbar_wait(&bars[2], parity);    // critic dgrad complete -> DF final
                fence_after_sync();
                if (it == 0) TC_STAMP(11);
                tmem_ld16w(lane_base + TM_DF + j0, df);
#pragma unroll
                for (int j = 0; j < TC_W; ++j) {
                    const float y = fmaf(zh0[j], g0w[j], g0b[j]);
                    const float s = fast_sigmoid(y);
                    df[j] = df[j] * s * fmaf(y, 1.0f - s, 1.0f);
                }
                {
                    float t[TC_W];
#pragma unroll
                    for (int j = 0; j < TC_W; ++j) t[j] = df[j] * zh0[j];
                    colsum16(t, acc);
                    colsum16(df, acc + HID);
                }
                gn_backward16(df, zh0, rs0, g0w);
                mma_ok &= mbar_wait(&bars[4], parity);    // critic wgrad complete -> DZ free again
                store_pieces16(sDZ, r, q, df);
                staged(&sbar[3]);
                if (it == 0) TC_STAMP(12);
            }
        }

        // ================= read the accumulators out: tensor memory -> this block's partial-gradient row
        TC_STAMP(13);
        TC_SPAN(1);
        if (it > 0) mma_ok &= mbar_wait(&bars[3], (it - 1) & 1);
        fence_after_sync();
        TC_STAMP(14);
        // DW_h: lanes 0..63 hold the first piece window's products, lanes 64..127 the second's; thread (r, q) reads 16 columns
        // of both heads and of DW0, parks them in shared memory (F and DZ regions, free now), then the CTA writes
        // dW = first + second with coalesced stores.  One pass: all three tensor-memory loads in flight, two barriers.
        constexpr int SS = HID + 4;
        float *scratch0 = reinterpret_cast<float *>(sF);              // [128][68] fp32 = 34 KB inside the 48 KB F region
        float *scratch1 = reinterpret_cast<float *>(sDZ);             // same, inside the 64 KB DZ region
        float *scratch2 = scratch1 + TC_ROWS * SS;                    // [128][17] for DW0
        {
            float v0[TC_W], v1[TC_W], v2[TC_W];
            tmem_ld16(lane_base + TM_DW + j0, v0);
            tmem_ld16(lane_base + TM_DW + 64 + j0, v1);
            tmem_ld16(lane_base + TM_DW0, v2);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                *reinterpret_cast<float4 *>(scratch0 + r * SS + j0 + 4 * k) = make_float4(v0[4 * k], v0[4 * k + 1], v0[4 * k + 2], v0[4 * k + 3]);
                *reinterpret_cast<float4 *>(scratch1 + r * SS + j0 + 4 * k) = make_float4(v1[4 * k], v1[4 * k + 1], v1[4 * k + 2], v1[4 * k + 3]);
            }
            if (q == 0) {
#pragma unroll
                for (int i = 0; i < TC_W; ++i) scratch2[r * 17 + i] = v2[i];
            }
            bar_compute();
            for (int idx = tid; idx < HID * HID; idx += TC_COMPUTE) {
                const int j = idx >> 6, k = idx & 63;
                put(L.head[0].w1 + idx, (it > 0) ? scratch0[j * SS + k] + scratch0[(64 + j) * SS + k] : 0.f);
                put(L.head[1].w1 + idx, (it > 0) ? scratch1[j * SS + k] + scratch1[(64 + j) * SS + k] : 0.f);
            }
            for (int idx = tid; idx < HID * O; idx += TC_COMPUTE) {
                const int j = idx / O, i = idx - j * O;
                put(L.w0 + idx, (it > 0) ? scratch2[j * 17 + i] + scratch2[(64 + j) * 17 + i] : 0.f);
            }
        }
        // column-sum accumulators (already in s_red / b2s): combine the four row quarters
        {
            bar_compute();
            for (int idx = tid; idx < NQ * HID; idx += TC_COMPUTE) {
                const int qq = idx / HID, j = idx - qq * HID;
                const float sm = (s_red[(0 * NQ + qq) * HID + j] + s_red[(1 * NQ + qq) * HID + j]) + (s_red[(2 * NQ + qq) * HID + j] + s_red[(3 * NQ + qq) * HID + j]);
                int off;
                if (qq == 0) off = L.g0w;
                else if (qq == 1) off = L.g0b;
                else {
                    int k = qq - 2, h = 0;
                    if (k >= 2 + L.head[0].out) { k -= 2 + L.head[0].out; h = 1; }
                    off = (k == 0) ? L.head[h].gw : (k == 1) ? L.head[h].gb : L.head[h].w2 + (k - 2) * HID;
                }
                put(off + j, sm);
            }
            if (tid < 2 * NA) {
                const int h = tid / NA, a = tid % NA;
                if (a < L.head[h].out) put(L.head[h].b2 + a, (b2s[0][h][a] + b2s[1][h][a]) + (b2s[2][h][a] + b2s[3][h][a]));
            }
        }
        TC_STAMP(15);
    }
    // everyone, the MMA warp included (block_sum synchronises the whole CTA)
    const double bp = block_sum<double>(l_pol, red);
    const double bv = block_sum<double>(l_val, red);
    const double be = block_sum<double>(l_ent, red);
    if (tid == 0) {
        if (fused) {
            st_tag_double(opt.tagw + blockIdx.x * 8 + 0, bp, epoch);
            st_tag_double(opt.tagw + blockIdx.x * 8 + 2, bv, epoch);
            st_tag_double(opt.tagw + blockIdx.x * 8 + 4, be, epoch);
        } else {
            loss_partials[blockIdx.x * 4 + 0] = bp;
            loss_partials[blockIdx.x * 4 + 1] = bv;
            loss_partials[blockIdx.x * 4 + 2] = be;
            loss_partials[blockIdx.x * 4 + 3] = 0.0;
        }
    }
    if (!mma_ok && lane == 0) atomicExch(status, 1);
    fence_before_sync();
    __syncthreads();
    if (is_mma_warp) tmem_dealloc(D.tmem, TM_COLS);
    TC_STAMP(16);
    TC_SPAN(2);
    TC_SPAN(3);
    if (!opt.params_rw) return;

    // ================= fused optimiser tail: reduce -> clip_grad_norm_ -> AdamW on this CTA's slices of the parameters.
    // No grid barrier anywhere: every CTA published its partial row, loss sums and (below) squared norm as TAGGED words, and
    // the consumers poll the words they need until they carry this launch's step number.  (All CTAs are co-resident -
    // cooperative launch - so the polls terminate; they are bounded anyway: status 2.)
    const int nb = gridDim.x;
    bool tags_ok = true;
    TC_SPAN0(0);
    // The parameters are cut into slices of RC = 64 (142 slices for P = 9 027); CTA c owns slices c, c + nb, ...  The cut does
    // not depend on the grid, so ranks whose minibatches have different row counts (different grids) agree on it.
    constexpr int RC = 64;                                            // parameters per slice
    const int nsl = (P + RC - 1) / RC;
    float *sl_part = reinterpret_cast<float *>(smem_raw);             // [RED_SL][RC] slice sums
    double *sq = reinterpret_cast<double *>(smem_raw + 8192);         // [RC] squared gradients
    double ssum = 0.0;                                                // thread 0: squared norm of this CTA's slices
    const bool sharded = opt.world > 1;
    // losses of the whole launch (one warp of CTA 0; the other CTAs' loss partials were written before the first grid barrier)
    if (blockIdx.x == 0 && is_mma_warp && opt.loss_out) {
        // lane l adds blocks l, l + 32, ... in ascending order, then a fixed shuffle tree: bit-reproducible
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int bl = lane; bl < nb; bl += 32) {
            a0 += ld_tag_double(opt.tagw + bl * 8 + 0, epoch, tags_ok);
            a1 += ld_tag_double(opt.tagw + bl * 8 + 2, epoch, tags_ok);
            a2 += ld_tag_double(opt.tagw + bl * 8 + 4, epoch, tags_ok);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) { opt.loss_out[0] += a0; opt.loss_out[1] += a1; opt.loss_out[2] += a2; opt.loss_out[3] += opt.rows; }
    }
    // sharded: exchange buffer of rank r = inbox[2 (step parity)][world (sender)][gstride] 8-byte words {float bits, step number}
    const unsigned int xepoch = (unsigned int)opt_step;   // the ranks agree on the optimiser step, not on launch counters
    const int inbox_off = (int)(opt_step & 1) * opt.world * opt.gstride;   // this step's inbox, in words
    __shared__ int peers_ok;
    if (tid == 0) peers_ok = 1;   // (made visible by the barriers inside the loop before anybody reads it)
    for (int sidx = blockIdx.x; sidx < nsl; sidx += nb) {
        const int p0 = sidx * RC, nc = min(RC, P - p0);
        for (int item = tid; item < nc * RED_SL; item += TC_THREADS) {
            const int sl = item / nc, pi = item - sl * nc;
            sl_part[sl * RC + pi] = reduce_slice_tagged(reinterpret_cast<const unsigned long long *>(partials), nb, part_stride, p0 + pi, sl, epoch, tags_ok);
        }
        __syncthreads();
        float gi = 0.f;
        if (tid < nc) {
            float t16[RED_SL];
#pragma unroll
            for (int u = 0; u < RED_SL; ++u) t16[u] = sl_part[u * RC + tid];
            gi = reduce_tree(t16);
        }
        if (sharded) {
            // ---- gradient exchange over peer memory, slice by slice, with no further grid-wide step and no separate flag:
            // every value travels as one 8-byte word {float bits, step number} (a naturally aligned 64-bit store is single-copy
            // atomic), PUSHED into every rank's inbox (slot = my rank); the receiver polls the words of its own inbox until
            // they carry this step's number and sums them in rank order (identical result on every rank) - the allreduce,
            // inside the kernel, one NVLink one-way latency long
            if (tid < nc) {
                const unsigned long long word = ((unsigned long long)xepoch << 32) | __float_as_uint(gi);
                for (int pr = 0; pr < opt.world; ++pr)
                    st_relaxed_sys_u64(reinterpret_cast<unsigned long long *>(opt.peers[pr]) + inbox_off + opt.rank * opt.gstride + p0 + tid, word);
                const unsigned long long *in = reinterpret_cast<const unsigned long long *>(opt.peers[opt.rank]) + inbox_off + p0 + tid;
                float g = 0.f;
                bool ok = true;
                for (int r0 = 0; r0 < opt.world; r0 += 8) {   // 8 ranks polled together, summed in rank order
                    unsigned long long pv[8];
                    bool all = false;
                    for (int itp = 0; itp < (1 << 22) && !all; ++itp) {
                        all = true;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            pv[u] = r0 + u < opt.world ? ld_relaxed_sys_u64(in + (r0 + u) * opt.gstride) : ((unsigned long long)xepoch << 32);
                            all = all && (unsigned int)(pv[u] >> 32) == xepoch;
                        }
                    }
                    ok = ok && all;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (r0 + u < opt.world) g += __uint_as_float((unsigned int)pv[u]);
                }
                if (!ok) peers_ok = 0;
                gi = g;
            }
            __syncthreads();
            if (!peers_ok) { if (tid == 0) atomicExch(status, 3); return; }
        }
        if (tid < nc) {
            opt.grad[p0 + tid] = gi;
            sq[tid] = (double)gi * gi;
        }
        __syncthreads();
        if (tid == 0)
            for (int k = 0; k < nc; ++k) ssum += sq[k];
        __syncthreads();
    }
    if (tid == 0) st_tag_double(opt.tagw + blockIdx.x * 8 + 6, ssum, epoch);
    TC_SPAN0(1);
    TC_SPAN0(2);
    __shared__ float coef_s;
    if (warp == 0) {
        // total squared norm in a fixed order: lane l adds the CTAs' (tagged) partials l, l + 32, ...; then a fixed shuffle tree
        double a = 0.0;
        for (int k = lane; k < nb; k += 32) a += ld_tag_double(opt.tagw + k * 8 + 6, epoch, tags_ok);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            const float total = (float)sqrt(a);
            coef_s = opt.max_norm > 0.f ? fminf(opt.max_norm / (total + 1e-6f), 1.0f) : 1.f;
            if (blockIdx.x == 0) {
                if (opt.norm_out) *opt.norm_out = (double)total;
                double *pw = reinterpret_cast<double *>(opt.clock) + 1;
                opt.clock[0] = opt_step; pw[0] = opt_p1; pw[1] = opt_p2;
                reinterpret_cast<unsigned int *>(status)[3] = epoch;   // the launch counter (every CTA read it at its start)
            }
        }
    }
    if (!tags_ok) atomicExch(status, 2);
    __syncthreads();
    TC_SPAN0(3);
    {
        const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
        const float step_size = (float)((double)opt.lr / (1.0 - opt_p1)), bc2_sqrt = (float)sqrt(1.0 - opt_p2);
        const float coef = coef_s, decay = 1.0f - opt.lr * opt.wd;
        for (int k = tid; k < RC * ((nsl - (int)blockIdx.x + nb - 1) / nb); k += TC_THREADS) {
            const int i = ((int)blockIdx.x + (k / RC) * nb) * RC + (k % RC);   // slice blockIdx.x + j nb, element k % RC
            if (i >= P) continue;
            const float g = opt.grad[i] * coef;      // written by this CTA above
            float pv = opt.params_rw[i] * decay;
            const float mi = opt.m[i] + (1.0f - b1) * (g - opt.m[i]);
            const float vi = fmaf(opt.v[i], b2, (1.0f - b2) * g * g);
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            pv = pv - step_size * (mi / denom);
            opt.params_rw[i] = pv; opt.m[i] = mi; opt.v[i] = vi;
        }
    }
    TC_SPAN(3);
}

// grad[i] = sum over blocks of partials[b][i] in a fixed order (bit-reproducible): 64 parameters x RED_SL block slices per
// CTA; slice sl adds blocks sl, sl + RED_SL, ... (RED_MAX independent loads in flight per thread), and the slice sums are
// combined in a fixed tree.  loss_out += block loss partials.  (Separate-kernel form of the fused
// tail above: used when the gradient is needed on its own, e.g. for the allreduce of the sharded path.)
__global__ void __launch_bounds__(64 * RED_SL)
k_reduce_partials_tc(const float *__restrict__ partials, int nblocks, int P, int stride, float *__restrict__ grad,
                     const double *__restrict__ loss_partials, double *__restrict__ loss_out, double rows) {
    __shared__ float part[RED_SL][64];
    const int p = threadIdx.x & 63, sl = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + p;
    part[sl][p] = i < P ? reduce_slice(partials, nblocks, stride, i, sl) : 0.f;
    __syncthreads();
    if (sl == 0 && i < P) {
        float t[RED_SL];
#pragma unroll
        for (int u = 0; u < RED_SL; ++u) t[u] = part[u][p];
        grad[i] = reduce_tree(t);
    }
    if (blockIdx.x == 0 && threadIdx.x < 32 && loss_out) add_loss_sums(loss_partials, nblocks, loss_out, rows, threadIdx.x);
}

// Row split.  The minibatch is cut into 32-row quarters (one per row-quarter warp group); CTA c owns the TC_QPC consecutive
// quarters from c * qpc on, i.e. floor(qpc / 4) full 128-row tiles and a last tile with qpc % 4 live quarters whose dead
// warps skip the CUDA-core work.  With qpc = ceil(quarters / SMs) every SM finishes within one quarter-tile of the others
// (65 536 rows on 148 SMs: 14 quarters = 3.5 tiles each, instead of 4 tiles on 68 CTAs and 3 on 80).
static void tc_split(int64_t b, int *grid, int *qpc) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t nq = (b + 31) / 32;
    int64_t q = (nq + sms - 1) / sms;
    if (q < 4) q = 4;   // small minibatches: whole tiles on fewer CTAs
    *qpc = (int)q;
    const int64_t gr = (nq + q - 1) / q;
    *grid = (int)(gr > 0 ? gr : 1);
}
// upper bound of the grid over every minibatch of at most b rows (the workspace is sized once for the largest one)
static int tc_grid_max(int64_t b) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t nt = (b + TC_ROWS - 1) / TC_ROWS;
    return (int)(nt < sms ? (nt > 0 ? nt : 1) : sms);
}
// header (4 words: status, 2 barrier words, launch counter) | partial rows (8-byte tagged words in the fused step, floats
// otherwise) | loss partials (4 doubles per CTA) | squared-norm partials (1 double) | tagged scalars (8 words per CTA)
static size_t tc_ws_floats(const PolicyLayout &L, int grid) {
    return 4 + (size_t)2 * grid * ((L.total + 3) & ~3) + (size_t)grid * 8 + (size_t)grid * 2 + (size_t)grid * 16 + 16;
}

}  // namespace prl

using namespace prl;

extern "C" {

int prl_ppo_grad_tc_supported(int is_continuous, int obs_dim, int action_dim) {
    return !is_continuous && obs_dim >= 1 && obs_dim <= TC_MAX_O && action_dim >= 1 && action_dim <= TC_MAX_A;
}

size_t prl_update_tc_ws_floats(int is_continuous, int obs_dim, int action_dim, int64_t batch) {
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    return tc_ws_floats(L, tc_grid_max(batch));
}

// shared launcher: gradient only (opt == nullptr: + separate reduction kernel) or fused optimiser step
static int launch_tc(const float *params, int is_continuous, int obs_dim, int action_dim, const float *states, const float *actions,
                     const float *old_logp, const float *adv, const float *returns, int64_t b, float policy_clip, float inv_count,
                     float *grad, double *loss_out, float *ws, size_t ws_floats, cudaStream_t st, const TcOptimizer *optp, const char *who) {
    PRL_REQUIRE(params && grad && ws && b >= 0 && (b > 0 || (optp && optp->world > 1)), "%s: bad arguments", who);
    PRL_REQUIRE(b == 0 || (states && actions && old_logp && adv && returns), "%s: null row pointer", who);
    PRL_REQUIRE(prl_ppo_grad_tc_supported(is_continuous, obs_dim, action_dim),
                "%s: only discrete policies with observ_dim <= %d and action_dim <= %d (got continuous=%d O=%d A=%d)", who, TC_MAX_O,
                TC_MAX_A, is_continuous, obs_dim, action_dim);
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    int grid, qpc;
    tc_split(b, &grid, &qpc);
    const int pstride = (L.total + 3) & ~3;   // per-CTA partial rows start 16-byte aligned
    PRL_REQUIRE(ws_floats >= tc_ws_floats(L, grid), "%s: workspace too small", who);
    PRL_REQUIRE(((uintptr_t)ws & 15) == 0, "%s: workspace must be 16-byte aligned", who);
    const int NA = action_dim <= 2 ? 2 : action_dim <= 4 ? 4 : 8;
    const size_t smem = tc_smem_bytes(L, NA);
    PRL_REQUIRE(smem <= 227 * 1024, "%s: needs %zu B shared memory (> 227 KB)", who, smem);
    // ws: [0] sticky status word, [1..2] grid-barrier counters (the caller zeroes the workspace once; the counters are
    // re-zeroed before every fused launch), then per-CTA partial gradients, loss partials, squared-norm partials
    int *status = reinterpret_cast<int *>(ws);
    float *partials = ws + 4;
    double *loss_partials = reinterpret_cast<double *>(partials + (size_t)2 * grid * pstride);
    TcOptimizer opt{};
    if (optp) {
        opt = *optp;
        opt.sync = reinterpret_cast<unsigned int *>(ws) + 1;
        opt.sumsq = loss_partials + (size_t)grid * 4;
        opt.tagw = reinterpret_cast<unsigned long long *>(opt.sumsq + grid);
    }
    auto launch = [&](auto kernel) -> int {
        PRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;   // the fused tail spins on grid barriers: every CTA must be resident
        attr[0].val.cooperative = optp ? 1 : 0;
        cfg.attrs = attr; cfg.numAttrs = 1;
        PRL_CUDA(cudaLaunchKernelEx(&cfg, kernel, params, L, states, actions, old_logp, adv, returns, b, policy_clip, inv_count, partials, pstride,
                                    loss_partials, status, opt, qpc));
        return PRL_OK;
    };
    const int rc = NA == 2 ? launch(k_ppo_grad_tc<2>) : NA == 4 ? launch(k_ppo_grad_tc<4>) : launch(k_ppo_grad_tc<8>);
    if (rc != PRL_OK) return rc;
    if (getenv("PRL_TC_TIMING")) {
        long long c[32];
        PRL_CUDA(cudaStreamSynchronize(st));
        PRL_CUDA(cudaMemcpyFromSymbol(c, g_tc_clock, sizeof c));
        fprintf(stderr, "[%s] CTA0 thread 0 cycles: setup %lld | tile 0: trunk fwd + stage F %lld, wait fwd MMA %lld, actor epilogue %lld, stage DZ %lld, "
                        "critic epilogue %lld, wait actor MMA + stage DZ %lld, wait critic MMA %lld, trunk bwd + stage %lld | all tiles %lld, final MMA wait %lld, "
                        "readout %lld, tail %lld | kernel %lld\n", who,
                c[1] - c[0], c[2] - c[1], c[4] - c[2], c[5] - c[4], c[6] - c[5], c[8] - c[6], c[9] - c[8], c[11] - c[9], c[12] - c[11], c[13] - c[1],
                c[14] - c[13], c[15] - c[14], c[16] - c[15], c[16] - c[0]);
        static unsigned long long sp[160 * 4 + 8];
        PRL_CUDA(cudaMemcpyFromSymbol(sp, g_tc_span, sizeof sp));
        unsigned long long t0 = ~0ull, s_max = 0, e1_min = ~0ull, e1_max = 0, e2_max = 0, e3_max = 0;
        for (int c2 = 0; c2 < grid && c2 < 160; ++c2) {
            t0 = sp[c2 * 4] < t0 ? sp[c2 * 4] : t0; s_max = sp[c2 * 4] > s_max ? sp[c2 * 4] : s_max;
            e1_min = sp[c2 * 4 + 1] < e1_min ? sp[c2 * 4 + 1] : e1_min; e1_max = sp[c2 * 4 + 1] > e1_max ? sp[c2 * 4 + 1] : e1_max;
            e2_max = sp[c2 * 4 + 2] > e2_max ? sp[c2 * 4 + 2] : e2_max; e3_max = sp[c2 * 4 + 3] > e3_max ? sp[c2 * 4 + 3] : e3_max;
        }
        fprintf(stderr, "[%s] grid %d wall clock (ns from the first CTA start): last CTA start %llu | tiles done: first %llu, last %llu | partials written %llu | "
                        "kernel end %llu | CTA0: start %llu tiles %llu end %llu\n", who, grid, s_max - t0, e1_min - t0, e1_max - t0, e2_max - t0, e3_max - t0,
                sp[0] - t0, sp[1] - t0, sp[3] - t0);
        if (optp) fprintf(stderr, "[%s] CTA0 tail: barrier 1 passed %llu, slice reduced %llu, barrier 2 passed %llu, norm known %llu, end %llu\n", who, sp[640] - t0,
                          sp[641] - t0, sp[642] - t0, sp[643] - t0, sp[3] - t0);
    }
    if (!optp) k_reduce_partials_tc<<<cdiv(L.total, 64), 64 * RED_SL, 0, st>>>(partials, grid, L.total, pstride, grad, loss_partials, loss_out, (double)b);
    return check_launch(who);
}

int prl_ppo_grad_tc(const float *params, int is_continuous, int obs_dim, int action_dim, const float *states, const float *actions,
                    const float *old_logp, const float *adv, const float *returns, int64_t b, float policy_clip, float inv_count,
                    float *grad, double *loss_out, float *ws, size_t ws_floats, void *stream) {
    return launch_tc(params, is_continuous, obs_dim, action_dim, states, actions, old_logp, adv, returns, b, policy_clip, inv_count, grad, loss_out,
                     ws, ws_floats, (cudaStream_t)stream, nullptr, "prl_ppo_grad_tc");
}

int prl_ppo_step_tc(float *params, int is_continuous, int obs_dim, int action_dim, const float *states, const float *actions,
                    const float *old_logp, const float *adv, const float *returns, int64_t b, float policy_clip, float inv_count,
                    float *grad, double *loss_out, float *exp_avg, float *exp_avg_sq, int64_t *step_counter, float lr, float weight_decay,
                    float max_norm, double *grad_norm_out, float *ws, size_t ws_floats, void *stream) {
    PRL_REQUIRE(exp_avg && exp_avg_sq && step_counter, "prl_ppo_step_tc: bad optimiser arguments");
    TcOptimizer opt{};
    opt.params_rw = params; opt.grad = grad; opt.m = exp_avg; opt.v = exp_avg_sq; opt.clock = step_counter; opt.norm_out = grad_norm_out;
    opt.lr = lr; opt.wd = weight_decay; opt.max_norm = max_norm; opt.loss_out = loss_out; opt.rows = (double)b;
    return launch_tc(params, is_continuous, obs_dim, action_dim, states, actions, old_logp, adv, returns, b, policy_clip, inv_count, grad, loss_out,
                     ws, ws_floats, (cudaStream_t)stream, &opt, "prl_ppo_step_tc");
}

int prl_ppo_step_tc_p2p(float *params, int is_continuous, int obs_dim, int action_dim, const float *states, const float *actions,
                        const float *old_logp, const float *adv, const float *returns, int64_t b, float policy_clip, float inv_count,
                        float *grad, double *loss_out, float *exp_avg, float *exp_avg_sq, int64_t *step_counter, float lr,
                        float weight_decay, float max_norm, double *grad_norm_out, void *const *peer_bufs, int rank, int world, float *ws,
                        size_t ws_floats, void *stream) {
    PRL_REQUIRE(exp_avg && exp_avg_sq && step_counter && peer_bufs && world >= 1 && world <= 64 && rank >= 0 && rank < world,
                "prl_ppo_step_tc_p2p: bad arguments");
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    TcOptimizer opt{};
    opt.params_rw = params; opt.grad = grad; opt.m = exp_avg; opt.v = exp_avg_sq; opt.clock = step_counter; opt.norm_out = grad_norm_out;
    opt.lr = lr; opt.wd = weight_decay; opt.max_norm = max_norm; opt.loss_out = loss_out; opt.rows = (double)b;
    opt.peers = reinterpret_cast<float *const *>(peer_bufs); opt.rank = rank; opt.world = world; opt.gstride = (L.total + 3) & ~3;
    return launch_tc(params, is_continuous, obs_dim, action_dim, states, actions, old_logp, adv, returns, b, policy_clip, inv_count, grad, loss_out,
                     ws, ws_floats, (cudaStream_t)stream, &opt, "prl_ppo_step_tc_p2p");
}

size_t prl_p2p_exchange_bytes(int is_continuous, int obs_dim, int action_dim, int world) {
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    const size_t gstride = (L.total + 3) & ~3;   // inbox[2][world][gstride] 8-byte words
    return (size_t)2 * world * gstride * 8 + 256;
}
int prl_p2p_alloc(size_t bytes, void **ptr) {
    PRL_REQUIRE(ptr && bytes > 0, "prl_p2p_alloc: bad arguments");
    PRL_CUDA(cudaMalloc(ptr, bytes));
    PRL_CUDA(cudaMemset(*ptr, 0, bytes));
    return PRL_OK;
}
int prl_p2p_free(void *ptr) {
    PRL_CUDA(cudaFree(ptr));
    return PRL_OK;
}
int prl_p2p_get_handle(void *ptr, unsigned char *handle64) {
    PRL_REQUIRE(ptr && handle64, "prl_p2p_get_handle: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PRL_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), ptr));
    return PRL_OK;
}
int prl_p2p_open_handle(const unsigned char *handle64, void **ptr) {
    PRL_REQUIRE(ptr && handle64, "prl_p2p_open_handle: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    PRL_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PRL_OK;
}
int prl_p2p_close_handle(void *ptr) {
    PRL_CUDA(cudaIpcCloseMemHandle(ptr));
    return PRL_OK;
}

/* ws[0]: 0 = every tensor-core phase of every call since the workspace was zeroed completed; 1 = an mbarrier wait timed
 * out (results invalid).  Host-synchronising. */
int prl_ppo_grad_tc_status(const float *ws, int *status_host, void *stream) {
    PRL_REQUIRE(ws && status_host, "prl_ppo_grad_tc_status: bad arguments");
    PRL_CUDA(cudaMemcpyAsync(status_host, ws, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PRL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    // a timed-out grid barrier leaves its arrival count behind: clear it so that the workspace can be used again
    if (*status_host == 2 || *status_host == 3)
        PRL_CUDA(cudaMemsetAsync(const_cast<float *>(ws) + 1, 0, 2 * sizeof(unsigned int), (cudaStream_t)stream));
    return PRL_OK;
}

}  // extern "C"
