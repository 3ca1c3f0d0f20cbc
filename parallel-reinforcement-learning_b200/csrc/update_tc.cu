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
// it staged operands through mbarriers (arrive, no wait) and only ever wait for MMA completion (mbarriers):
//     forward   Z[r][(h,j)]  = sum_k F[r][k] W1c_h[j][k]         M=128 N=64  K=64    (per head)
//     dgrad     DF[r][k]    += sum_j DZ_h[r][j] W1c_h[j][k]      M=128 N=64  K=64    (B = MN-major view of the same W1 bytes)
//     wgrad     DW_h[j][k]  += sum_r DZ_h[r][j] F[r][k]          M=128 N=64  K=128   (A, B = MN-major views of the same DZ / F bytes)
//     wgrad0    DW0[j][i]   += sum_r DZ0[r][j] X[r][i]           M=128 N=16  K=128
// The weight-gradient accumulators stay in tensor memory across ALL tiles of the CTA and are read out once.
//
// What remains on the CUDA cores are the row-wise nonlinearities (GroupNorm, SiLU, softmax / loss and their backward)
// and the narrow column sums (GroupNorm affine and output-layer gradients).  That part is written for issue slots, the
// resource that bounds the kernel (ncu, round 1: 41 M warp instructions per 65 536-row launch, issue 33 % busy, top stall
// "no instruction" on 229 KB of SASS):
//   * every element-wise step works on float2 feature pairs (FFMA2 / FADD2 / FMUL2, sm_100: half the instructions to fetch
//     and decode; measured on B200 a packed instruction occupies the FMA pipe for two issue cycles, so this buys code size,
//     not issue slots);
//   * GroupNorm without a mean pass: the hidden-layer weights are staged CENTRED per GroupNorm group
//     (W1c[j][:] = W1[j][:] - mean over the 8 rows j' of j's group), so the GEMM delivers z - mean_group(z) directly.
//     GroupNorm is invariant to that shift, its input gradient has zero group mean, so dgrad may use the same centred
//     bytes and the weight gradient is unchanged;
//   * column sums as warp butterfly reductions into per-lane REGISTER accumulators that live across all tiles of the CTA
//     and are combined once per launch in a fixed order (bit-reproducible).  (Measured and rejected: a per-warp
//     shared-memory transposition instead of the shuffles - 2 KB in and 2 KB out per quantity and warp is shared-memory
//     bandwidth the tensor core needs for its operands: 1.08 ms against 0.86 ms in tools/micro/ubench.cu);
//   * nothing is recomputed: what the backward needs of a forward stays in registers or is parked in this thread's
//     lane of tensor memory (tcgen05.st / tcgen05.ld: 1 instruction per 16 values);
//   * the four threads of a row exchange their partial head outputs through shared memory behind a 128-thread named
//     barrier per row quarter (the quarters run independently between MMA hand-offs).
#include <stdlib.h>
#include <string.h>

#include "policy.cuh"
#include "tiled_mlp.cuh"
#include "umma.cuh"

namespace prl {
using namespace umma;

// 16 compute warps + 1 MMA-issue warp; rows per tile; features per thread
constexpr int TC_COMPUTE = 512, TC_THREADS = TC_COMPUTE + 32, TC_ROWS = 128, TC_W = 16;
constexpr int PIECE = 8 * CHUNK;     // one bf16 piece of a [128][64] matrix: 8 chunks x 2048 B = 16 KB
constexpr int XPIECE = 2 * CHUNK;    // one bf16 piece of the [128][16] input matrix
constexpr int TC_MAX_O = 16, TC_MAX_A = 8;

// tensor-memory columns.  DF (trunk-activation gradient) reuses the actor's Z columns: the actor epilogue has read Z_0
// before it hands DZ_0 over, and only then is the first dgrad MMA issued.  ZH0 / DS0 / HJ are per-thread parking space
// (lane = row, 16 columns per feature quarter): the trunk's zhat and SiLU derivative for the trunk backward, a head's
// SiLU output for its dW2.
constexpr uint32_t TM_Z = 0, TM_DF = 0, TM_DW = 128 /* + 64 h */, TM_DW0 = 256, TM_ZH0 = 288, TM_DS0 = 352, TM_HJ = 416, TM_COLS = 512;

__host__ __device__ inline int tc_small_floats(const PolicyLayout &L) {
    int n = L.O * HID + 2 * HID;
    for (int h = 0; h < 2; ++h) n += 2 * HID + L.head[h].out * HID + round4(L.head[h].out);
    return n;
}
__host__ __device__ inline int tc_num_q(const PolicyLayout &L) { return 2 + 2 + L.head[0].out + 2 + L.head[1].out; }
// W (3 pieces) | F (3) | DZ (3 + a zero piece) | X (3 small) | small parameters | partial head outputs [4 q][128][NA] + [4 q][128]
__host__ __device__ inline size_t tc_smem_bytes(const PolicyLayout &L, int NA) {
    return 3 * PIECE + 3 * PIECE + 4 * PIECE + 3 * XPIECE + (size_t)((tc_small_floats(L) + 3) & ~3) * 4 + (size_t)4 * TC_ROWS * (NA + 1) * 4 + 128;
}

// ---- small math ----------------------------------------------------------------------------------------------------
typedef float2 f2;
__device__ __forceinline__ f2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 dup2(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 1 / (1 + exp(-y)) on a feature pair: FMUL2, 2 x MUFU.EX2, FADD2, 2 x MUFU.RCP (2 ulp; saturates cleanly at both ends)
__device__ __forceinline__ f2 sigmoid2(f2 y) {
    const f2 t = mul2(y, dup2(-1.4426950408889634f));
    const f2 d = add2(mk2(ex2_approx(t.x), ex2_approx(t.y)), dup2(1.0f));
    return mk2(rcp_approx(d.x), rcp_approx(d.y));
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float fast_exp(float x) { return ex2_approx(x * 1.4426950408889634f); }
__device__ __forceinline__ float fast_log(float x) { return lg2_approx(x) * 0.6931471805599453f; }

// write 16 fp32 values (features 16q..16q+15 of row r, as 8 pairs) as three bf16 pieces: chunks 2q, 2q+1 of a [128][64] piece triple
__device__ __forceinline__ void store_pieces16(unsigned char *base, int r, int q, const f2 (&v)[8]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t q0[4], q1[4], q2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const f2 x = v[4 * c + i];
            q0[i] = pack_bf16x2(x.x, x.y);
            const f2 r0 = fma2(mk2(__uint_as_float(q0[i] << 16), __uint_as_float(q0[i] & 0xffff0000u)), dup2(-1.0f), x);   // exact
            q1[i] = pack_bf16x2(r0.x, r0.y);
            const f2 r1 = fma2(mk2(__uint_as_float(q1[i] << 16), __uint_as_float(q1[i] & 0xffff0000u)), dup2(-1.0f), r0);  // exact
            q2[i] = pack_bf16x2(r1.x, r1.y);
        }
        unsigned char *p = base + (2 * q + c) * CHUNK + r * 16;
        *reinterpret_cast<uint4 *>(p) = make_uint4(q0[0], q0[1], q0[2], q0[3]);
        *reinterpret_cast<uint4 *>(p + PIECE) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
        *reinterpret_cast<uint4 *>(p + 2 * PIECE) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
    }
}

// column sum of 8 per-row values over the warp's 32 rows, first three butterfly levels: on return lane l holds the sum over the 8
// lanes that share l's bits 0 and 1 of feature f(lane) = 4*bit4 + 2*bit3 + bit2.  The sum over the remaining four lanes (xor 2,
// xor 1) is linear, so it is taken ONCE per launch on the accumulated values (colsum_finish) instead of once per tile.  7 shuffles.
// (A/B, round 2: the same function NOT inlined - 9 KB less code - costs 4.9 %: call overhead beats instruction-cache relief)
__device__ __forceinline__ float colsum8(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7) {
    const int lane = threadIdx.x & 31;
    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
    const float a0 = (u16 ? v4 : v0) + __shfl_xor_sync(0xffffffffu, u16 ? v0 : v4, 16);
    const float a1 = (u16 ? v5 : v1) + __shfl_xor_sync(0xffffffffu, u16 ? v1 : v5, 16);
    const float a2 = (u16 ? v6 : v2) + __shfl_xor_sync(0xffffffffu, u16 ? v2 : v6, 16);
    const float a3 = (u16 ? v7 : v3) + __shfl_xor_sync(0xffffffffu, u16 ? v3 : v7, 16);
    const float b0 = (u8 ? a2 : a0) + __shfl_xor_sync(0xffffffffu, u8 ? a0 : a2, 8);
    const float b1 = (u8 ? a3 : a1) + __shfl_xor_sync(0xffffffffu, u8 ? a1 : a3, 8);
    return (u4 ? b1 : b0) + __shfl_xor_sync(0xffffffffu, u4 ? b0 : b1, 4);
}
// the deferred levels of colsum8, applied to a register accumulator at the end of the launch: afterwards all four lanes holding a
// feature hold its total
__device__ __forceinline__ void colsum_finish(f2 &acc) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
}
// column sums of the thread's 16 features (8 pairs) over the warp's 32 rows, added to the lane's register accumulators:
// acc.x += partial total of feature f(lane) of group 0, acc.y += the same of group 1 (no divergent update, nothing in memory; the
// accumulators live across all tiles of the CTA and are completed by colsum_finish)
__device__ __forceinline__ void colsum16(const f2 (&v)[8], f2 &acc) {
    acc.x += colsum8(v[0].x, v[0].y, v[1].x, v[1].y, v[2].x, v[2].y, v[3].x, v[3].y);
    acc.y += colsum8(v[4].x, v[4].y, v[5].x, v[5].y, v[6].x, v[6].y, v[7].x, v[7].y);
}

// GroupNorm on the thread's two groups when the GEMM delivered z - mean_group(z) (centred weights): z -> zhat in place
__device__ __forceinline__ void gn_normalize_centred(f2 (&z)[8], float (&rstd)[2]) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        f2 v = mul2(z[4 * g], z[4 * g]);
#pragma unroll
        for (int i = 1; i < 4; ++i) v = fma2(z[4 * g + i], z[4 * g + i], v);
        const float r = rsqrtf(fmaf(v.x + v.y, 1.0f / GSIZE, GN_EPS));
        rstd[g] = r;
#pragma unroll
        for (int i = 0; i < 4; ++i) z[4 * g + i] = mul2(z[4 * g + i], dup2(r));
    }
}
// dy -> dz in place: d = dy * gamma, dz = rstd * (d - mean(d) - zhat * mean(d * zhat)) per group
__device__ __forceinline__ void gn_backward16(f2 (&d)[8], const f2 (&zhat)[8], const float (&rstd)[2], const float *gamma16) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const float4 ga = *reinterpret_cast<const float4 *>(gamma16 + 8 * g), gb = *reinterpret_cast<const float4 *>(gamma16 + 8 * g + 4);
        d[4 * g] = mul2(d[4 * g], mk2(ga.x, ga.y));
        d[4 * g + 1] = mul2(d[4 * g + 1], mk2(ga.z, ga.w));
        d[4 * g + 2] = mul2(d[4 * g + 2], mk2(gb.x, gb.y));
        d[4 * g + 3] = mul2(d[4 * g + 3], mk2(gb.z, gb.w));
        f2 s1 = add2(d[4 * g], d[4 * g + 1]), s2 = mul2(d[4 * g], zhat[4 * g]);
        s1 = add2(s1, add2(d[4 * g + 2], d[4 * g + 3]));
#pragma unroll
        for (int i = 1; i < 4; ++i) s2 = fma2(d[4 * g + i], zhat[4 * g + i], s2);
        const float a = rstd[g];
        const float c1 = -(s1.x + s1.y) * (1.0f / GSIZE) * a, c2 = -(s2.x + s2.y) * (1.0f / GSIZE) * a;
#pragma unroll
        for (int i = 0; i < 4; ++i) d[4 * g + i] = fma2(d[4 * g + i], dup2(a), fma2(zhat[4 * g + i], dup2(c2), dup2(c1)));
    }
}
__device__ __forceinline__ void load16(const float *src, f2 (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = reinterpret_cast<const float4 *>(src)[i];
        v[2 * i] = mk2(t.x, t.y); v[2 * i + 1] = mk2(t.z, t.w);
    }
}

// named barrier 1 = the 512 compute threads among themselves (the MMA warp never joins it); 2 + rq = the four warps of row quarter rq
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void bar_quarter(int rq) { asm volatile("bar.sync %0, 128;" ::"r"(2 + rq) : "memory"); }
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

// ---- tensor memory <-> registers, 16 columns of this thread's lane as 8 float pairs ----------------------------------------
__device__ __forceinline__ void tmem_ld16_f2(uint32_t taddr, f2 (&v)[8]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = mk2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
}
// two loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x2_f2(uint32_t ta, f2 (&a)[8], uint32_t tb, f2 (&b)[8]) {
    uint32_t r[16], s[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]), "=r"(s[9]),
          "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15])
        : "r"(ta), "r"(tb)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = mk2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        b[i] = mk2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
    }
}
// park 16 values in this thread's lane (completes before the thread goes on: the same thread reads them back later)
__device__ __forceinline__ void tmem_st16_f2(uint32_t taddr, const f2 (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n\t"
        "tcgen05.wait::st.sync.aligned;"
        :
        : "r"(taddr), "r"(__float_as_uint(v[0].x)), "r"(__float_as_uint(v[0].y)), "r"(__float_as_uint(v[1].x)), "r"(__float_as_uint(v[1].y)),
          "r"(__float_as_uint(v[2].x)), "r"(__float_as_uint(v[2].y)), "r"(__float_as_uint(v[3].x)), "r"(__float_as_uint(v[3].y)),
          "r"(__float_as_uint(v[4].x)), "r"(__float_as_uint(v[4].y)), "r"(__float_as_uint(v[5].x)), "r"(__float_as_uint(v[5].y)),
          "r"(__float_as_uint(v[6].x)), "r"(__float_as_uint(v[6].y)), "r"(__float_as_uint(v[7].x)), "r"(__float_as_uint(v[7].y))
        : "memory");
}

// Fused optimiser tail: every CTA writes its partial-gradient row and loss sums, then adds one to an arrival counter
// (release); every CTA waits for the counter to reach the grid size (ONE word polled by one lane per CTA - polling a word
// per producer from every CTA hot-spots a handful of L2 lines: measured 3 us per exchange), reduces its 64-parameter
// slices over the CTAs' partial rows in a fixed order and publishes the squared norm of its slices as a tagged word.
// The LEADER CTA (the last one: it owns no slice when the grid is larger than the slice count) collects those, and
// publishes the clip coefficient as one tagged word that everybody else polls; then every CTA applies clip_grad_norm_ +
// AdamW to its slices.  One launch per optimiser step instead of three; no grid barrier, no host-side reset.
// Continuous policies (mu head, log_std head, critic: ActorCritic.py:28-42) run through this two-head kernel as TWO passes over
// the minibatch.  Their loss couples the mu and log_std outputs, so a forward pre-pass (k_policy_dout) evaluates the loss and
// leaves d loss / d mu and d loss / d log_std per row; given those, the heads are independent: pass 1 = {mu head with its
// external output gradient, critic with its own SmoothL1}, pass 2 = {log_std head with its external gradient, critic weighted 0}.
// Everything downstream of the output gradient is linear in it, so the trunk's gradient is the sum of the two passes'.
struct TcExternal {
    const float *dout;      // [b][out of head 0] gradient of the loss w.r.t. head 0's outputs (inv_count included); nullptr: discrete loss
    float critic_weight;    // 1: the critic's SmoothL1 term takes part; 0: the critic head is skipped altogether (second pass)
    int zero_off, zero_len; // range of the partial-gradient row that belongs to the head this pass does not touch: written as zeros
};

struct TcOptimizer {
    float *params_rw, *grad, *m, *v;          // params_rw == nullptr: gradient only (the reduction runs as a separate kernel)
    int64_t *clock;                           // {int64 step, double beta1^step, double beta2^step}
    unsigned int *count;                      // arrival counter behind the partial rows (zero between launches: the leader CTA resets it)
    unsigned long long *tagw;                 // [grid][2] tagged words: squared norm of the CTA's slices (a double in two halves)
    unsigned long long *coefw;                // 2 tagged words published by the leader CTA: clip coefficient, total gradient norm
    double *norm_out;
    float lr, wd, max_norm;
    double *loss_out;                         // 4 doubles, accumulated
    double rows;
    unsigned long long timeout_ns;            // bound of every cross-CTA / cross-GPU wait (wall clock, %globaltimer)
    // sharded runs: the gradient exchange happens inside this kernel over NVLink peer memory instead of an NCCL allreduce.
    // peers[r] = base of rank r's exchange buffer: inbox[2 (step parity)][world (sender)][gstride] words {float bits, step number}
    float *const *peers;
    int rank, world, gstride;
};

__device__ __noinline__ double ipow(double base, int64_t n) {
    double r = 1.0;
    for (; n > 0; n >>= 1, base *= base)
        if (n & 1) r *= base;
    return r;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
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
__device__ __forceinline__ void red_release_gpu_add(unsigned int *p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// tagged words inside one GPU: {32 payload bits, launch number}; a naturally aligned 64-bit access is single-copy atomic, so a
// reader that sees this launch's number also sees the payload - no fence, no barrier between producer and consumer
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
constexpr int TC_MAX_GRID = 160;   // lanes x 5: a warp holds one word per CTA in flight
// one thread: wait until *count == nb, then acquire (the partial rows behind the counter).  false on time-out.
__device__ __forceinline__ bool wait_count(const unsigned int *count, unsigned int nb, unsigned long long timeout_ns) {
    const unsigned long long t0 = globaltimer_ns();
    bool ok = true;
    for (int it = 0; ld_relaxed_gpu(count) != nb; ++it)
        if ((it & 63) == 63 && globaltimer_ns() - t0 > timeout_ns) { ok = false; break; }
    __threadfence();
    return ok;
}
// one thread: poll a tagged word until it carries this launch's number; returns the payload
__device__ __forceinline__ unsigned int wait_tag(const unsigned long long *p, unsigned int epoch, unsigned long long timeout_ns, bool &ok) {
    const unsigned long long t0 = globaltimer_ns();
    unsigned long long v = ld_tag(p);
    for (int it = 0; (unsigned int)(v >> 32) != epoch; ++it) {
        if ((it & 63) == 63 && globaltimer_ns() - t0 > timeout_ns) { ok = false; break; }
        v = ld_tag(p);
    }
    return (unsigned int)v;
}
// one warp: sum over the CTAs of a tagged double (lane l adds CTAs l, l + 32, ... in ascending order, then a fixed shuffle
// tree: bit-reproducible); every word is polled until it carries this launch's number, all of a lane's words in flight
__device__ __forceinline__ double sum_tagged_doubles(const unsigned long long *tagw, int nb, unsigned int epoch, int lane, unsigned long long timeout_ns,
                                                     bool &ok) {
    const unsigned long long t0 = globaltimer_ns();
    unsigned long long hi[TC_MAX_GRID / 32], lo[TC_MAX_GRID / 32];
    bool all = false;
    for (int it = 0; !all; ++it) {
        bool mine = true;
#pragma unroll
        for (int k = 0; k < TC_MAX_GRID / 32; ++k) {
            const int bl = lane + 32 * k;
            hi[k] = lo[k] = (unsigned long long)epoch << 32;
            if (bl < nb) { hi[k] = ld_tag(tagw + 2 * bl); lo[k] = ld_tag(tagw + 2 * bl + 1); }
            mine = mine && (unsigned int)(hi[k] >> 32) == epoch && (unsigned int)(lo[k] >> 32) == epoch;
        }
        all = __all_sync(0xffffffffu, mine);
        if (!all && (it & 63) == 63 && globaltimer_ns() - t0 > timeout_ns) break;
    }
    ok = ok && all;
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < TC_MAX_GRID / 32; ++k)
        if (lane + 32 * k < nb) a += __longlong_as_double((long long)(((hi[k] & 0xffffffffull) << 32) | (lo[k] & 0xffffffffull)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    return a;
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
// per-CTA wall-clock marks (ns): [cta][0] first instruction, [1] tiles done, [2] partials written, [3] last instruction,
// [4] READY flag raised, [5] all flags seen, [6] slices reduced, [7] norm known
__device__ unsigned long long g_tc_span[160 * 8];
#define TC_SPAN(i) do { if (threadIdx.x == 0 && blockIdx.x < 160) g_tc_span[blockIdx.x * 8 + (i)] = globaltimer_ns(); } while (0)

// ===================================================================================================== the kernel
// NA = compile-time bound of the actor's output width (action_dim rounded up to 2, 4 or 8): the per-output loops unroll over
// it.  XR = observation values kept in registers (4 or 8; observ_dim > XR reads the rest on demand).
// SHARDED: the build with the cross-GPU gradient exchange in its tail (prl_ppo_step_tc_p2p); the single-GPU build leaves that code out
// (the instruction stream is larger than the instruction cache as it is).
// SKIP1: the second pass of a continuous policy's update (below), where the critic head is left out altogether; a template
// parameter because even a kernel-uniform runtime flag costs the main build 1.6 % (A/B: registers).
template <int NA, int XR, bool SHARDED, bool SKIP1>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_ppo_grad_tc(const float *__restrict__ params, PolicyLayout L, const float *__restrict__ states, const float *__restrict__ actions,
              const float *__restrict__ old_logp, const float *__restrict__ adv, const float *__restrict__ returns, int64_t b,
              float clip, float inv_count, float *__restrict__ partials, int part_stride, double *__restrict__ loss_partials,
              int *__restrict__ status, TcOptimizer opt, int qpc, TcExternal ext) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[6];          // MMA completion: actor forward, actor backward, critic dgrad, trunk wgrad, critic wgrad, critic forward
    __shared__ uint64_t sbar[4];          // operands staged (512 arrivals): F + X, actor DZ, critic DZ, trunk DZ
    __shared__ uint32_t tmem_slot;
    __shared__ double red[3][4];          // loss sums per row quarter
    __shared__ float b2s[4][TC_MAX_A + 1];
    __shared__ int tail_ok_s;
    __shared__ float coef_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_mma_warp = warp == TC_COMPUTE / 32;
    const int rq = warp & 3, q = (warp >> 2) & 3;   // row quarter (= tensor-memory lane quarter), feature quarter
    const int r = rq * 32 + lane, j0 = TC_W * q;    // row of the tile, first feature of this thread
    const int O = L.O, A = L.A, nout0 = L.head[0].out;
    TC_STAMP(0);
    TC_SPAN(0);
    // second pass of a continuous policy's update: the critic contributes nothing (its gradient came from the first pass), so the
    // whole head - forward, epilogue, dgrad, wgrad - is left out and its partial-gradient entries are written as zeros
    constexpr bool skip1 = SKIP1;
    if (is_mma_warp) tmem_alloc(&tmem_slot, TM_COLS);   // overlaps the parameter staging of the other warps
    else if (q == 0 && lane == 0) {
        // the first tile's per-row scalars start their way from HBM to L2 while the parameters are staged
        const int64_t row0 = (int64_t)blockIdx.x * qpc * 32 + r;
        if (row0 < b) {
            if (ext.dout == nullptr) { prefetch_l2(adv + row0); prefetch_l2(old_logp + row0); prefetch_l2(actions + row0); }
            else prefetch_l2(ext.dout + row0 * nout0);
            prefetch_l2(returns + row0);
        }
    }

    // ---- everything the launch needs from global memory besides the rows is requested up front (one exposed latency)
    const bool fused = opt.params_rw != nullptr;
    int64_t opt_step = 0;
    double opt_p1 = 0.0, opt_p2 = 0.0;
    unsigned int epoch = 0;
    if (fused) {
        // optimiser clock, read before anybody can advance it (CTA 0 does, at the very end); the launch number tags the
        // READY flags and the squared-norm words of this launch (kept in the workspace header, advanced by CTA 0 at the end)
        opt_step = opt.clock[0];
        opt_p1 = reinterpret_cast<const double *>(opt.clock)[1];
        opt_p2 = reinterpret_cast<const double *>(opt.clock)[2];
        epoch = __ldcg(reinterpret_cast<const unsigned int *>(status) + 3) + 1u;
    }

    // ---- carve shared memory (all pointers derive from smem_raw so they stay in the shared state space)
    unsigned char *sW = smem_raw, *sF = sW + 3 * PIECE, *sDZ = sF + 3 * PIECE, *sX = sDZ + 4 * PIECE;
    float *sSmall = reinterpret_cast<float *>(sX + 3 * XPIECE);
    float *s_w0t = sSmall, *s_g0w = s_w0t + O * HID, *s_g0b = s_g0w + HID;
    float *s_head0 = s_g0b + HID;                                   // per head: gw[64] gb[64] w2[out][64] b2[round4(out)]
    const int head0_floats = 2 * HID + nout0 * HID + round4(nout0);
    float *s_head1 = s_head0 + head0_floats;
    float *s_po0 = sSmall + ((tc_small_floats(L) + 3) & ~3);        // [4 q][128 r][NA] partial actor outputs
    float *s_po1 = s_po0 + 4 * TC_ROWS * NA;                        // [4 q][128 r]     partial critic outputs

    // ---- stage parameters: fp32 small ones; W1 of both heads as bf16x3 in the operand layout (row n = h*64 + j).
    // The two hidden-layer matrices and W0 are staged CENTRED over each GroupNorm group of output features (see the header).
    {
        f2 v[8];
        const int n_small0 = HID * O, n_small1 = n_small0 + 2 * HID, n_h0 = 2 * HID + nout0 * HID + nout0,
                  n_h1 = 2 * HID + L.head[1].out * HID + L.head[1].out, n_small = n_small1 + n_h0 + n_h1;
        if (!is_mma_warp) {
            // thread (r, q) stages features 16q..16q+15 of row n = r of the stacked [128][64] W1 (n < 64: actor, else critic)
            // (A/B, round 2: rotating the row every CTA starts with, so that the 147 CTAs do not ask L2 for the same lines at once, changes
            // nothing - 0.1 %: the set-up is not bound by an L2 hot spot)
            const int w1off = (r >> 6) ? L.head[1].w1 : L.head[0].w1;   // (no dynamic index into the kernel-parameter struct)
            const float *wrow = params + w1off + (r & 63) * HID + j0;
            TC_STAMP(19);
            if ((w1off & 3) == 0) {   // warp-uniform (a warp's rows belong to one head)
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float4 t = __ldg(reinterpret_cast<const float4 *>(wrow) + k); v[2 * k] = mk2(t.x, t.y); v[2 * k + 1] = mk2(t.z, t.w); }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = mk2(__ldg(wrow + 2 * k), __ldg(wrow + 2 * k + 1));
            }
        }
        // (timing aid: CTA 0's first warp waits for its W1 row here.  A/B on one box, tools/ab_build.sh: the build WITHOUT this line is
        // 0.5 % slower - 60.9 against 60.6 us per launch - so it stays)
        if (blockIdx.x == 0 && tid == 0 && v[0].x == 123456.789f) g_tc_clock[31] = 1;
        TC_STAMP(20);
        // small parameters: w0 (stored transposed and centred), then three blocks that are contiguous both in `params` and in
        // shared memory: {g0w, g0b}, head 0 {gw, gb, w2, b2}, head 1 {gw, gb, w2, b2}
        for (int base = 0; base < n_small; base += 2 * TC_THREADS) {
            float sv[2];
            float *dst[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = base + u * TC_THREADS + tid;
                dst[u] = nullptr;
                if (i < n_small0) {
                    const int j = i / O, o = i - j * O;
                    const float *gp = params + L.w0 + (j & ~(GSIZE - 1)) * O + o;   // the 8 rows of j's group, column o
                    float m = 0.f;
#pragma unroll
                    for (int t = 0; t < GSIZE; ++t) m += __ldg(gp + t * O);
                    dst[u] = s_w0t + o * HID + j;
                    sv[u] = __ldg(params + L.w0 + i) - m * (1.0f / GSIZE);
                }
                else if (i < n_small1) { dst[u] = s_g0w + (i - n_small0); sv[u] = __ldg(params + L.g0w + (i - n_small0)); }
                else if (i < n_small1 + n_h0) { dst[u] = s_head0 + (i - n_small1); sv[u] = __ldg(params + L.head[0].gw + (i - n_small1)); }
                else if (i < n_small) { dst[u] = s_head1 + (i - n_small1 - n_h0); sv[u] = __ldg(params + L.head[1].gw + (i - n_small1 - n_h0)); }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (dst[u]) *dst[u] = sv[u];
        }
        TC_STAMP(17);
        if (!is_mma_warp) {
            // centre W1 over the 8 rows of the GroupNorm group: lanes 8i..8i+7 of a warp hold rows 8i..8i+7 of one head
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                f2 s = v[k];
#pragma unroll
                for (int o = 1; o < GSIZE; o <<= 1) { s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o); }
                v[k] = fma2(s, dup2(-1.0f / GSIZE), v[k]);
            }
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
        tail_ok_s = 1;
    }
    TC_STAMP(18);
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
    const int P = L.total;
    float *part = partials + (size_t)blockIdx.x * part_stride;

    if (is_mma_warp) {
        // =============================================================================== MMA-issue warp
        for (int tile = 0; tile < ntiles; ++tile, ++it) {
            const uint32_t parity = it & 1;
            mma_ok &= mbar_wait(&sbar[0], parity);
            fence_after_sync();
            if (elect_one()) { issue_forward(D, 0); mma_commit(&bars[0]); if (!skip1) issue_forward(D, 1); mma_commit(&bars[5]); }
            __syncwarp();
            mma_ok &= mbar_wait(&sbar[1], parity);
            fence_after_sync();
            if (elect_one()) { issue_head_dgrad(D, 0); issue_head_wgrad(D, 0, it == 0); mma_commit(&bars[1]); }
            __syncwarp();
            mma_ok &= mbar_wait(&sbar[2], parity);
            fence_after_sync();
            // the trunk backward needs DF (both dgrads) only: the critic's wgrad gets its own completion barrier
            // (skip1: the commits still mark the completion of the first head's MMAs, which is what the trunk backward waits for)
            if (elect_one()) {
                if (!skip1) issue_head_dgrad(D, 1);
                mma_commit(&bars[2]);
                if (!skip1) issue_head_wgrad(D, 1, it == 0);
                mma_commit(&bars[4]);
            }
            __syncwarp();
            mma_ok &= mbar_wait(&sbar[3], parity);
            fence_after_sync();
            if (elect_one()) { issue_trunk_wgrad(D, it == 0); mma_commit(&bars[3]); }
            __syncwarp();
        }
    } else {
        // =============================================================================== compute warps
        // column-sum accumulators: registers, live across all tiles (.x: feature f(lane) of this thread's group 0, .y: of group 1)
        f2 acc_g0 = dup2(0.f), acc_b0 = dup2(0.f), acc_gh[2] = {dup2(0.f), dup2(0.f)}, acc_bh[2] = {dup2(0.f), dup2(0.f)}, acc_w2c = dup2(0.f);
        f2 acc_w2a[NA];
        float b2a[NA], b2c = 0.f;
#pragma unroll
        for (int a = 0; a < NA; ++a) { acc_w2a[a] = dup2(0.f); b2a[a] = 0.f; }
        float l_pol = 0.f, l_val = 0.f, l_ent = 0.f;
        const bool softmax_rows_sum_to_zero = ext.dout == nullptr;   // head 0 is a softmax head evaluated here (not an external gradient)
        const bool own_loss_pf = ext.dout == nullptr;

        // inputs of the first tile; inside the loop the next tile's are prefetched while the current one computes
        // (A/B, round 2: requesting them before the parameter staging instead gains nothing, L2-resident or streaming)
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
                staged(&sbar[1]);
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
                // the next tile's per-row scalars (one 128-byte line per array and row quarter) are pulled into L2 now: their loads
                // at the top of the next tile would otherwise wait for HBM inside the actor epilogue (ncu: long-scoreboard samples
                // on the first use of `actions`)
                if (q == 0 && lane == 0 && nlive) {
                    if (own_loss_pf) { prefetch_l2(adv + rown); prefetch_l2(old_logp + rown); prefetch_l2(actions + rown); }
                    else prefetch_l2(ext.dout + rown * nout0);
                    prefetch_l2(returns + rown);
                }
            }
            const bool own_loss = ext.dout == nullptr;   // (warp-uniform) false: head 0's output gradient comes from the pre-pass
            const float adv_i = (live && own_loss) ? __ldg(adv + row) : 0.f, old_i = (live && own_loss) ? __ldg(old_logp + row) : 0.f;
            const float ret_i = live ? __ldg(returns + row) : 0.f;
            const int act = (live && own_loss) ? (int)__ldg(actions + row) : 0;

            // ================= trunk forward (CUDA cores): z0 = W0c x, GroupNorm, SiLU -> F pieces, X pieces; zhat0 and SiLU'
            // are parked in tensor memory for the trunk backward
            float rs0[2];
            {
                f2 z[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) z[k] = dup2(0.f);
#pragma unroll
                for (int i = 0; i < XR; ++i) {
                    if (i < O) {
                        f2 w[8];
                        load16(s_w0t + i * HID + j0, w);
#pragma unroll
                        for (int k = 0; k < 8; ++k) z[k] = fma2(dup2(x[i]), w[k], z[k]);
                    }
                }
                for (int i = XR; i < O; ++i) {
                    const float xi = xmask * __ldg(xrow + i);
                    f2 w[8];
                    load16(s_w0t + i * HID + j0, w);
#pragma unroll
                    for (int k = 0; k < 8; ++k) z[k] = fma2(dup2(xi), w[k], z[k]);
                }
                gn_normalize_centred(z, rs0);
                f2 f[8], d0[8];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 g = *reinterpret_cast<const float4 *>(s_g0w + j0 + 4 * k4), bt = *reinterpret_cast<const float4 *>(s_g0b + j0 + 4 * k4);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int k = 2 * k4 + hf;
                        const f2 y = fma2(z[k], hf ? mk2(g.z, g.w) : mk2(g.x, g.y), hf ? mk2(bt.z, bt.w) : mk2(bt.x, bt.y));
                        const f2 s = sigmoid2(y);
                        f[k] = mul2(y, s);
                        d0[k] = fma2(f[k], fma2(s, dup2(-1.0f), dup2(1.0f)), s);   // SiLU'(y) = s + y s (1 - s)
                    }
                }
                tmem_st16_f2(lane_base + TM_ZH0 + j0, z);
                tmem_st16_f2(lane_base + TM_DS0 + j0, d0);
                // the previous tile's trunk-wgrad MMAs read X and DZ; its head MMAs (already waited for) read F
                if (it > 0) mma_ok &= mbar_wait(&bars[3], parity ^ 1);
                store_pieces16(sF, r, q, f);
                if (q == 0) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (8 * c < O) {   // (columns >= O stay zero from the set-up)
                            uint32_t q0[4], q1[4], q2[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int e = 8 * c + 2 * i;
                                float xa, xb;
                                if (e + 1 < XR) { xa = x[e < XR ? e : 0]; xb = x[e + 1 < XR ? e + 1 : 0]; }
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
            }
            staged(&sbar[0]);
            if (it == 0) TC_STAMP(2);
            mma_ok &= mbar_wait(&bars[0], parity);
            fence_after_sync();
            if (it == 0) TC_STAMP(4);

            // ================= heads: forward epilogue, loss, backward epilogue -> DZ pieces, tensor-core dgrad + wgrad
            auto head = [&](auto HC) {
                constexpr int H = decltype(HC)::value;
                constexpr int NO = H == 0 ? NA : 1;             // compile-time bound of this head's output width
                const int nout = H == 0 ? nout0 : 1;
                const float *sh = H ? s_head1 : s_head0;
                const float *w2 = sh + 2 * HID;
                float *spo = H ? s_po1 : s_po0;
                if (H == 1) { mma_ok &= mbar_wait(&bars[5], parity); fence_after_sync(); }   // critic forward complete
                f2 zh[8], ds[8];
                float rstd[2];
                tmem_ld16_f2(lane_base + TM_Z + 64 * H + j0, zh);
                gn_normalize_centred(zh, rstd);
                {
                    f2 hj[8], po2[NO];
#pragma unroll
                    for (int a = 0; a < NO; ++a) po2[a] = dup2(0.f);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const float4 g = *reinterpret_cast<const float4 *>(sh + j0 + 4 * k4), bt = *reinterpret_cast<const float4 *>(sh + HID + j0 + 4 * k4);
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int k = 2 * k4 + hf;
                            const f2 y = fma2(zh[k], hf ? mk2(g.z, g.w) : mk2(g.x, g.y), hf ? mk2(bt.z, bt.w) : mk2(bt.x, bt.y));
                            const f2 s = sigmoid2(y);
                            hj[k] = mul2(y, s);
                            ds[k] = fma2(hj[k], fma2(s, dup2(-1.0f), dup2(1.0f)), s);
                        }
#pragma unroll
                        for (int a = 0; a < NO; ++a) {
                            if (a < nout) {
                                const float4 w = *reinterpret_cast<const float4 *>(w2 + a * HID + j0 + 4 * k4);
                                po2[a] = fma2(hj[2 * k4], mk2(w.x, w.y), po2[a]);
                                po2[a] = fma2(hj[2 * k4 + 1], mk2(w.z, w.w), po2[a]);
                            }
                        }
                    }
                    tmem_st16_f2(lane_base + TM_HJ + j0, hj);   // parked for dW2
#pragma unroll
                    for (int a = 0; a < NO; ++a)
                        if (a < nout) spo[(q * TC_ROWS + r) * NO + a] = po2[a].x + po2[a].y;
                }
                bar_quarter(rq);   // the four threads of a row have published their partial outputs
                float out[NO];
#pragma unroll
                for (int a = 0; a < NO; ++a)
                    out[a] = (a < nout) ? w2[nout * HID + a] + ((spo[(0 * TC_ROWS + r) * NO + a] + spo[(1 * TC_ROWS + r) * NO + a]) +
                                                              (spo[(2 * TC_ROWS + r) * NO + a] + spo[(3 * TC_ROWS + r) * NO + a]))
                                        : 0.f;
                // ---- loss and output gradients (the four threads of a row compute them redundantly; q == 0 keeps the sums)
                float dout[NO];
#pragma unroll
                for (int a = 0; a < NO; ++a) dout[a] = 0.f;
                if (H == 0 && !own_loss) {
                    if (live) {
#pragma unroll
                        for (int a = 0; a < NO; ++a)
                            if (a < nout) dout[a] = __ldg(ext.dout + row * nout + a);
                    }
                } else if (H == 0) {
                    if (live) {
                        float m = out[0];
#pragma unroll
                        for (int a = 1; a < NO; ++a)
                            if (a < A) m = fmaxf(m, out[a]);
                        float p[NO], Ssum = 0.f, Psum = 0.f;
#pragma unroll
                        for (int a = 0; a < NO; ++a) { p[a] = (a < A) ? fast_exp(out[a] - m) : 0.f; Ssum += p[a]; }
                        const float iS = rcp_approx(Ssum);
#pragma unroll
                        for (int a = 0; a < NO; ++a) { p[a] *= iS; Psum += p[a]; }
                        const float iP = rcp_approx(Psum);
                        float pa = 0.f, ent = 0.f;
#pragma unroll
                        for (int a = 0; a < NO; ++a) {
                            if (a < A) {
                                p[a] *= iP;
                                const float l = fast_log(fminf(fmaxf(p[a], F32_EPS), 1.0f - F32_EPS));
                                ent = fmaf(-l, p[a], ent);
                                if (a == act) pa = p[a];
                            }
                        }
                        const float logp = fast_log(fminf(fmaxf(pa, F32_EPS), 1.0f - F32_EPS));
                        const float dl = logp - old_i;
                        const float rr = fast_exp(fminf(fmaxf(dl, -20.f), 20.f));
                        const float s1 = rr * adv_i;
                        const float s2 = fminf(fmaxf(rr, 1.0f - clip), 1.0f + clip) * adv_i;
                        const float g1 = s1 < s2 ? 1.f : (s1 > s2 ? 0.f : 0.5f);   // torch.min splits ties evenly
                        const float in_clip = (rr >= 1.0f - clip && rr <= 1.0f + clip) ? 1.f : 0.f;
                        const float in20 = (dl >= -20.f && dl <= 20.f) ? 1.f : 0.f;
                        float dlogp = -inv_count * adv_i * (g1 + (1.f - g1) * in_clip) * rr * in20;
                        if (!(pa >= F32_EPS && pa <= 1.0f - F32_EPS)) dlogp = 0.f;   // clamp in probs_to_logits blocks the gradient
#pragma unroll
                        for (int a = 0; a < NO; ++a)
                            if (a < A) dout[a] = dlogp * ((a == act ? 1.f : 0.f) - p[a]);
                        if (q == 0) { l_pol += -fminf(s1, s2); l_ent += ent; }
                    }
                } else if (live) {
                    const float dv = out[0] - ret_i, ad = fabsf(dv);
                    if (q == 0) l_val += ext.critic_weight * (ad < 1.f ? 0.5f * dv * dv : ad - 0.5f);
                    dout[0] = ext.critic_weight * (0.5f * inv_count * (ad < 1.f ? dv : (dv > 0.f ? 1.f : -1.f)));
                }
                // ---- dy = (dout . W2) * SiLU'(y), in place of ds
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    f2 dh0 = dup2(0.f), dh1 = dup2(0.f);
#pragma unroll
                    for (int a = 0; a < NO; ++a) {
                        if (a < nout) {
                            const float4 w = *reinterpret_cast<const float4 *>(w2 + a * HID + j0 + 4 * k4);
                            dh0 = fma2(dup2(dout[a]), mk2(w.x, w.y), dh0);
                            dh1 = fma2(dup2(dout[a]), mk2(w.z, w.w), dh1);
                        }
                    }
                    ds[2 * k4] = mul2(dh0, ds[2 * k4]);
                    ds[2 * k4 + 1] = mul2(dh1, ds[2 * k4 + 1]);
                }
                // ---- GroupNorm backward -> dz, handed to the tensor core BEFORE the column sums below: the dgrad / wgrad MMAs
                // run under them (what the trunk backward waits for is the critic's dgrad)
                {
                    f2 dz[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) dz[k] = ds[k];
                    gn_backward16(dz, zh, rstd, sh + j0);
                    if (it == 0) TC_STAMP(5 + 3 * H);
                    // the actor's MMAs read DZ: they must have completed before the critic overwrites it
                    if (H == 1) mma_ok &= mbar_wait(&bars[1], parity);
                    store_pieces16(sDZ, r, q, dz);
                    staged(&sbar[1 + H]);
                    if (it == 0) TC_STAMP(6 + 3 * H);
                }
                // ---- output-layer gradients: dW2[a][j] = sum_r dout[a] h_j, db2[a] = sum_r dout[a]
                {
                    f2 hj[8];
                    tmem_ld16_f2(lane_base + TM_HJ + j0, hj);
#pragma unroll
                    for (int a = 0; a < NO; ++a) {
                        if (a < nout - ((H == 0 && softmax_rows_sum_to_zero) ? 1 : 0)) {   // (softmax head: the last action follows from the others)
                            f2 t[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) t[k] = mul2(hj[k], dup2(dout[a]));
                            if (H == 0) { colsum16(t, acc_w2a[a]); b2a[a] += dout[a]; }
                            else { colsum16(t, acc_w2c); b2c += dout[a]; }
                        }
                    }
                }
                // ---- GroupNorm-affine gradients
                {
                    f2 t[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) t[k] = mul2(ds[k], zh[k]);
                    colsum16(t, acc_gh[H]);
                    colsum16(ds, acc_bh[H]);
                }
            };
            head(std::integral_constant<int, 0>{});
            if (!skip1) head(std::integral_constant<int, 1>{});
            else staged(&sbar[2]);   // keep the hand-off (nothing was staged: the MMA warp issues nothing for it)

            // ================= trunk backward: DF -> dy0 -> GroupNorm backward -> DZ pieces, tensor-core wgrad against X
            {
                f2 df[8], zh0[8], d0[8];
                tmem_ld16x2_f2(lane_base + TM_ZH0 + j0, zh0, lane_base + TM_DS0 + j0, d0);
                mma_ok &= mbar_wait(&bars[2], parity);    // critic dgrad complete -> DF final
                fence_after_sync();
                if (it == 0) TC_STAMP(11);
                tmem_ld16_f2(lane_base + TM_DF + j0, df);
#pragma unroll
                for (int k = 0; k < 8; ++k) df[k] = mul2(df[k], d0[k]);
                {
                    f2 dz[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) dz[k] = df[k];
                    gn_backward16(dz, zh0, rs0, s_g0w + j0);
                    mma_ok &= mbar_wait(&bars[4], parity);    // critic wgrad complete -> DZ free again
                    store_pieces16(sDZ, r, q, dz);
                    staged(&sbar[3]);
                }
                {   // column sums under the trunk-wgrad MMAs
                    f2 t[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) t[k] = mul2(df[k], zh0[k]);
                    colsum16(t, acc_g0);
                    colsum16(df, acc_b0);
                }
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
        float *s_red = reinterpret_cast<float *>(sW);                 // [4 rq][NQ][64] column sums (the W region is free: every MMA has completed)
        const int NQ = tc_num_q(L);
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
        }
        // column sums: complete the butterflies (all lanes), then one of the four lanes holding feature f of a group reports it;
        // quantities: 0 dgamma0, 1 dbeta0, then per head: dgamma, dbeta, dW2[a] (a < out)
        colsum_finish(acc_g0); colsum_finish(acc_b0); colsum_finish(acc_gh[0]); colsum_finish(acc_bh[0]);
        colsum_finish(acc_gh[1]); colsum_finish(acc_bh[1]); colsum_finish(acc_w2c);
#pragma unroll
        for (int a = 0; a < NA; ++a) colsum_finish(acc_w2a[a]);
        if (softmax_rows_sum_to_zero) {
            // the softmax head's output gradients of a row sum to zero (dout_a = dlogp (1[a = act] - p_a), sum_a p_a = 1), so the
            // last action's dW2 row and db2 are minus the sum of the others': its column sums were skipped in the tile loop
            f2 w = dup2(0.f);
            float bsum = 0.f;
#pragma unroll
            for (int a = 0; a < NA; ++a)
                if (a < nout0 - 1) { w = add2(w, acc_w2a[a]); bsum += b2a[a]; }
#pragma unroll
            for (int a = 0; a < NA; ++a)
                if (a == nout0 - 1) { acc_w2a[a] = mk2(-w.x, -w.y); b2a[a] = -bsum; }
        }
        if ((lane & 3) == 0) {
            const int f = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            float *dst = s_red + (size_t)rq * NQ * HID + j0 + f;
            auto rep = [&](int qi, f2 a) { dst[qi * HID] = a.x; dst[qi * HID + 8] = a.y; };
            rep(0, acc_g0); rep(1, acc_b0);
            rep(2, acc_gh[0]); rep(3, acc_bh[0]);
#pragma unroll
            for (int a = 0; a < NA; ++a)
                if (a < nout0) rep(4 + a, acc_w2a[a]);
            rep(4 + nout0, acc_gh[1]); rep(5 + nout0, acc_bh[1]); rep(6 + nout0, acc_w2c);
        }
        if (q == 0) {
            // db2 and the loss sums of this row quarter (fixed shuffle tree)
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const float sb = warp_sum(b2a[a]);
                if (lane == 0) b2s[rq][a] = sb;
            }
            const float sc = warp_sum(b2c);
            const double sp = warp_sum((double)l_pol), sv = warp_sum((double)l_val), se = warp_sum((double)l_ent);
            if (lane == 0) { b2s[rq][TC_MAX_A] = sc; red[0][rq] = sp; red[1][rq] = sv; red[2][rq] = se; }
        }
        bar_compute();
        for (int idx = tid; idx < HID * HID; idx += TC_COMPUTE) {
            const int j = idx >> 6, k = idx & 63;
            part[L.head[0].w1 + idx] = (it > 0) ? scratch0[j * SS + k] + scratch0[(64 + j) * SS + k] : 0.f;
            part[L.head[1].w1 + idx] = (it > 0 && !skip1) ? scratch1[j * SS + k] + scratch1[(64 + j) * SS + k] : 0.f;
        }
        for (int idx = tid; idx < HID * O; idx += TC_COMPUTE) {
            const int j = idx / O, i = idx - j * O;
            part[L.w0 + idx] = (it > 0) ? scratch2[j * 17 + i] + scratch2[(64 + j) * 17 + i] : 0.f;
        }
        for (int idx = tid; idx < NQ * HID; idx += TC_COMPUTE) {
            const int qq = idx / HID, j = idx - qq * HID;
            const float sm = (s_red[(0 * NQ + qq) * HID + j] + s_red[(1 * NQ + qq) * HID + j]) + (s_red[(2 * NQ + qq) * HID + j] + s_red[(3 * NQ + qq) * HID + j]);
            int off;
            if (qq == 0) off = L.g0w;
            else if (qq == 1) off = L.g0b;
            else {
                int k = qq - 2, h = 0;
                if (k >= 2 + nout0) { k -= 2 + nout0; h = 1; }
                const int hgw = h ? L.head[1].gw : L.head[0].gw, hgb = h ? L.head[1].gb : L.head[0].gb, hw2 = h ? L.head[1].w2 : L.head[0].w2;
                off = (k == 0) ? hgw : (k == 1) ? hgb : hw2 + (k - 2) * HID;
            }
            part[off + j] = sm;
        }
        for (int idx = tid; idx < ext.zero_len; idx += TC_COMPUTE) part[ext.zero_off + idx] = 0.f;
        if (tid < NA) {
            if (tid < nout0) part[L.head[0].b2 + tid] = (b2s[0][tid] + b2s[1][tid]) + (b2s[2][tid] + b2s[3][tid]);
        } else if (tid == NA) {
            part[L.head[1].b2] = (b2s[0][TC_MAX_A] + b2s[1][TC_MAX_A]) + (b2s[2][TC_MAX_A] + b2s[3][TC_MAX_A]);
        } else if (tid >= 32 && tid < 35) {
            const int k = tid - 32;
            loss_partials[blockIdx.x * 4 + k] = (red[k][0] + red[k][1]) + (red[k][2] + red[k][3]);
        }
        TC_STAMP(15);
    }
    if (!mma_ok && lane == 0) atomicExch(status, 1);
    fence_before_sync();
    __syncthreads();   // every thread's partial-row stores happen before the READY flag below
    if (is_mma_warp) tmem_dealloc(D.tmem, TM_COLS);
    TC_STAMP(16);
    TC_SPAN(2);
    if (!fused) { TC_SPAN(3); return; }

    // ================= fused optimiser tail: reduce -> clip_grad_norm_ -> AdamW on this CTA's slices of the parameters.
    // No grid barrier anywhere: READY flags behind the partial rows, tagged words for the squared norm (all CTAs are
    // co-resident - cooperative launch - so the waits terminate; they are bounded in wall-clock time anyway: status 2).
    const int nb = gridDim.x;
    const bool leader = blockIdx.x == (unsigned)nb - 1;
    if (tid == 0) {
        red_release_gpu_add(opt.count, 1u);
        TC_SPAN(4);
        if (!wait_count(opt.count, (unsigned)nb, opt.timeout_ns)) tail_ok_s = 0;
    }
    __syncthreads();
    TC_SPAN(5);
    bool tags_ok = tail_ok_s != 0;
    {   // this step's bias-correction powers (kept off the critical path above)
        const bool have = opt_step > 0 && opt_p1 > 0.0;
        opt_step += 1;
        // (first step, or a clock loaded without its powers: beta^step by binary exponentiation - an inlined double-precision pow() is
        // ~6 KB of code in a kernel whose instruction stream is already larger than it should be; A/B: 0.6 %)
        opt_p1 = have ? opt_p1 * 0.9 : ipow(0.9, opt_step);
        opt_p2 = have ? opt_p2 * 0.999 : ipow(0.999, opt_step);
    }
    // The parameters are cut into slices of RC = 64 (142 slices for P = 9 027); CTA c owns slices c, c + nb, ...  The cut does
    // not depend on the grid, so ranks whose minibatches have different row counts (different grids) agree on it.
    constexpr int RC = 64;                                            // parameters per slice
    const int nsl = (P + RC - 1) / RC;
    float *sl_part = reinterpret_cast<float *>(smem_raw);             // [RED_SL][RC] slice sums
    double *sq = reinterpret_cast<double *>(smem_raw + 8192);         // [RC] squared gradients
    double ssum = 0.0;                                                // thread 0: squared norm of this CTA's slices
    const bool sharded = SHARDED && opt.world > 1;
    // losses of the whole launch (one warp of CTA 0)
    if (blockIdx.x == 0 && is_mma_warp && opt.loss_out && tags_ok) add_loss_sums(loss_partials, nb, opt.loss_out, opt.rows, lane);
    // sharded: exchange buffer of rank r = inbox[2 (step parity)][world (sender)][gstride] 8-byte words {float bits, step number}
    const unsigned int xepoch = (unsigned int)opt_step;   // the ranks agree on the optimiser step, not on launch counters
    const int inbox_off = (int)(opt_step & 1) * opt.world * opt.gstride;   // this step's inbox, in words
    for (int sidx = blockIdx.x; sidx < nsl && tags_ok; sidx += nb) {
        const int p0 = sidx * RC, nc = min(RC, P - p0);
        for (int item = tid; item < nc * RED_SL; item += TC_THREADS) {
            const int sl = item / nc, pi = item - sl * nc;
            sl_part[sl * RC + pi] = reduce_slice(partials, nb, part_stride, p0 + pi, sl);
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
                const unsigned long long t0 = globaltimer_ns();
                for (int r0 = 0; r0 < opt.world; r0 += 8) {   // 8 ranks polled together, summed in rank order
                    unsigned long long pv[8];
                    bool all = false;
                    for (int itp = 0; !all; ++itp) {
                        all = true;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            pv[u] = r0 + u < opt.world ? ld_relaxed_sys_u64(in + (r0 + u) * opt.gstride) : ((unsigned long long)xepoch << 32);
                            all = all && (unsigned int)(pv[u] >> 32) == xepoch;
                        }
                        if (!all && (itp & 63) == 63 && globaltimer_ns() - t0 > opt.timeout_ns) break;
                    }
                    ok = ok && all;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (r0 + u < opt.world) g += __uint_as_float((unsigned int)pv[u]);
                }
                if (!ok) tail_ok_s = 0;
                gi = g;
            }
            __syncthreads();
            if (!tail_ok_s) { tags_ok = false; if (tid == 0) atomicExch(status, 3); break; }
        }
        if (tid < nc) {
            opt.grad[p0 + tid] = gi;
            sq[tid] = (double)gi * gi;
        }
        __syncthreads();
        // the 64 squared gradients of the slice: one warp, fixed tree (A/B: 0.4 % over one thread adding them serially)
        if (warp == 0) {
            double a = (lane < nc ? sq[lane] : 0.0) + (lane + 32 < nc ? sq[lane + 32] : 0.0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) ssum += a;
        }
        __syncthreads();
    }
    // a CTA that failed publishes nothing: the others time out on its squared norm, nobody applies the step
    if (tid == 0 && tags_ok) st_tag_double(opt.tagw + blockIdx.x * 2, ssum, epoch);
    TC_SPAN(6);
    if (warp == 0) {
        if (leader) {
            // total squared norm in a fixed order; the leader also advances the optimiser clock and the launch counter and
            // re-arms the arrival counter (every CTA has passed it: it published its squared norm afterwards)
            const double a = tags_ok ? sum_tagged_doubles(opt.tagw, nb, epoch, lane, opt.timeout_ns, tags_ok) : 0.0;
            if (lane == 0) {
                const float total = (float)sqrt(a);
                const float coef = opt.max_norm > 0.f ? fminf(opt.max_norm / (total + 1e-6f), 1.0f) : 1.f;
                coef_s = coef;
                if (!tags_ok) tail_ok_s = 0;
                else {
                    st_tag(opt.coefw, __float_as_uint(coef), epoch);
                    if (opt.norm_out) *opt.norm_out = (double)total;
                    double *pw = reinterpret_cast<double *>(opt.clock) + 1;
                    opt.clock[0] = opt_step; pw[0] = opt_p1; pw[1] = opt_p2;
                    reinterpret_cast<unsigned int *>(status)[3] = epoch;   // the launch counter (every CTA read it at its start)
                    *opt.count = 0u;
                }
            }
        } else if (lane == 0) {
            if (tags_ok) coef_s = __uint_as_float(wait_tag(opt.coefw, epoch, opt.timeout_ns, tags_ok));
            if (!tags_ok) tail_ok_s = 0;
        }
    }
    __syncthreads();
    TC_SPAN(7);
    if (!tail_ok_s) {   // some wait timed out: the step is NOT applied by this CTA (status 2 unless a peer wait already said 3)
        if (tid == 0) atomicCAS(status, 0, 2);
        TC_SPAN(3);
        return;
    }
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
                     const double *__restrict__ loss_partials, double *__restrict__ loss_out, double rows, int accumulate) {
    __shared__ float part[RED_SL][64];
    const int p = threadIdx.x & 63, sl = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + p;
    part[sl][p] = i < P ? reduce_slice(partials, nblocks, stride, i, sl) : 0.f;
    __syncthreads();
    if (sl == 0 && i < P) {
        float t[RED_SL];
#pragma unroll
        for (int u = 0; u < RED_SL; ++u) t[u] = part[u][p];
        const float g = reduce_tree(t);
        grad[i] = accumulate ? grad[i] + g : g;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32 && loss_out) add_loss_sums(loss_partials, nblocks, loss_out, rows, threadIdx.x);
}

// ---- pre-pass of the continuous (tanh-Gaussian) update: forward of trunk + mu head + log_std head in IEEE float32 (the
// register-tiled forward of the old-policy evaluation, tiled_mlp.cuh), the clipped-surrogate loss of every row and its gradient
// with respect to the two heads' outputs - the operation sequence of csrc/update_ppo.cu's continuous branch (PPO.py:219-252,
// ActorCritic.py:118-146).  dout_mu / dout_ls [b][A]; loss_partials[block][4] = {policy term sum, 0, entropy sum, 0}.
// RPT: rows of the forward per thread, 32 RPT rows per tile (tiled_mlp.cuh) - 7 where 224-row tiles spread more evenly over the
// persistent CTAs than 256-row ones (262 144 rows on 296 CTAs: 3.96 tiles of 224 each instead of 3.46 -> 4 tiles of 256).
template <int RPT>
__global__ void __launch_bounds__(EV_THREADS, 2)
k_policy_dout(const float *__restrict__ params, PolicyLayout L, const float *__restrict__ states, const float *__restrict__ actions,
              const float *__restrict__ old_logp, const float *__restrict__ adv, int64_t n, float clip, float inv_count,
              float *__restrict__ dout_mu, float *__restrict__ dout_ls, double *__restrict__ loss_partials) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double red[32];
    const EvSmem S = ev_layout(L, 2);   // mu and log_std heads (the critic's loss needs no pre-pass)
    const int tid = threadIdx.x, O = L.O, A = L.A;
    ev_stage_weights(smem, S, params, L);
    float *sX = smem + S.x;
    const float *sO = smem + S.out;
    constexpr int ROWS = 32 * RPT;
    const int64_t ntiles = (n + ROWS - 1) / ROWS;
    double pol_acc = 0.0, ent_acc = 0.0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * ROWS;
        const int rows = (int)min((int64_t)ROWS, n - row0);
        __syncthreads();
        for (int i = tid; i < ROWS * O; i += EV_THREADS) sX[i] = i < rows * O ? __ldg(states + row0 * O + i) : 0.f;
        __syncthreads();
        ev_forward_tile<RPT>(smem, S, L);
        __syncthreads();
        if (tid < rows) {
            const int64_t row = row0 + tid;
            const float *o = sO + tid * S.so;
            // From the float32 head outputs on, the loss and its output gradients are evaluated in DOUBLE: ratio = exp(logp - old_logp)
            // carries ulp(|logp|) of float32 noise per row otherwise, which the Gaussian log-density's (z^2 - 1) / sigma factors
            // amplify (measured on the learn_continuous rows: 5e-5 of the largest gradient component in float32, torch-float32
            // autograd itself 1e-5; a few dozen double operations per row cost nothing next to the forward)
            double q = 0.0, hld = 0.0;
            for (int a = 0; a < A; ++a) {
                const double mu = (double)o[S.col[0] + a], ls = (double)o[S.col[1] + a];
                const double lc = fmin(fmax(ls, -2.0), 2.0);
                const double tril = log1p(exp(lc));            // softplus; sqrt(sd^2) = sd > 0
                const double zt = ((double)actions[row * A + a] - mu) / tril;
                q = fma(zt, zt, q);
                hld += log(tril);
            }
            const double logp = -0.5 * (A * 1.8378770664093454836 + q) - hld;
            // d(-min(r A, clamp(r) A) * inv_count) / dlogp
            const double adv_i = (double)adv[row], dl = logp - (double)old_logp[row], lo = 1.0 - (double)clip, hi = 1.0 + (double)clip;
            const double r = exp(fmin(fmax(dl, -20.0), 20.0));
            const double s1 = r * adv_i, s2 = fmin(fmax(r, lo), hi) * adv_i;
            const double g1 = s1 < s2 ? 1.0 : (s1 > s2 ? 0.0 : 0.5);   // torch.min splits ties evenly
            const double in_clip = (r >= lo && r <= hi) ? 1.0 : 0.0;
            const double in20 = (dl >= -20.0 && dl <= 20.0) ? 1.0 : 0.0;
            const double dlogp = -(double)inv_count * adv_i * (g1 + (1.0 - g1) * in_clip) * r * in20;
            for (int a = 0; a < A; ++a) {
                const double mu = (double)o[S.col[0] + a], ls = (double)o[S.col[1] + a];
                const double lc = fmin(fmax(ls, -2.0), 2.0);
                const double tril = log1p(exp(lc));
                const double zt = ((double)actions[row * A + a] - mu) / tril;
                const double in2 = (ls >= -2.0 && ls <= 2.0) ? 1.0 : 0.0;
                dout_mu[row * A + a] = (float)(dlogp * zt / tril);
                dout_ls[row * A + a] = (float)(dlogp * (zt * zt - 1.0) / tril * (1.0 / (1.0 + exp(-lc))) * in2);
            }
            pol_acc += -fmin(s1, s2);
            ent_acc += 0.5 * A * (1.0 + 1.8378770664093454836) + hld;
        }
    }
    const double bp = block_sum<double>(pol_acc, red);
    const double be = block_sum<double>(ent_acc, red);
    if (tid == 0) {
        loss_partials[blockIdx.x * 4 + 0] = bp; loss_partials[blockIdx.x * 4 + 1] = 0.0;
        loss_partials[blockIdx.x * 4 + 2] = be; loss_partials[blockIdx.x * 4 + 3] = 0.0;
    }
}
__global__ void k_add_loss_sums(const double *__restrict__ loss_partials, int nblocks, double *__restrict__ loss_out) {
    add_loss_sums(loss_partials, nblocks, loss_out, 0.0, threadIdx.x);
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
// Fused optimiser step: the tail (slice reduction, squared norm, clip + AdamW) is spread over the CTAs of the launch, one
// 64-parameter slice after the other per CTA - a small minibatch (512 rows = 4 CTAs) would walk 36 slices per CTA, ~1 us each.
// So the launch is widened to one CTA per slice plus the leader: the CTAs beyond those that own rows stage nothing but the
// tail's share (they write an all-zero partial row, which changes no bit of the fixed-order sums: x + 0 = x, and row b stays in
// reduction slice b % RED_SL whatever the grid is).  configs[0] (mini_batch 512): 66 -> ~30 us per optimiser step.
static int tc_tail_grid(int grid_rows, int P) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = (P + 63) / 64 + 1;
    const int g = want < sms ? want : sms;
    return grid_rows > g ? grid_rows : g;
}
// upper bound of the grid over every minibatch of at most b rows (the workspace is sized once for the largest one)
static int tc_grid_max(int64_t b, int P) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t nt = (b + TC_ROWS - 1) / TC_ROWS;
    return tc_tail_grid((int)(nt < sms ? (nt > 0 ? nt : 1) : sms), P);
}
// header (4 words: status, arrival counter, unused, launch counter) | partial rows (floats, 16-byte aligned rows) | loss partials
// (4 doubles per CTA) | squared-norm words (2 tagged 8-byte words per CTA) | clip-coefficient words (2)
static size_t tc_ws_floats(const PolicyLayout &L, int grid) {
    return 4 + (size_t)grid * ((L.total + 3) & ~3) + (size_t)grid * 8 + (size_t)grid * 4 + (size_t)grid + 16;
}
// bound of the in-kernel waits on other CTAs / other GPUs, wall clock (PRL_TC_TIMEOUT_MS, default 20 s: a stalled peer process
// - profiler, lazy module load - must not be mistaken for a dead one)
static unsigned long long tc_timeout_ns() {
    static unsigned long long ns = 0;
    if (!ns) {
        const char *e = getenv("PRL_TC_TIMEOUT_MS");
        const double ms = e ? atof(e) : 20000.0;
        ns = (unsigned long long)((ms > 0.0 ? ms : 20000.0) * 1e6);
    }
    return ns;
}

}  // namespace prl

using namespace prl;

extern "C" {

// 1: discrete policy - every form (prl_ppo_grad_tc, prl_ppo_step_tc, prl_ppo_step_tc_p2p); 2: continuous policy - prl_ppo_grad_tc
// only (pre-pass + two passes of the two-head kernel); 0: not supported
int prl_ppo_grad_tc_supported(int is_continuous, int obs_dim, int action_dim) {
    const bool fits = obs_dim >= 1 && obs_dim <= TC_MAX_O && action_dim >= 1 && action_dim <= TC_MAX_A;
    return fits ? (is_continuous ? 2 : 1) : 0;
}

// continuous: behind the kernel's workspace sit d loss / d mu and d loss / d log_std of every row ([b][A] each) and the pre-pass's
// loss partials (4 doubles per CTA, at most 2 CTAs per SM)
static size_t tc_cont_extra_floats(int action_dim, int64_t batch) { return (size_t)2 * action_dim * batch + 8 * 2 * 160 + 16; }

size_t prl_update_tc_ws_floats(int is_continuous, int obs_dim, int action_dim, int64_t batch) {
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    return tc_ws_floats(L, tc_grid_max(batch, L.total)) + (is_continuous ? tc_cont_extra_floats(action_dim, batch) : 0);
}

// shared launcher: gradient only (opt == nullptr: + separate reduction kernel) or fused optimiser step
// L: the two heads the kernel works on (a discrete policy's own layout, or two of a continuous policy's three heads inside the
// full parameter vector); ext: where head 0's output gradient comes from; accumulate: add the reduced gradient to `grad`
static int launch_tc(const float *params, const PolicyLayout &L, const float *states, const float *actions,
                     const float *old_logp, const float *adv, const float *returns, int64_t b, float policy_clip, float inv_count,
                     float *grad, double *loss_out, float *ws, size_t ws_floats, cudaStream_t st, const TcOptimizer *optp, const char *who,
                     TcExternal ext = TcExternal{nullptr, 1.0f, 0, 0}, int accumulate = 0) {
    PRL_REQUIRE(params && grad && ws && b >= 0 && (b > 0 || (optp && optp->world > 1)), "%s: bad arguments", who);
    PRL_REQUIRE(b == 0 || (states && returns && (ext.dout || (actions && old_logp && adv))), "%s: null row pointer", who);
    const int obs_dim = L.O, action_dim = L.A;
    PRL_REQUIRE(prl_ppo_grad_tc_supported(0, obs_dim, action_dim) && L.n_heads == 2,
                "%s: policies with observ_dim <= %d and action_dim <= %d only (got O=%d A=%d)", who, TC_MAX_O, TC_MAX_A, obs_dim, action_dim);
    int grid, qpc;
    tc_split(b, &grid, &qpc);
    if (optp && !getenv("PRL_TC_NARROW_TAIL")) grid = tc_tail_grid(grid, L.total);   // (the variable: A/B against the row CTAs only)
    const int pstride = (L.total + 3) & ~3;   // per-CTA partial rows start 16-byte aligned
    PRL_REQUIRE(ws_floats >= tc_ws_floats(L, grid), "%s: workspace too small", who);
    PRL_REQUIRE(((uintptr_t)ws & 15) == 0, "%s: workspace must be 16-byte aligned", who);
    const int NA = action_dim <= 2 ? 2 : (action_dim == 3 && obs_dim > 4 && obs_dim <= 6) ? 3 : action_dim <= 4 ? 4 : 8;
    const size_t smem = tc_smem_bytes(L, NA);
    PRL_REQUIRE(smem <= 227 * 1024, "%s: needs %zu B shared memory (> 227 KB)", who, smem);
    PRL_REQUIRE(grid <= TC_MAX_GRID, "%s: grid %d > %d", who, grid, TC_MAX_GRID);
    // ws: [0] sticky status word, [1] arrival counter, [3] launch counter (the caller zeroes the workspace once), then per-CTA
    // partial gradients, loss partials, squared-norm words, clip-coefficient words
    int *status = reinterpret_cast<int *>(ws);
    float *partials = ws + 4;
    double *loss_partials = reinterpret_cast<double *>(partials + (size_t)grid * pstride);
    TcOptimizer opt{};
    if (optp) {
        opt = *optp;
        opt.tagw = reinterpret_cast<unsigned long long *>(loss_partials + (size_t)grid * 4);
        opt.coefw = opt.tagw + (size_t)grid * 2;
        opt.count = reinterpret_cast<unsigned int *>(ws) + 1;
        opt.timeout_ns = tc_timeout_ns();
    }
    auto launch = [&](auto kernel) -> int {
        PRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;   // the fused tail waits on the other CTAs: every CTA must be resident
        attr[0].val.cooperative = optp ? 1 : 0;
        cfg.attrs = attr; cfg.numAttrs = 1;
        PRL_CUDA(cudaLaunchKernelEx(&cfg, kernel, params, L, states, actions, old_logp, adv, returns, b, policy_clip, inv_count, partials, pstride,
                                    loss_partials, status, opt, qpc, ext));
        return PRL_OK;
    };
    // builds: actor width NA x observation registers XR {(2,4) (2,8) (3,6) (4,4) (4,8) (8,8)} x {single GPU | sharded | second pass of a
    // continuous update (no critic head)}.  (3,6) exists for Acrobot's shapes (A = 3, O = 6): A/B 66.0 against 69.1 us per launch with (4,8).
    const bool x4 = obs_dim <= 4;
    const bool v36 = action_dim == 3 && obs_dim > 4 && obs_dim <= 6;
#define PRL_TC_PICK(SH, SK)                                                                                                \
    (v36 ? launch(k_ppo_grad_tc<3, 6, SH, SK>)                                                                             \
     : NA == 2 ? (x4 ? launch(k_ppo_grad_tc<2, 4, SH, SK>) : launch(k_ppo_grad_tc<2, 8, SH, SK>))                          \
     : NA == 4 ? (x4 ? launch(k_ppo_grad_tc<4, 4, SH, SK>) : launch(k_ppo_grad_tc<4, 8, SH, SK>))                          \
               : launch(k_ppo_grad_tc<8, 8, SH, SK>))
    int rc;
    if (optp && optp->world > 1) rc = PRL_TC_PICK(true, false);
    else if (ext.dout != nullptr && ext.critic_weight == 0.f) rc = PRL_TC_PICK(false, true);
    else rc = PRL_TC_PICK(false, false);
#undef PRL_TC_PICK
    if (rc != PRL_OK) return rc;
    if (getenv("PRL_TC_TIMING")) {
        long long c[32];
        PRL_CUDA(cudaStreamSynchronize(st));
        PRL_CUDA(cudaMemcpyFromSymbol(c, g_tc_clock, sizeof c));
        fprintf(stderr, "[%s] CTA0 thread 0 cycles: setup %lld (W1 loads issued %lld, W1 row arrived %lld, parameters in shared memory %lld, pieces stored %lld) | tile 0: trunk fwd + stage F %lld, wait fwd MMA %lld, actor epilogue %lld, stage DZ %lld, "
                        "critic epilogue %lld, wait actor MMA + stage DZ %lld, wait critic MMA %lld, trunk bwd + stage %lld | all tiles %lld, final MMA wait %lld, "
                        "readout %lld, tail %lld | kernel %lld\n", who,
                c[1] - c[0], c[19] - c[0], c[20] - c[0], c[17] - c[0], c[18] - c[0], c[2] - c[1], c[4] - c[2], c[5] - c[4], c[6] - c[5], c[8] - c[6], c[9] - c[8], c[11] - c[9], c[12] - c[11], c[13] - c[1],
                c[14] - c[13], c[15] - c[14], c[16] - c[15], c[16] - c[0]);
        static unsigned long long sp[160 * 8];
        PRL_CUDA(cudaMemcpyFromSymbol(sp, g_tc_span, sizeof sp));
        unsigned long long t0 = ~0ull, mn[8], mx[8];
        for (int k = 0; k < 8; ++k) { mn[k] = ~0ull; mx[k] = 0; }
        for (int c2 = 0; c2 < grid && c2 < 160; ++c2) t0 = sp[c2 * 8] < t0 ? sp[c2 * 8] : t0;
        for (int c2 = 0; c2 < grid && c2 < 160; ++c2)
            for (int k = 0; k < 8; ++k) {
                const unsigned long long v = sp[c2 * 8 + k] - t0;
                mn[k] = v < mn[k] ? v : mn[k]; mx[k] = v > mx[k] ? v : mx[k];
            }
        fprintf(stderr, "[%s] grid %d wall clock, ns from the first CTA start, min..max over CTAs: start %llu..%llu | tiles done %llu..%llu | partials written "
                        "%llu..%llu | kernel end %llu..%llu\n", who, grid, mn[0], mx[0], mn[1], mx[1], mn[2], mx[2], mn[3], mx[3]);
        if (optp) fprintf(stderr, "[%s] tail: arrival counted %llu..%llu | all arrivals seen %llu..%llu | slices reduced %llu..%llu | norm known %llu..%llu\n", who,
                          mn[4], mx[4], mn[5], mx[5], mn[6], mx[6], mn[7], mx[7]);
    }
    if (!optp) k_reduce_partials_tc<<<cdiv(L.total, 64), 64 * RED_SL, 0, st>>>(partials, grid, L.total, pstride, grad, loss_partials, loss_out,
                                                                               ext.critic_weight != 0.f ? (double)b : 0.0, accumulate);
    return check_launch(who);
}

int prl_ppo_grad_tc(const float *params, int is_continuous, int obs_dim, int action_dim, const float *states, const float *actions,
                    const float *old_logp, const float *adv, const float *returns, int64_t b, float policy_clip, float inv_count,
                    float *grad, double *loss_out, float *ws, size_t ws_floats, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!is_continuous)
        return launch_tc(params, make_policy_layout(0, obs_dim, action_dim), states, actions, old_logp, adv, returns, b, policy_clip, inv_count,
                         grad, loss_out, ws, ws_floats, st, nullptr, "prl_ppo_grad_tc");
    // ---- continuous policy: pre-pass (loss + output gradients of the mu / log_std heads), then two passes of the two-head kernel
    const char *who = "prl_ppo_grad_tc (continuous)";
    PRL_REQUIRE(prl_ppo_grad_tc_supported(1, obs_dim, action_dim), "%s: observ_dim <= %d and action_dim <= %d only", who, TC_MAX_O, TC_MAX_A);
    PRL_REQUIRE(params && states && actions && old_logp && adv && returns && grad && ws && b > 0, "%s: bad arguments", who);
    const PolicyLayout LC = make_policy_layout(1, obs_dim, action_dim);
    const size_t base = tc_ws_floats(LC, tc_grid_max(b, LC.total));
    PRL_REQUIRE(ws_floats >= base + tc_cont_extra_floats(action_dim, b) && ((uintptr_t)ws & 15) == 0, "%s: workspace too small / unaligned", who);
    const size_t off_mu = (base + 3) & ~(size_t)3, off_ls = off_mu + (size_t)action_dim * b, off_pre = (off_ls + (size_t)action_dim * b + 3) & ~(size_t)3;
    float *dmu = ws + off_mu, *dls = ws + off_ls;
    double *pre_partials = reinterpret_cast<double *>(ws + off_pre);   // (ws is 16-byte aligned, off_pre a multiple of 4 floats)
    {
        const size_t smem = (size_t)ev_layout(LC, 2).total * sizeof(float);
        PRL_REQUIRE(smem <= 227 * 1024, "%s: pre-pass needs %zu B shared memory", who, smem);
        int pgrid = 1;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int slots = 2 * sms < 320 ? 2 * sms : 320;   // (pre_partials holds 320 CTAs' sums)
        auto rounds = [&](int rpt) { return cdiv(cdiv(b, (int64_t)32 * rpt), (int64_t)slots) * rpt; };   // tiles per CTA x their size
        const int rpt = rounds(7) < rounds(8) && !getenv("PRL_DOUT_RPT8") ? 7 : 8;
        const int64_t ntiles = cdiv(b, (int64_t)32 * rpt);
        pgrid = (int)(ntiles < slots ? ntiles : slots);
        if (rpt == 7) {
            PRL_CUDA(cudaFuncSetAttribute(k_policy_dout<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_policy_dout<7><<<pgrid, EV_THREADS, smem, st>>>(params, LC, states, actions, old_logp, adv, b, policy_clip, inv_count, dmu, dls, pre_partials);
        } else {
            PRL_CUDA(cudaFuncSetAttribute(k_policy_dout<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_policy_dout<8><<<pgrid, EV_THREADS, smem, st>>>(params, LC, states, actions, old_logp, adv, b, policy_clip, inv_count, dmu, dls, pre_partials);
        }
        if (loss_out) k_add_loss_sums<<<1, 32, 0, st>>>(pre_partials, pgrid, loss_out);
        if (check_launch("k_policy_dout") != PRL_OK) return PRL_ERR_CUDA;
    }
    // the kernel's view of the parameter vector: trunk + {one policy head, critic}
    auto two_heads = [&](int h) {
        PolicyLayout L2 = LC;
        L2.cont = 0; L2.n_heads = 2;
        L2.head[0] = LC.head[h]; L2.head[1] = LC.head[2];
        return L2;
    };
    const int span = HID * HID + 2 * HID + action_dim * HID + action_dim;   // parameters of one policy head (contiguous: w1 gw gb w2 b2)
    int rc = launch_tc(params, two_heads(0), states, nullptr, nullptr, nullptr, returns, b, policy_clip, inv_count, grad, loss_out, ws, base, st, nullptr, who,
                       TcExternal{dmu, 1.0f, LC.head[1].w1, span}, 0);
    if (rc != PRL_OK) return rc;
    return launch_tc(params, two_heads(1), states, nullptr, nullptr, nullptr, returns, b, policy_clip, inv_count, grad, nullptr, ws, base, st, nullptr, who,
                     TcExternal{dls, 0.0f, LC.head[0].w1, span}, 1);
}

int prl_ppo_step_tc(float *params, int is_continuous, int obs_dim, int action_dim, const float *states, const float *actions,
                    const float *old_logp, const float *adv, const float *returns, int64_t b, float policy_clip, float inv_count,
                    float *grad, double *loss_out, float *exp_avg, float *exp_avg_sq, int64_t *step_counter, float lr, float weight_decay,
                    float max_norm, double *grad_norm_out, float *ws, size_t ws_floats, void *stream) {
    PRL_REQUIRE(exp_avg && exp_avg_sq && step_counter, "prl_ppo_step_tc: bad optimiser arguments");
    TcOptimizer opt{};
    opt.params_rw = params; opt.grad = grad; opt.m = exp_avg; opt.v = exp_avg_sq; opt.clock = step_counter; opt.norm_out = grad_norm_out;
    opt.lr = lr; opt.wd = weight_decay; opt.max_norm = max_norm; opt.loss_out = loss_out; opt.rows = (double)b;
    PRL_REQUIRE(!is_continuous, "prl_ppo_step_tc: discrete policies only (continuous ones: prl_ppo_grad_tc + prl_adamw_step_dev)");
    return launch_tc(params, make_policy_layout(0, obs_dim, action_dim), states, actions, old_logp, adv, returns, b, policy_clip, inv_count, grad, loss_out,
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
    PRL_REQUIRE(!is_continuous, "prl_ppo_step_tc_p2p: discrete policies only (continuous ones: prl_ppo_grad_tc + allreduce + prl_adamw_step_dev)");
    return launch_tc(params, L, states, actions, old_logp, adv, returns, b, policy_clip, inv_count, grad, loss_out,
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
    // a timed-out launch leaves its arrival count behind: clear it so that the workspace can be used again
    if (*status_host == 2 || *status_host == 3)
        PRL_CUDA(cudaMemsetAsync(const_cast<float *>(ws) + 1, 0, sizeof(unsigned int), (cudaStream_t)stream));
    return PRL_OK;
}

}  // extern "C"
