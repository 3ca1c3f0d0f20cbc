// GAE / returns reverse scan and advantage normalisation.
//
// Reference: PPO.compute_gae (/root/reference/PPO/PPO.py:107-120) - a flat reverse loop over the env-major
// buffer in float32 (numpy-2 promotion: gamma and gamma*lambda are rounded to float32 once, every intermediate
// rounds to float32, evaluation order of lines 113-114) - and the normalisation at PPO.py:198-199.
//
// Bit-exactness: gae_t = delta_t + (gl * nd_t) * gae_{t+1} is not associative in floating point, so the scan is
// parallel ACROSS segments (a segment ends where done == 1, which zeroes both the bootstrap and the carry) and
// strictly sequential WITHIN one, in the reference's operation order.  Two forms:
//   prl_gae_columns  time-major [T][E]: one env per thread walking t backwards - fully coalesced across envs.
//   prl_gae          flat env-major [N] (the compute_gae signature): one segment per lane, staged through shared memory.
#include <stdlib.h>

#include "common.cuh"

namespace prl {

__device__ __forceinline__ float gae_step(float r, float d, float v, float nv, float g, float gl, float &gae) {
    const float nd = __fsub_rn(1.0f, d);
    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(g, nv), nd)), v);
    gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nd), gae));
    return __fadd_rn(gae, v);
}

// U time steps of loads are issued together before the (strictly sequential) arithmetic of those steps retires: with one
// thread per env there are only ~14 warps per SM at E = 65 536, so memory-level parallelism has to come from each thread
// (3 loads x U x 128 B per warp in flight).
constexpr int GCU = 8;
__global__ void __launch_bounds__(256)
k_gae_columns(const float *__restrict__ rewards, const float *__restrict__ dones, const float *__restrict__ values,
              const int32_t *__restrict__ lengths, int E, int T_cap, float g, float gl, float *__restrict__ returns) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int len = lengths ? min(lengths[e], T_cap) : T_cap;
    if (len <= 0) return;
    float gae = 0.f;
    float nv = values[(size_t)(len - 1) * E + e];
    for (int t0 = len - 1; t0 >= 0; t0 -= GCU) {
        float r[GCU], d[GCU], v[GCU];
#pragma unroll
        for (int u = 0; u < GCU; ++u) {
            const int t = t0 - u;
            const size_t i = (size_t)(t < 0 ? 0 : t) * E + e;
            r[u] = rewards[i]; d[u] = dones[i]; v[u] = values[i];
        }
#pragma unroll
        for (int u = 0; u < GCU; ++u) {
            const int t = t0 - u;
            if (t >= 0) {
                returns[(size_t)t * E + e] = gae_step(r[u], d[u], v[u], nv, g, gl, gae);
                nv = v[u];
            }
        }
    }
}

// ---- cp.async helpers (LDGSTS: global -> shared without a register round trip; completion tracked per thread) ----------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {   // both addresses 16-byte aligned
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- time-major form, ring-buffered ---------------------------------------------------------------------------------
// One warp = 32 neighbouring envs, one CTA = one warp (no CTA-wide synchronisation anywhere).  The warp streams its
// [T][32] column block backwards through a shared-memory ring of GR_STAGES x GR_ROWS time rows filled by 16-byte
// cp.async copies (lane l fetches envs 4(l%8)..+3 of row l/8), so ~12 KB per warp (~170 KB per SM at 14 warps) are in
// flight while the strictly sequential float32 recurrence of each env retires from registers; returns leave as 128-byte
// coalesced stores.  Needs E % 4 == 0 and 16-byte aligned arrays (k_gae_columns above serves everything else).
constexpr int GR_ROWS = 4, GR_STAGES = 8;
__global__ void __launch_bounds__(32)
k_gae_columns_ring(const float *__restrict__ rewards, const float *__restrict__ dones, const float *__restrict__ values,
                   const int32_t *__restrict__ lengths, int E, int T_cap, float g, float gl, float *__restrict__ returns) {
    __shared__ __align__(16) float ring[GR_STAGES][3][GR_ROWS][32];
    const int lane = threadIdx.x, e0 = blockIdx.x * 32, e = e0 + lane;
    const int len = e < E ? (lengths ? max(min(lengths[e], T_cap), 0) : T_cap) : 0;
    const int Tw = __reduce_max_sync(0xffffffffu, len);
    if (Tw <= 0) return;
    const int nst = (Tw + GR_ROWS - 1) / GR_ROWS;
    const int lrow = lane >> 3, lcol = (lane & 7) * 4;
    const bool col_ok = e0 + lcol < E;
    auto issue = [&](int k) {   // stage k = time rows Tw-1-4k .. Tw-4-4k (rows below 0 re-read row 0 and are ignored)
        if (k < nst && col_ok) {
            const int t = max(Tw - 1 - GR_ROWS * k - lrow, 0);
            const size_t gi = (size_t)t * E + e0 + lcol;
            const int s = k % GR_STAGES;
            cp_async16(&ring[s][0][lrow][lcol], rewards + gi);
            cp_async16(&ring[s][1][lrow][lcol], dones + gi);
            cp_async16(&ring[s][2][lrow][lcol], values + gi);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < GR_STAGES; ++k) issue(k);
    float gae = 0.f, nv = 0.f;
    for (int k = 0; k < nst; ++k) {
        cp_async_wait<GR_STAGES - 1>();
        __syncwarp();
        const int s = k % GR_STAGES;
        float r[GR_ROWS], d[GR_ROWS], v[GR_ROWS];
#pragma unroll
        for (int u = 0; u < GR_ROWS; ++u) { r[u] = ring[s][0][u][lane]; d[u] = ring[s][1][u][lane]; v[u] = ring[s][2][u][lane]; }
        __syncwarp();            // every lane has read the stage before it is refilled
        issue(k + GR_STAGES);
#pragma unroll
        for (int u = 0; u < GR_ROWS; ++u) {
            const int t = Tw - 1 - GR_ROWS * k - u;
            if (t >= 0 && t < len) {
                if (t == len - 1) nv = v[u];   // bootstrap of the last stored step = its own value (as k_gae_columns)
                returns[(size_t)t * E + e] = gae_step(r[u], d[u], v[u], nv, g, gl, gae);
                nv = v[u];
            }
        }
    }
}

// ---- flat form ---------------------------------------------------------------------------------------------------
// gae_t = delta_t + (gl * nd_t) * gae_{t+1},  delta_t = (r_t + (g * v_{t+1}) * nd_t) - v_t,  return_t = gae_t + v_t.
// Only the first recurrence is serial, and only inside a segment (a segment ends where done != 0, which zeroes both the
// bootstrap and the carry); delta and the final "+ v" are element-wise.  Inside a segment nd == 1 exactly, so the serial
// part is  gae_t = delta_t + gl * gae_{t+1}:  one multiply and one add per element, the reference's own roundings.
//
// One CTA (8 warps) per chunk of FL consecutive transitions plus a halo of the FH transitions before it (4 CTAs per SM):
//   1. every thread loads float4s of rewards / dones / values (all of its loads in flight at once), computes delta and the
//      segment-end flags in registers and stores x = delta (x = gae for segment ends) and v to shared memory;
//   2. the segment ends inside the chunk are compacted in ascending order (warp ballots + one scan of 32 block counts);
//   3. thread k walks segment k backwards in shared memory: x_t <- x_t + gl * x_{t+1}  (1 load, 2 float ops, 1 store per
//      element; 32-element blocks at base + immediate addresses);
//   4. returns = x + v leave with coalesced stores.
// A chunk owns exactly the segments that END in it.  The first of them usually begins in an earlier chunk: the halo holds
// that part, so it is walked like any other; only if the segment is longer than the halo does warp 0 continue backwards
// through further (smaller) windows, lane 0 walking.  Elements after the chunk's last end belong to a later chunk.
constexpr int FL = 3072, FH = 128, FW = FL + FH, FT = 256, FX = 256;
// shared-memory index of window element i: one pad word per 32 and per 128 elements, so that threads walking segments
// whose ends are 32, 64, 128, ... elements apart (equal-length episodes) hit different banks, and so that the scalar
// stores of float4 lanes (elements 4 l + c) are conflict-free.  fpad(i + 128 k) = fpad(i) + 133 k.
__device__ __forceinline__ int fpad(int i) { return i + (i >> 5) + (i >> 7); }
constexpr int FWP = FW + FW / 32 + FW / 128 + 2;
constexpr int FNQ = (FW / 4 + FT - 1) / FT;                       // float4 groups per thread (4); group q = tid + k FT
constexpr int FQS = 4 * FT + (4 * FT) / 32 + (4 * FT) / 128;      // shared-memory words between a thread's groups
static_assert(FH == 128 && FT == 256 && FNQ * (FT / 32) == 32, "the block bookkeeping below relies on these (all 32 block counts are written)");

// x of one element: delta, or gae for a segment end (gae_{t+1} = 0 there; the products keep the reference's NaN / -0 behaviour)
__device__ __forceinline__ float gae_x(float r, float d, float v, float nv, bool is_end, float g, float gl) {
    const float nd = __fsub_rn(1.0f, d);
    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(g, nv), nd)), v);
    return is_end ? __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nd), 0.f)) : delta;
}
// interior elements t, t-1, ..., s (window indices) of one segment, in place: x_t <- x_t + gl * gae
__device__ __forceinline__ void gae_walk(float *sx, int t, int s, float gl, float &gae) {
    while (t >= s) {
        const int b0 = t & ~31, lo = max(s, b0);       // one 32-element block: contiguous in shared memory
        float *px = sx + fpad(b0) - b0;
        if (lo == b0 && t == b0 + 31) {
            // a whole block: 32 independent loads in flight, the chain retires from registers, then 32 stores (the chain
            // never waits for shared memory; the 8-element form below interleaves a load latency into every batch)
            float x[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) x[u] = px[b0 + u];
#pragma unroll
            for (int u = 31; u >= 0; --u) {
                gae = __fadd_rn(x[u], __fmul_rn(gl, gae));
                x[u] = gae;
            }
#pragma unroll
            for (int u = 0; u < 32; ++u) px[b0 + u] = x[u];
            t = b0 - 1;
            continue;
        }
        for (; t - 7 >= lo; t -= 8) {
            float x8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x8[u] = px[t - u];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                gae = __fadd_rn(x8[u], __fmul_rn(gl, gae));
                px[t - u] = gae;
            }
        }
        for (; t >= lo; --t) {
            gae = __fadd_rn(px[t], __fmul_rn(gl, gae));
            px[t] = gae;
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(FT, 4)
k_gae_flat(const float *__restrict__ rewards, const float *__restrict__ dones, const float *__restrict__ values,
           const float *__restrict__ next_value_ptr, int64_t N, float g, float gl, float *__restrict__ returns) {
    __shared__ float sx[FWP], sv[FWP], sd[FWP];   // sd: staging / flags of the generic path, scratch of the long-segment path
    __shared__ uint16_t ends[FL];
    __shared__ int bcnt[32], s_hend;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
        const int64_t c0 = (int64_t)blockIdx.x * FL, base = c0 - FH;   // window element i = global element base + i
    const int len = (int)min((int64_t)FL, N - c0), lim = FH + len;
    const int h0 = base < 0 ? (int)-base : 0;                       // first valid window element
    const int ilast = N - 1 - base < (int64_t)lim ? (int)(N - 1 - base) : -1;   // window index of the very last transition
    const float nv_last = ilast >= 0 ? (next_value_ptr ? *next_value_ptr : values[N - 1]) : 0.f;
    unsigned nib[FNQ];   // bit c of nib[k]: element 4 q + c (q = tid + k FT) is a segment end
#pragma unroll
    for (int k = 0; k < FNQ; ++k) nib[k] = 0;
    if (VEC && base >= 0 && lim == FW) {
        // ---- interior chunk of 16-byte aligned arrays
        const float4 *gr = reinterpret_cast<const float4 *>(rewards + base), *gd = reinterpret_cast<const float4 *>(dones + base),
                     *gv = reinterpret_cast<const float4 *>(values + base);
        float4 a[FNQ], bq[FNQ], c[FNQ];
        float vx[FNQ];   // lane 31: the value after its group (lane 0 of the next warp holds it, out of shuffle reach)
#pragma unroll
        for (int k = 0; k < FNQ; ++k) {
            const int q = tid + k * FT;
            vx[k] = 0.f;
            if (q < FW / 4) {
                a[k] = __ldcs(gr + q); bq[k] = __ldcs(gd + q); c[k] = __ldcs(gv + q);
                if (lane == 31) vx[k] = 4 * q + 4 == ilast + 1 ? nv_last : values[base + 4 * q + 4];
            }
        }
        const int p0 = fpad(4 * tid);
#pragma unroll
        for (int k = 0; k < FNQ; ++k) {
            const int q = tid + k * FT;
            if (q < FW / 4) {   // warp-uniform (FW / 4 is a multiple of 32)
                float vn = __shfl_down_sync(0xffffffffu, c[k].x, 1);
                if (lane == 31) vn = vx[k];
                const int i = 4 * q, p = p0 + k * FQS;
                const bool e0 = bq[k].x != 0.f, e1 = bq[k].y != 0.f, e2 = bq[k].z != 0.f, e3 = bq[k].w != 0.f || i + 3 == ilast;
                sx[p] = gae_x(a[k].x, bq[k].x, c[k].x, c[k].y, e0, g, gl);
                sx[p + 1] = gae_x(a[k].y, bq[k].y, c[k].y, c[k].z, e1, g, gl);
                sx[p + 2] = gae_x(a[k].z, bq[k].z, c[k].z, c[k].w, e2, g, gl);
                sx[p + 3] = gae_x(a[k].w, bq[k].w, c[k].w, vn, e3, g, gl);
                sv[p] = c[k].x; sv[p + 1] = c[k].y; sv[p + 2] = c[k].z; sv[p + 3] = c[k].w;
                nib[k] = (e0 ? 1u : 0u) | (e1 ? 2u : 0u) | (e2 ? 4u : 0u) | (e3 ? 8u : 0u);
            }
        }
    } else {
        // ---- first / last chunk, unaligned arrays: scalar staging, then the same element-wise pass from shared memory
        for (int i = h0 + tid; i < lim; i += FT) {
            const int p = fpad(i);
            sx[p] = rewards[base + i]; sd[p] = dones[base + i]; sv[p] = values[base + i];
        }
        __syncthreads();
        for (int i = h0 + tid; i < lim; i += FT) {
            const int p = fpad(i);
            const float d = sd[p];
            const float nv = i == ilast ? nv_last : (i + 1 < lim ? sv[fpad(i + 1)] : values[base + i + 1]);
            const bool is_end = d != 0.f || i == ilast;
            sx[p] = gae_x(sx[p], d, sv[p], nv, is_end, g, gl);
            sd[p] = is_end ? 1.f : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < FNQ; ++k) {
            const int i = 4 * (tid + k * FT);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
                if (i + cc >= h0 && i + cc < lim && sd[fpad(i + cc)] != 0.f) nib[k] |= 1u << cc;
        }
    }
    // ---- compaction of the segment ends, ascending.  128-element block b = 8 k + w (warp w, group round k); block 0 is the
    // halo.  Per block: one warp prefix sum of the lanes' flag counts; then one scan of the 32 block counts.
    int pre[FNQ];        // ends of the block in lower lanes
    {
        int hend = -1;
#pragma unroll
        for (int k = 0; k < FNQ; ++k) {
            const int mine = __popc(nib[k]);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            pre[k] = incl - mine;
            if (lane == 31) bcnt[8 * k + w] = (k == 0 && w == 0) ? 0 : incl;
            if (k == 0 && w == 0) {   // the halo: its last end
                const unsigned any = __ballot_sync(0xffffffffu, nib[0] != 0);
                if (any) {
                    const int hl = 31 - __clz(any);
                    const unsigned hn = __shfl_sync(0xffffffffu, nib[0], hl);
                    hend = 4 * hl + 31 - __clz(hn);
                }
            }
        }
        if (tid == 0) s_hend = hend;
    }
    __syncthreads();
    int excl, ns;
    {
        const int v = bcnt[lane];
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        excl = incl - v;
        ns = __shfl_sync(0xffffffffu, incl, 31);
    }
    if (ns == 0) return;   // no segment ends here: a later chunk owns everything in this one
#pragma unroll
    for (int k = 0; k < FNQ; ++k) {
        const int off = __shfl_sync(0xffffffffu, excl, 8 * k + w) + pre[k];
        if ((k == 0 && w == 0) || nib[k] == 0) continue;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
            if (nib[k] >> cc & 1) ends[off + __popc(nib[k] & ((1u << cc) - 1))] = (uint16_t)(4 * (tid + k * FT) + cc);
    }
    __syncthreads();
    const int hend = s_hend;
    const int s_first = hend >= 0 ? hend + 1 : h0;
    const bool open = hend < 0 && base > 0;   // the first segment begins before the window
    float gae_open = 0.f;
    // ---- one segment per thread
    for (int k = tid; k < ns; k += FT) {
        const int e = ends[k], s = k ? ends[k - 1] + 1 : s_first;
        float gae = sx[fpad(e)];
        gae_walk(sx, e - 1, s, gl, gae);
        if (k == 0) gae_open = gae;
    }
    __syncthreads();
    {   // returns of [s_first, last end]: coalesced; element i + FT k sits (FT + FT/32 + FT/128) k words further
        const int last = ends[ns - 1];
        int i = s_first + tid;
        const float *px = sx + fpad(i), *pv = sv + fpad(i);
        float *go = returns + base + i;
        constexpr int ps = FT + FT / 32 + FT / 128;
        for (; i + 3 * FT <= last; i += 4 * FT, px += 4 * ps, pv += 4 * ps, go += 4 * FT) {
            const float a0 = __fadd_rn(px[0], pv[0]), a1 = __fadd_rn(px[ps], pv[ps]), a2 = __fadd_rn(px[2 * ps], pv[2 * ps]),
                        a3 = __fadd_rn(px[3 * ps], pv[3 * ps]);
            go[0] = a0; go[FT] = a1; go[2 * FT] = a2; go[3 * FT] = a3;
        }
        for (; i <= last; i += FT, px += ps, pv += ps, go += FT) *go = __fadd_rn(*px, *pv);
    }
    if (!open || w != 0) return;
    // ---- the rest of a first segment longer than the halo: further windows of FX elements; warp 0 alone, lane 0 walks.
    // The other warps may still be storing from sx[] / sv[]: the windows live in sd[], which nobody else reads any more.
    float gae = __shfl_sync(0xffffffffu, gae_open, 0);
    float nv = values[base + s_first];    // value of the element after the window's top
    for (int64_t top = base; top > 0;) {
        const int cnt2 = (int)min((int64_t)FX, top);
        const int64_t w0 = top - cnt2;
        int he = -1;
        __syncwarp();
        for (int it = 0; it * 32 < cnt2; ++it) {
            const int i = it * 32 + lane;
            const bool in = i < cnt2;
            const float dd = in ? dones[w0 + i] : 0.f;
            if (in) {
                const float v = values[w0 + i];
                const float vn = i + 1 < cnt2 ? values[w0 + i + 1] : nv;
                sd[fpad(i)] = gae_x(rewards[w0 + i], 0.f, v, vn, false, g, gl);   // delta (only elements above the last end are used)
                sd[FWP / 2 + fpad(i)] = v;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, in && dd != 0.f);
            if (bal) he = it * 32 + 31 - __clz(bal);
        }
        __syncwarp();
        if (lane == 0) gae_walk(sd, cnt2 - 1, he + 1, gl, gae);
        __syncwarp();
        for (int i = he + 1 + lane; i < cnt2; i += 32) returns[w0 + i] = __fadd_rn(sd[fpad(i)], sd[FWP / 2 + fpad(i)]);
        if (he >= 0) break;
        nv = values[w0];
        top = w0;
    }
}

// ---- flat form, persistent with prefetch ---------------------------------------------------------------------------
// Same algorithm, same shared-memory walk and the same bits as k_gae_flat; what changes is how the chunk arrives.
// k_gae_flat holds its 12 float4 loads in registers, so a CTA has loads in flight only during the first phase of its
// life (measured: ~45 % of the copy peak - too few bytes in flight per SM).  Here a CTA is persistent (3 per SM), the raw
// rewards / dones / values window of its NEXT chunk is fetched by 16-byte cp.async copies into a separate shared-memory
// staging area (no registers) as soon as the current chunk has been consumed from it, and the copies fly while the CTA
// compacts, walks and stores the current chunk: ~38 KB per CTA, ~115 KB per SM are in flight all the time.
// The staging area doubles as k_gae_flat's sd[] (edge chunks, long-segment windows): a prefetch is only issued when
// nobody uses it in that role any more.
constexpr int FRAW = FW + 4;   // + the value after the window (only word 0 of the 4 is used)
constexpr size_t GAE_PF_SMEM = (size_t)(3 * FRAW + 2 * FWP) * 4 + FL * 2 + 16;
static_assert((3 * FRAW) * 4 >= FWP * 4, "staging area too small to stand in for sd[]");

__global__ void __launch_bounds__(FT, 3)
k_gae_flat_pf(const float *__restrict__ rewards, const float *__restrict__ dones, const float *__restrict__ values,
              const float *__restrict__ next_value_ptr, int64_t N, int L, float g, float gl, float *__restrict__ returns) {
    extern __shared__ __align__(16) float dyn[];
    float *raw = dyn, *sd = dyn;                         // raw: [3][FRAW] rewards, dones, values of the window, unpadded
    float *sx = dyn + 3 * FRAW, *sv = sx + FWP;
    uint16_t *ends = reinterpret_cast<uint16_t *>(sv + FWP);
    __shared__ int bcnt[32], s_hend;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t nchunks = (N + L - 1) / L;
    const int W4 = (FH + L) / 4;   // float4 groups of a full window (L % 4 == 0)
    const float nv_end = next_value_ptr ? *next_value_ptr : values[N - 1];
    auto interior = [&](int64_t ch) { return ch > 0 && ch < nchunks && (ch + 1) * L <= N; };
    auto prefetch = [&](int64_t ch) {   // the whole window of an interior chunk; every thread commits one group
        if (interior(ch)) {
            const int64_t b = ch * L - FH;
#pragma unroll
            for (int k = 0; k < FNQ; ++k) {
                const int q = tid + k * FT;
                if (q < W4) {
                    cp_async16(raw + 4 * q, rewards + b + 4 * q);
                    cp_async16(raw + FRAW + 4 * q, dones + b + 4 * q);
                    cp_async16(raw + 2 * FRAW + 4 * q, values + b + 4 * q);
                }
            }
            if (tid == 0 && b + 4 * W4 < N) cp_async4(raw + 2 * FRAW + 4 * W4, values + b + 4 * W4);
        }
        cp_async_commit();
    };
    prefetch(blockIdx.x);
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t c0 = chunk * L, base = c0 - FH;   // window element i = global element base + i
        const int len = (int)min((int64_t)L, N - c0), lim = FH + len;
        const int h0 = base < 0 ? (int)-base : 0;                       // first valid window element
        const int ilast = N - 1 - base < (int64_t)lim ? (int)(N - 1 - base) : -1;   // window index of the very last transition
        const float nv_last = ilast >= 0 ? nv_end : 0.f;
        unsigned nib[FNQ];   // bit c of nib[k]: element 4 q + c (q = tid + k FT) is a segment end
#pragma unroll
        for (int k = 0; k < FNQ; ++k) nib[k] = 0;
        cp_async_wait<0>();
        __syncthreads();     // the window has landed; everybody is done with the previous chunk's sx / sv / ends
        if (interior(chunk)) {
            const float4 *gr = reinterpret_cast<const float4 *>(raw), *gd = reinterpret_cast<const float4 *>(raw + FRAW),
                         *gv = reinterpret_cast<const float4 *>(raw + 2 * FRAW);
            const int p0 = fpad(4 * tid);
#pragma unroll
            for (int k = 0; k < FNQ; ++k) {
                const int q = tid + k * FT;
                if (q < W4) {
                    const float4 a = gr[q], bq = gd[q], c = gv[q];
                    const int i = 4 * q, p = p0 + k * FQS;
                    const float vn = i + 3 == ilast ? nv_last : raw[2 * FRAW + i + 4];
                    const bool e0 = bq.x != 0.f, e1 = bq.y != 0.f, e2 = bq.z != 0.f, e3 = bq.w != 0.f || i + 3 == ilast;
                    sx[p] = gae_x(a.x, bq.x, c.x, c.y, e0, g, gl);
                    sx[p + 1] = gae_x(a.y, bq.y, c.y, c.z, e1, g, gl);
                    sx[p + 2] = gae_x(a.z, bq.z, c.z, c.w, e2, g, gl);
                    sx[p + 3] = gae_x(a.w, bq.w, c.w, vn, e3, g, gl);
                    sv[p] = c.x; sv[p + 1] = c.y; sv[p + 2] = c.z; sv[p + 3] = c.w;
                    nib[k] = (e0 ? 1u : 0u) | (e1 ? 2u : 0u) | (e2 ? 4u : 0u) | (e3 ? 8u : 0u);
                }
            }
        } else {
            // ---- first / last chunk: scalar staging (sd[] = the staging area; nothing is in flight into it)
            for (int i = h0 + tid; i < lim; i += FT) {
                const int p = fpad(i);
                sx[p] = rewards[base + i]; sd[p] = dones[base + i]; sv[p] = values[base + i];
            }
            __syncthreads();
            for (int i = h0 + tid; i < lim; i += FT) {
                const int p = fpad(i);
                const float d = sd[p];
                const float nv = i == ilast ? nv_last : (i + 1 < lim ? sv[fpad(i + 1)] : values[base + i + 1]);
                const bool is_end = d != 0.f || i == ilast;
                sx[p] = gae_x(sx[p], d, sv[p], nv, is_end, g, gl);
                sd[p] = is_end ? 1.f : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < FNQ; ++k) {
                const int i = 4 * (tid + k * FT);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    if (i + cc >= h0 && i + cc < lim && sd[fpad(i + cc)] != 0.f) nib[k] |= 1u << cc;
            }
        }
        // ---- compaction of the segment ends, ascending (as k_gae_flat)
        int pre[FNQ];
        {
            int hend = -1;
#pragma unroll
            for (int k = 0; k < FNQ; ++k) {
                const int mine = __popc(nib[k]);
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                pre[k] = incl - mine;
                if (lane == 31) bcnt[8 * k + w] = (k == 0 && w == 0) ? 0 : incl;
                if (k == 0 && w == 0) {   // the halo: its last end
                    const unsigned any = __ballot_sync(0xffffffffu, nib[0] != 0);
                    if (any) {
                        const int hl = 31 - __clz(any);
                        const unsigned hn = __shfl_sync(0xffffffffu, nib[0], hl);
                        hend = 4 * hl + 31 - __clz(hn);
                    }
                }
            }
            if (tid == 0) s_hend = hend;
        }
        __syncthreads();     // bcnt, s_hend visible; the staging area has been consumed by everybody
        int excl, ns;
        {
            const int v = bcnt[lane];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            excl = incl - v;
            ns = __shfl_sync(0xffffffffu, incl, 31);
        }
        const int hend = s_hend;
        const int s_first = hend >= 0 ? hend + 1 : h0;
        const bool open = ns > 0 && hend < 0 && base > 0;   // the first segment begins before the window (CTA-uniform)
        const int64_t next = chunk + gridDim.x;
        if (!open) prefetch(next);   // flies during the walk and the stores
        if (ns == 0) continue;       // no segment ends here: a later chunk owns everything in this one
#pragma unroll
        for (int k = 0; k < FNQ; ++k) {
            const int off = __shfl_sync(0xffffffffu, excl, 8 * k + w) + pre[k];
            if ((k == 0 && w == 0) || nib[k] == 0) continue;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
                if (nib[k] >> cc & 1) ends[off + __popc(nib[k] & ((1u << cc) - 1))] = (uint16_t)(4 * (tid + k * FT) + cc);
        }
        __syncthreads();
        float gae_open = 0.f;
        // ---- one segment per thread
        for (int k = tid; k < ns; k += FT) {
            const int e = ends[k], s = k ? ends[k - 1] + 1 : s_first;
            float gae = sx[fpad(e)];
#ifndef GAE_NOWALK
            gae_walk(sx, e - 1, s, gl, gae);
#endif
            if (k == 0) gae_open = gae;
        }
        __syncthreads();
        {   // returns of [s_first, last end]: coalesced
            const int last = ends[ns - 1];
            int i = s_first + tid;
            const float *px = sx + fpad(i), *pv = sv + fpad(i);
            float *go = returns + base + i;
            constexpr int ps = FT + FT / 32 + FT / 128;
            for (; i + 3 * FT <= last; i += 4 * FT, px += 4 * ps, pv += 4 * ps, go += 4 * FT) {
                const float a0 = __fadd_rn(px[0], pv[0]), a1 = __fadd_rn(px[ps], pv[ps]), a2 = __fadd_rn(px[2 * ps], pv[2 * ps]),
                            a3 = __fadd_rn(px[3 * ps], pv[3 * ps]);
                __stcs(go, a0); __stcs(go + FT, a1); __stcs(go + 2 * FT, a2); __stcs(go + 3 * FT, a3);
            }
            for (; i <= last; i += FT, px += ps, pv += ps, go += FT) __stcs(go, __fadd_rn(*px, *pv));
        }
        if (!open) continue;
        // ---- the rest of a first segment longer than the halo (as k_gae_flat: warp 0, windows in sd[]); the prefetch of
        // the next chunk waits until these windows are done with the staging area
        if (w == 0) {
            float gae = __shfl_sync(0xffffffffu, gae_open, 0);
            float nv = values[base + s_first];    // value of the element after the window's top
            for (int64_t top = base; top > 0;) {
                const int cnt2 = (int)min((int64_t)FX, top);
                const int64_t w0 = top - cnt2;
                int he = -1;
                __syncwarp();
                for (int it = 0; it * 32 < cnt2; ++it) {
                    const int i = it * 32 + lane;
                    const bool in = i < cnt2;
                    const float dd = in ? dones[w0 + i] : 0.f;
                    if (in) {
                        const float v = values[w0 + i];
                        const float vn = i + 1 < cnt2 ? values[w0 + i + 1] : nv;
                        sd[fpad(i)] = gae_x(rewards[w0 + i], 0.f, v, vn, false, g, gl);
                        sd[FWP / 2 + fpad(i)] = v;
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, in && dd != 0.f);
                    if (bal) he = it * 32 + 31 - __clz(bal);
                }
                __syncwarp();
                if (lane == 0) gae_walk(sd, cnt2 - 1, he + 1, gl, gae);
                __syncwarp();
                for (int i = he + 1 + lane; i < cnt2; i += 32) returns[w0 + i] = __fadd_rn(sd[fpad(i)], sd[FWP / 2 + fpad(i)]);
                if (he >= 0) break;
                nv = values[w0];
                top = w0;
            }
        }
        __syncthreads();
        prefetch(next);
    }
    cp_async_wait<0>();
}

// ---- advantage normalisation -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_adv_stats(const float *__restrict__ returns, const float *__restrict__ values, int64_t N, double *__restrict__ stats) {
    __shared__ double sc[32];
    double s1 = 0.0, s2 = 0.0;
    const int64_t n4 = N >> 2;
    const float4 *r4 = reinterpret_cast<const float4 *>(returns), *v4 = reinterpret_cast<const float4 *>(values);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 r = r4[i], v = v4[i];
        const float a0 = __fsub_rn(r.x, v.x), a1 = __fsub_rn(r.y, v.y), a2 = __fsub_rn(r.z, v.z), a3 = __fsub_rn(r.w, v.w);
        s1 += (double)a0 + (double)a1 + (double)a2 + (double)a3;
        s2 += (double)a0 * a0 + (double)a1 * a1 + (double)a2 * a2 + (double)a3 * a3;
    }
    if (blockIdx.x == 0 && threadIdx.x < (N & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        const float a = __fsub_rn(returns[i], values[i]);
        s1 += a;
        s2 += (double)a * a;
    }
    s1 = block_sum<double>(s1, sc);
    s2 = block_sum<double>(s2, sc);
    if (threadIdx.x == 0) {
        atomicAdd(stats + 0, s1);
        atomicAdd(stats + 1, s2);
        if (blockIdx.x == 0) atomicAdd(stats + 2, (double)N);
    }
}

__global__ void __launch_bounds__(256)
k_adv_apply(const float *__restrict__ returns, const float *__restrict__ values, int64_t N, const double *__restrict__ stats,
            float *__restrict__ adv) {
    const double cnt = stats[2];
    const double mean_d = stats[0] / cnt;
    const double var_d = fmax((stats[1] - stats[0] * mean_d) / (cnt - 1.0), 0.0);
    const float mean = (float)mean_d;
    const float denom = __fadd_rn((float)sqrt(var_d), 1e-8f);
    const int64_t n4 = N >> 2;
    const float4 *r4 = reinterpret_cast<const float4 *>(returns), *v4 = reinterpret_cast<const float4 *>(values);
    float4 *o4 = reinterpret_cast<float4 *>(adv);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 r = r4[i], v = v4[i];
        float4 o;
        o.x = __fdiv_rn(__fsub_rn(__fsub_rn(r.x, v.x), mean), denom);
        o.y = __fdiv_rn(__fsub_rn(__fsub_rn(r.y, v.y), mean), denom);
        o.z = __fdiv_rn(__fsub_rn(__fsub_rn(r.z, v.z), mean), denom);
        o.w = __fdiv_rn(__fsub_rn(__fsub_rn(r.w, v.w), mean), denom);
        o4[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < (N & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        adv[i] = __fdiv_rn(__fsub_rn(__fsub_rn(returns[i], values[i]), mean), denom);
    }
}

}  // namespace prl

using namespace prl;

extern "C" {

size_t prl_gae_ws_bytes(int64_t N) {
    (void)N;
    return 16;   // the chunked kernel needs no workspace; the parameter stays in the ABI
}

int prl_gae(const float *rewards, const float *dones, const float *values, const float *next_value_ptr, double gamma,
            double gae_lambda, int64_t N, float *returns, void *ws, size_t ws_bytes, void *stream) {
    (void)ws; (void)ws_bytes;
    PRL_REQUIRE(N >= 0, "prl_gae: negative N");
    if (N == 0) return PRL_OK;
    PRL_REQUIRE(rewards && dones && values && returns, "prl_gae: null pointer");
    // 46 KB of static shared memory per CTA: ask for the largest carve-out so that 3-4 of them fit per SM
    static const bool carve = (cudaFuncSetAttribute(k_gae_flat<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), true);
    (void)carve;
    const bool vec = (((uintptr_t)rewards | (uintptr_t)dones | (uintptr_t)values) & 15) == 0;
    const float g = (float)gamma, gl = (float)(gamma * gae_lambda);
    if (vec) {
        PRL_CUDA(cudaFuncSetAttribute(k_gae_flat_pf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GAE_PF_SMEM));   // per device
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        // persistent CTAs, 3 per SM.  Chunk length: FL once there is more than one round of chunks (measured: shorter
        // chunks that even out the rounds are slower - 47.7 us at L = 2816 against 42.5 us at L = 3072 for 8.4 M
        // transitions); a single partial round is spread over all CTAs in multiples of 128.
        const int64_t ctas = (int64_t)sms * 3;
        int L = FL;
        if (N < ctas * FL) L = (int)max((int64_t)512, min((int64_t)FL, ((N + ctas - 1) / ctas + 127) / 128 * 128));
        const int grid = (int)min((int64_t)cdiv(N, L), ctas);
        k_gae_flat_pf<<<grid, FT, GAE_PF_SMEM, (cudaStream_t)stream>>>(rewards, dones, values, next_value_ptr, N, L, g, gl, returns);
    }
    else k_gae_flat<false><<<cdiv(N, FL), FT, 0, (cudaStream_t)stream>>>(rewards, dones, values, next_value_ptr, N, g, gl, returns);
    return check_launch("k_gae_flat");
}

int prl_gae_columns(const float *rewards, const float *dones, const float *values, const int32_t *lengths, int E, int T_cap,
                    double gamma, double gae_lambda, float *returns, void *stream) {
    PRL_REQUIRE(E > 0 && T_cap > 0 && rewards && dones && values && returns, "prl_gae_columns: bad arguments");
    const bool ring = E % 4 == 0 && (((uintptr_t)rewards | (uintptr_t)dones | (uintptr_t)values) & 15) == 0;
    static const bool carve = (cudaFuncSetAttribute(k_gae_columns_ring, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), true);
    (void)carve;
    if (ring)
        k_gae_columns_ring<<<cdiv(E, 32), 32, 0, (cudaStream_t)stream>>>(rewards, dones, values, lengths, E, T_cap, (float)gamma,
                                                                       (float)(gamma * gae_lambda), returns);
    else
        k_gae_columns<<<cdiv(E, 256), 256, 0, (cudaStream_t)stream>>>(rewards, dones, values, lengths, E, T_cap, (float)gamma,
                                                                     (float)(gamma * gae_lambda), returns);
    return check_launch("k_gae_columns");
}

int prl_adv_normalize(const float *returns, const float *values, int64_t N, float *adv, double *stats, int phase, void *stream) {
    PRL_REQUIRE(N >= 0 && stats && (phase >= 1 && phase <= 3), "prl_adv_normalize: bad arguments");
    if (N == 0) return PRL_OK;
    PRL_REQUIRE(returns && values, "prl_adv_normalize: null pointer");
    PRL_REQUIRE(((uintptr_t)returns & 15) == 0 && ((uintptr_t)values & 15) == 0 && ((uintptr_t)adv & 15) == 0,
                "prl_adv_normalize: arrays must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = (int)min((int64_t)148 * 8, (N / 4 + 255) / 256 + 1);
    if (phase & 1) k_adv_stats<<<nb, 256, 0, st>>>(returns, values, N, stats);
    if (phase & 2) {
        PRL_REQUIRE(adv, "prl_adv_normalize: adv is NULL");
        k_adv_apply<<<nb, 256, 0, st>>>(returns, values, N, stats, adv);
    }
    return check_launch("k_adv");
}

}  // extern "C"
