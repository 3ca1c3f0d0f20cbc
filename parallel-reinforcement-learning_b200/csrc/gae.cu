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
//   prl_gae          flat env-major [N] (the compute_gae signature): one segment per thread.
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

// ---- flat form ---------------------------------------------------------------------------------------------------
// One CTA per chunk of GC consecutive transitions.  The chunk's rewards / dones / values are loaded into shared memory
// with coalesced loads, the segment ends inside the chunk are compacted (warp ballots), each thread then walks one
// segment backwards IN SHARED MEMORY in the reference's operation order, and the returns leave with coalesced stores.
// A chunk owns exactly the segments that END in it: the first of them may begin in an earlier chunk - its owner follows
// it back through global memory (rare: one segment per chunk); elements after the chunk's last end belong to a later chunk.
constexpr int GC = 4096, GT = 256;
// shared-memory index of chunk element i: one pad word per 32 and per 128 elements, so that threads walking segments
// whose starts are 32, 64, 128, ... elements apart (equal-length episodes) hit different banks
__device__ __forceinline__ int gpad(int i) { return i + (i >> 5) + (i >> 7); }
constexpr int GCP = GC + GC / 32 + GC / 128;

__global__ void __launch_bounds__(GT)
k_gae_chunks(const float *__restrict__ rewards, const float *__restrict__ dones, const float *__restrict__ values,
             const float *__restrict__ next_value_ptr, int64_t N, float g, float gl, float *__restrict__ returns) {
    extern __shared__ __align__(16) float gsm[];
    float *sr = gsm, *sd = sr + GCP, *sv = sd + GCP;   // the return of element i overwrites sr[i] once r[i] has been used
    uint16_t *ends = reinterpret_cast<uint16_t *>(sv + GCP);
    __shared__ int wcount[GT / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * GC;
    const int len = (int)min((int64_t)GC, N - c0);
    const bool vec = (((uintptr_t)rewards | (uintptr_t)dones | (uintptr_t)values) & 15) == 0 && len == GC;
    if (vec) {
        const float4 *r4 = reinterpret_cast<const float4 *>(rewards + c0), *d4 = reinterpret_cast<const float4 *>(dones + c0),
                     *v4 = reinterpret_cast<const float4 *>(values + c0);
#pragma unroll 4
        for (int i = tid; i < GC / 4; i += GT) {
            const float4 a = r4[i], bq = d4[i], c = v4[i];
            const int p = gpad(4 * i);   // 4 consecutive elements never straddle a pad word (pads follow multiples of 32)
            sr[p] = a.x; sr[p + 1] = a.y; sr[p + 2] = a.z; sr[p + 3] = a.w;
            sd[p] = bq.x; sd[p + 1] = bq.y; sd[p + 2] = bq.z; sd[p + 3] = bq.w;
            sv[p] = c.x; sv[p + 1] = c.y; sv[p + 2] = c.z; sv[p + 3] = c.w;
        }
    } else {
        for (int i = tid; i < len; i += GT) {
            const int p = gpad(i);
            sr[p] = rewards[c0 + i];
            sd[p] = dones[c0 + i];
            sv[p] = values[c0 + i];
        }
    }
    __syncthreads();
    // ---- compact the segment ends of the chunk, ascending
    int nends = 0;
    for (int it = 0; it < GC / GT; ++it) {
        const int i = it * GT + tid;
        const bool f = i < len && (sd[gpad(i)] != 0.f || c0 + i == N - 1);
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wcount[w] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int j = 0; j < GT / 32; ++j) {
            const int c = wcount[j];
            before += j < w ? c : 0;
            total += c;
        }
        if (f) ends[nends + before + __popc(bal & ((1u << lane) - 1))] = (uint16_t)i;
        nends += total;
        __syncthreads();
    }
    // ---- one segment per thread
    for (int k = tid; k < nends; k += GT) {
        const int e = ends[k], s = k ? ends[k - 1] + 1 : 0;
        const int64_t ge = c0 + e;
        float nv = (ge == N - 1) ? (next_value_ptr ? *next_value_ptr : values[N - 1]) : (e + 1 < len ? sv[gpad(e + 1)] : values[ge + 1]);
        float gae = 0.f;
        // 8 steps of shared-memory loads and of the gae-independent part (delta) are in flight per iteration; only
        // gae = delta + (gl * nd) * gae is a serial chain
        int t = e;
        for (; t - 7 >= s; t -= 8) {
            float r8[8], d8[8], v8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int p = gpad(t - u);
                r8[u] = sr[p]; d8[u] = sd[p]; v8[u] = sv[p];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                sr[gpad(t - u)] = gae_step(r8[u], d8[u], v8[u], nv, g, gl, gae);
                nv = v8[u];
            }
        }
        for (; t >= s; --t) {
            const int p = gpad(t);
            const float v = sv[p];
            sr[p] = gae_step(sr[p], sd[p], v, nv, g, gl, gae);
            nv = v;
        }
        if (k == 0) {   // the part of this segment that lies in earlier chunks
            for (int64_t t = c0 - 1; t >= 0 && dones[t] == 0.f; --t) {
                const float v = values[t];
                returns[t] = gae_step(rewards[t], 0.f, v, nv, g, gl, gae);
                nv = v;
            }
        }
    }
    __syncthreads();
    const int last = nends ? ends[nends - 1] : -1;
    for (int i = tid; i <= last; i += GT) returns[c0 + i] = sr[gpad(i)];
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
    const size_t smem = (size_t)3 * GCP * sizeof(float) + (size_t)GC * sizeof(uint16_t);
    PRL_CUDA(cudaFuncSetAttribute(k_gae_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gae_chunks<<<cdiv(N, GC), GT, smem, (cudaStream_t)stream>>>(rewards, dones, values, next_value_ptr, N, (float)gamma,
                                                               (float)(gamma * gae_lambda), returns);
    return check_launch("k_gae_chunks");
}

int prl_gae_columns(const float *rewards, const float *dones, const float *values, const int32_t *lengths, int E, int T_cap,
                    double gamma, double gae_lambda, float *returns, void *stream) {
    PRL_REQUIRE(E > 0 && T_cap > 0 && rewards && dones && values && returns, "prl_gae_columns: bad arguments");
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
