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

__global__ void __launch_bounds__(256)
k_gae_columns(const float *__restrict__ rewards, const float *__restrict__ dones, const float *__restrict__ values,
              const int32_t *__restrict__ lengths, int E, int T_cap, float g, float gl, float *__restrict__ returns) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int len = lengths ? min(lengths[e], T_cap) : T_cap;
    if (len <= 0) return;
    float gae = 0.f;
    float nv = values[(size_t)(len - 1) * E + e];
    // software prefetch one step ahead: the three loads of step t-1 are issued before step t's arithmetic retires
    size_t i = (size_t)(len - 1) * E + e;
    float r = rewards[i], d = dones[i], v = values[i];
    for (int t = len - 1; t >= 0; --t) {
        float r2 = 0.f, d2 = 0.f, v2 = 0.f;
        if (t > 0) {
            const size_t j = i - E;
            r2 = rewards[j]; d2 = dones[j]; v2 = values[j];
        }
        returns[i] = gae_step(r, d, v, nv, g, gl, gae);
        nv = v;
        r = r2; d = d2; v = v2;
        i -= E;
    }
}

// ---- flat form ---------------------------------------------------------------------------------------------------
constexpr int SEG_TPB = 1024;

__device__ __forceinline__ int is_seg_end(const float *__restrict__ dones, int64_t i, int64_t N) {
    return (i < N) && (dones[i] != 0.f || i == N - 1);
}

__global__ void k_seg_counts(const float *__restrict__ dones, int64_t N, int32_t *__restrict__ counts) {
    __shared__ int32_t sc[32];
    const int64_t i = (int64_t)blockIdx.x * SEG_TPB + threadIdx.x;
    const int c = block_sum<int32_t>(is_seg_end(dones, i, N), sc);
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void k_seg_ends(const float *__restrict__ dones, int64_t N, const int32_t *__restrict__ counts,
                           int64_t *__restrict__ ends, int64_t *__restrict__ nseg) {
    __shared__ int64_t sc64[32];
    __shared__ int32_t wsum[32];
    __shared__ int64_t base_s;
    int64_t part = 0;
    for (int j = threadIdx.x; j < (int)blockIdx.x; j += blockDim.x) part += counts[j];
    part = block_sum<int64_t>(part, sc64);
    if (threadIdx.x == 0) base_s = part;
    __syncthreads();
    const int64_t base = base_s;
    const int64_t i = (int64_t)blockIdx.x * SEG_TPB + threadIdx.x;
    const int f = is_seg_end(dones, i, N);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    if (w == 0) {
        int v = wsum[lane], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        wsum[lane] = incl - v;
        if (lane == 31 && blockIdx.x == gridDim.x - 1) *nseg = base + incl;
    }
    __syncthreads();
    if (f) ends[base + wsum[w] + __popc(bal & ((1u << lane) - 1))] = i;
}

__global__ void __launch_bounds__(128)
k_gae_segments(const float *__restrict__ rewards, const float *__restrict__ dones, const float *__restrict__ values,
               const float *__restrict__ next_value_ptr, const int64_t *__restrict__ ends, const int64_t *__restrict__ nseg,
               int64_t N, float g, float gl, float *__restrict__ returns) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *nseg) return;
    const int64_t end = ends[k];
    const int64_t start = k > 0 ? ends[k - 1] + 1 : 0;
    float nv = (end == N - 1) ? (next_value_ptr ? *next_value_ptr : values[N - 1]) : values[end + 1];
    float gae = 0.f;
    for (int64_t t = end; t >= start; --t) {
        const float v = values[t];
        returns[t] = gae_step(rewards[t], dones[t], v, nv, g, gl, gae);
        nv = v;
    }
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
    return (size_t)(cdiv(N, SEG_TPB) + 2) * sizeof(int32_t) + 16 + (size_t)(N + 2) * sizeof(int64_t);
}

int prl_gae(const float *rewards, const float *dones, const float *values, const float *next_value_ptr, double gamma,
            double gae_lambda, int64_t N, float *returns, void *ws, size_t ws_bytes, void *stream) {
    PRL_REQUIRE(N >= 0, "prl_gae: negative N");
    if (N == 0) return PRL_OK;
    PRL_REQUIRE(rewards && dones && values && returns && ws && ws_bytes >= prl_gae_ws_bytes(N), "prl_gae: null pointer or workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = cdiv(N, SEG_TPB);
    int32_t *counts = static_cast<int32_t *>(ws);
    char *p = static_cast<char *>(ws) + (((size_t)(nb + 2) * sizeof(int32_t) + 7) & ~(size_t)7);
    int64_t *nseg = reinterpret_cast<int64_t *>(p);
    int64_t *ends = nseg + 1;
    k_seg_counts<<<nb, SEG_TPB, 0, st>>>(dones, N, counts);
    k_seg_ends<<<nb, SEG_TPB, 0, st>>>(dones, N, counts, ends, nseg);
    // at most N segments; launch for the worst case, threads beyond *nseg exit
    k_gae_segments<<<cdiv(N, 128), 128, 0, st>>>(rewards, dones, values, next_value_ptr, ends, nseg, N, (float)gamma,
                                                (float)(gamma * gae_lambda), returns);
    return check_launch("k_gae_segments");
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
