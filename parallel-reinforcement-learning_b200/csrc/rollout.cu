// Rollout half of the hot path: vectorised env reset/step, active-mask bookkeeping (stream compaction), the
// device-resident VecMemory, PPO.get_action, and the fused one-launch AsyncPPO.worker.
//
// Reference: /root/reference/AsyncTools/AsyncPPO.py:11-146, AsyncTools/utils.py:1-50, PPO/PPO.py:82-96.
#include "envs.cuh"
#include "np_rng.cuh"
#include "policy.cuh"
#include "tiled_mlp.cuh"

namespace prl {

constexpr int TPB = 128;  // threads per block for the one-env-per-thread kernels

// ------------------------------------------------------------------------------------------------ state access
template <class ENV>
__device__ __forceinline__ void load_state(const double *__restrict__ state, int E, int e, double (&s)[ENV::S]) {
#pragma unroll
    for (int i = 0; i < ENV::S; ++i) s[i] = state[(size_t)i * E + e];
}
template <class ENV>
__device__ __forceinline__ void store_state(double *__restrict__ state, int E, int e, const double (&s)[ENV::S]) {
#pragma unroll
    for (int i = 0; i < ENV::S; ++i) state[(size_t)i * E + e] = s[i];
}
template <class ENV>
__device__ __forceinline__ void store_obs_row(float *__restrict__ obs, size_t row, const float (&o)[ENV::O]) {
    if constexpr (ENV::O == 4) {
        reinterpret_cast<float4 *>(obs)[row] = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int i = 0; i < ENV::O; ++i) obs[row * ENV::O + i] = o[i];
    }
}

// ------------------------------------------------------------------------------------------------ reset
// start state of env e in episode `episode`: numpy Generator.uniform(low, high) arithmetic on Philox(seed; e, episode) doubles
template <class ENV>
__device__ __forceinline__ void draw_reset(const Philox &ph, int e, uint64_t episode, double (&s)[ENV::S]) {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < ENV::S; ++i) {
        if (i >= ENV::RESET_DRAWS) { s[i] = 0.0; continue; }
        if ((i & 1) == 0) ph((uint32_t)e, (uint32_t)episode, (STREAM_RESET << 24) | (uint32_t)(i >> 1), (uint32_t)(episode >> 32), r);
        const double u = u01d(r[2 * (i & 1)], r[2 * (i & 1) + 1]);
        // numpy Generator.uniform: low + (high - low) * u
        const double hi = ENV::reset_hi(i), lo = ENV::reset_lo(i);
        double v = dadd(lo, dmul(dsub(hi, lo), u));
        if (ENV::RESET_F32) v = (double)(float)v;
        s[i] = v;
    }
}

template <class ENV>
__global__ void k_env_reset(int E, uint64_t seed, uint64_t episode, double *__restrict__ state,
                            int32_t *__restrict__ elapsed, uint8_t *__restrict__ terminal, float *__restrict__ obs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    Philox ph(seed);
    double s[ENV::S];
    draw_reset<ENV>(ph, e, episode, s);
    store_state<ENV>(state, E, e, s);
    elapsed[e] = 0;
    terminal[e] = 0;
    float o[ENV::O];
    ENV::obs(s, o);
    store_obs_row<ENV>(obs, e, o);
}

// reset() from numpy's seeded stream (np_rng.cuh): env e owns a PCG64 generator in rng[4][E]; with `seeds` it is first
// (re)created as PCG64(SeedSequence(seeds[e])) - gymnasium's env.reset(seed=...) - then RESET_DRAWS doubles are drawn
// as Generator.uniform(low, high) does and the advanced generator is stored back, so later resets continue the stream.
template <class ENV>
__global__ void k_env_reset_numpy(int E, const uint64_t *__restrict__ seeds, uint64_t *__restrict__ rng, double *__restrict__ state,
                                  int32_t *__restrict__ elapsed, uint8_t *__restrict__ terminal, float *__restrict__ obs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    Pcg64 g;
    if (seeds) g = Pcg64::from_seed(seeds[e]);
    else { g.shi = rng[e]; g.slo = rng[(size_t)E + e]; g.ihi = rng[2 * (size_t)E + e]; g.ilo = rng[3 * (size_t)E + e]; }
    double s[ENV::S];
#pragma unroll
    for (int i = 0; i < ENV::S; ++i) {
        if (i >= ENV::RESET_DRAWS) { s[i] = 0.0; continue; }
        const double hi = ENV::reset_hi(i), lo = ENV::reset_lo(i);
        double v = dadd(lo, dmul(dsub(hi, lo), g.next_double()));   // random_uniform: low + range * next_double
        if (ENV::RESET_F32) v = (double)(float)v;
        s[i] = v;
    }
    rng[e] = g.shi; rng[(size_t)E + e] = g.slo; rng[2 * (size_t)E + e] = g.ihi; rng[3 * (size_t)E + e] = g.ilo;
    store_state<ENV>(state, E, e, s);
    elapsed[e] = 0;
    terminal[e] = 0;
    float o[ENV::O];
    ENV::obs(s, o);
    store_obs_row<ENV>(obs, e, o);
}

template <class ENV>
__global__ void k_env_set_state(int E, const double *__restrict__ aos, double *__restrict__ state,
                                int32_t *__restrict__ elapsed, uint8_t *__restrict__ terminal, float *__restrict__ obs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    double s[ENV::S];
#pragma unroll
    for (int i = 0; i < ENV::S; ++i) s[i] = aos[(size_t)e * ENV::S + i];
    store_state<ENV>(state, E, e, s);
    elapsed[e] = 0;
    terminal[e] = 0;
    float o[ENV::O];
    ENV::obs(s, o);
    store_obs_row<ENV>(obs, e, o);
}

template <class ENV>
__global__ void k_env_get_state(int E, const double *__restrict__ state, double *__restrict__ aos) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
#pragma unroll
    for (int i = 0; i < ENV::S; ++i) aos[(size_t)e * ENV::S + i] = state[(size_t)i * E + e];
}

// ------------------------------------------------------------------------------------------------ compact step
template <class ENV, int ADT>
__global__ void k_env_step(int E, int n, const int32_t *__restrict__ active_idx, const void *__restrict__ actions,
                           double *__restrict__ state, int32_t *__restrict__ elapsed, int max_steps,
                           float *__restrict__ obs, double *__restrict__ rewards, uint8_t *__restrict__ dones,
                           uint8_t *__restrict__ truncs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int e = active_idx[i];
    double s[ENV::S];
    load_state<ENV>(state, E, e, s);
    typename ENV::Action a;
    if constexpr (ENV::CONT) a = static_cast<const float *>(actions)[(size_t)i * ENV::A];
    else if constexpr (ADT == PRL_ACT_I64) a = (int)static_cast<const int64_t *>(actions)[i];
    else a = static_cast<const int32_t *>(actions)[i];
    double r;
    const bool term = ENV::step(s, a, r);
    store_state<ENV>(state, E, e, s);
    const int el = elapsed[e] + 1;
    elapsed[e] = el;
    float o[ENV::O];
    ENV::obs(s, o);
    store_obs_row<ENV>(obs, i, o);
    rewards[i] = r;
    dones[i] = term;
    truncs[i] = el >= max_steps;
}

// ------------------------------------------------------------------------------------------------ stream compaction
// Two launches: per-block counts, then (block offset = sum of earlier counts) + warp-ballot ranks.  Output order is
// ascending index, as numpy boolean indexing gives (utils.py:4, :15).
constexpr int SCAN_TPB = 1024;

__global__ void k_flag_counts(const uint8_t *__restrict__ flags, int64_t n, int want, int32_t *__restrict__ counts) {
    __shared__ int32_t sc[32];
    const int64_t i = (int64_t)blockIdx.x * SCAN_TPB + threadIdx.x;
    const int f = (i < n) && ((flags[i] != 0) == (want != 0));
    const int c = block_sum<int32_t>(f, sc);
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__device__ __forceinline__ int64_t block_prefix_of_counts(const int32_t *__restrict__ counts, int b, int64_t *sc) {
    int64_t part = 0;
    for (int j = threadIdx.x; j < b; j += blockDim.x) part += counts[j];
    part = block_sum<int64_t>(part, sc);
    __shared__ int64_t base;
    if (threadIdx.x == 0) base = part;
    __syncthreads();
    return base;
}

__global__ void k_flag_compact(const uint8_t *__restrict__ flags, int64_t n, int want, const int32_t *__restrict__ counts,
                               int32_t *__restrict__ idx, int32_t *__restrict__ count_out) {
    __shared__ int64_t sc64[32];
    __shared__ int32_t wsum[32];
    const int64_t base = block_prefix_of_counts(counts, blockIdx.x, sc64);
    const int64_t i = (int64_t)blockIdx.x * SCAN_TPB + threadIdx.x;
    const int f = (i < n) && ((flags[i] != 0) == (want != 0));
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
        wsum[lane] = incl - v;  // exclusive
        if (lane == 31 && blockIdx.x == gridDim.x - 1) *count_out = (int32_t)(base + incl);
    }
    __syncthreads();
    if (f) idx[base + wsum[w] + __popc(bal & ((1u << lane) - 1))] = (int32_t)i;
}

__global__ void k_gather_rows(const float *__restrict__ rows, const int32_t *__restrict__ idx,
                              const int32_t *__restrict__ count, int width, float *__restrict__ out) {
    const int64_t total = (int64_t)(*count) * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / width;
        const int c = (int)(i - r * width);
        out[i] = rows[(int64_t)idx[r] * width + c];
    }
}

__global__ void k_mask_update(uint8_t *__restrict__ terminal, const int32_t *__restrict__ active_idx,
                              const uint8_t *__restrict__ dones, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) terminal[active_idx[i]] = dones[i] != 0;
}

// ------------------------------------------------------------------------------------------------ device VecMemory
__global__ void k_buffer_append(int E, int T_cap, int n, const int32_t *__restrict__ active_idx,
                                const float *__restrict__ states, int O, const float *__restrict__ actions, int AW,
                                const float *__restrict__ rewards, const float *__restrict__ dones,
                                float *__restrict__ bs, float *__restrict__ ba, float *__restrict__ br,
                                float *__restrict__ bd, int32_t *__restrict__ lengths, int32_t *__restrict__ overflow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int e = active_idx[i];
    const int t = lengths[e];
    if (t >= T_cap) { atomicExch(overflow, 1); return; }
    for (int c = 0; c < O; ++c) bs[((size_t)t * O + c) * E + e] = states[(size_t)i * O + c];
    for (int c = 0; c < AW; ++c) ba[((size_t)t * AW + c) * E + e] = actions[(size_t)i * AW + c];
    br[(size_t)t * E + e] = rewards[i];
    bd[(size_t)t * E + e] = dones[i];
    lengths[e] = t + 1;
}

// exclusive scan of int32 lengths -> int64 offsets (two launches, same structure as the compaction)
__global__ void k_len_block_sums(const int32_t *__restrict__ len, int E, int32_t *__restrict__ sums) {
    __shared__ int32_t sc[32];
    const int e = blockIdx.x * SCAN_TPB + threadIdx.x;
    const int c = block_sum<int32_t>(e < E ? len[e] : 0, sc);
    if (threadIdx.x == 0) sums[blockIdx.x] = c;
}

__global__ void k_len_offsets(const int32_t *__restrict__ len, int E, const int32_t *__restrict__ sums, int64_t base0,
                              int64_t *__restrict__ offsets, int64_t *__restrict__ total) {
    __shared__ int64_t sc64[32];
    __shared__ int32_t wsum[32];
    const int64_t base = base0 + block_prefix_of_counts(sums, blockIdx.x, sc64);
    const int e = blockIdx.x * SCAN_TPB + threadIdx.x;
    const int v = e < E ? len[e] : 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        int x = wsum[lane], in2 = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, in2, o);
            if (lane >= o) in2 += t;
        }
        wsum[lane] = in2 - x;
        if (lane == 31 && blockIdx.x == gridDim.x - 1) *total = base + in2;
    }
    __syncthreads();
    if (e < E) offsets[e] = base + wsum[w] + (incl - v);
}

// time-major [T][C][E] -> env-major flat [off[e] + t][C]: a warp transposes a 32(t) x 32(e) tile through shared
// memory so that both the reads (along e) and the writes (along t) are contiguous.
__global__ void k_transfer(int E, int T_cap, int C, const float *__restrict__ src, const int32_t *__restrict__ len,
                           const int64_t *__restrict__ offsets, float *__restrict__ dst, int64_t capacity) {
    __shared__ float tile[4][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e0 = (blockIdx.x * 4 + w) * 32;
    if (e0 >= E) return;
    const int e = e0 + lane;
    const int my_len = e < E ? len[e] : 0;
    int max_len = my_len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) max_len = max(max_len, __shfl_xor_sync(0xffffffffu, max_len, o));
    const int64_t my_off = e < E ? offsets[e] : 0;
    for (int c = 0; c < C; ++c) {
        for (int t0 = 0; t0 < max_len; t0 += 32) {
#pragma unroll 4
            for (int tt = 0; tt < 32; ++tt) {
                const int t = t0 + tt;
                tile[w][tt][lane] = (t < my_len) ? src[((size_t)t * C + c) * E + e] : 0.f;
            }
            __syncwarp();
            for (int ee = 0; ee < 32; ++ee) {
                const int l = __shfl_sync(0xffffffffu, my_len, ee);
                const int64_t off = __shfl_sync(0xffffffffu, my_off, ee);
                const int t = t0 + lane;
                if (t < l && off + t < capacity) dst[(off + t) * C + c] = tile[w][lane][ee];
            }
            __syncwarp();
        }
    }
}

// All fields in one launch.  The channels of the four fields (observation components, action components, reward, done)
// are packed into groups of <= 4; a warp moves one 32(t) x 32(e) tile of one group at a time: float4 loads (4 time rows x
// 128 bytes per warp instruction, 8 in flight per lane) fill [channel][t][e] shared-memory tiles, the warp then turns
// around (lane = t) and writes env-major rows - a whole float4 row per lane when the group is channels
// 4k..4k+3 of a field whose width is a multiple of 4 (CartPole observations), so the writes are 512 contiguous bytes.
struct XferGroup {
    const float *src[4];   // channel plane: element (t, e) at src[t * sstride + e]
    float *dst[4];         // element (n) at dst[n * dC]
    int64_t sstride[4];
    int dC[4];
    int n, vec;            // channels in the group; vec: one float4 store at dst[0] + n * 4 ... (dC == 4 k, 16-byte aligned)
};
constexpr int XF_MAX_GROUPS = 8;
struct XferPlan {
    int ngroups;
    XferGroup g[XF_MAX_GROUPS];
};
constexpr int XF_WARPS = 4;
__global__ void __launch_bounds__(XF_WARPS * 32)
k_transfer_tiles(int E, int T_cap, XferPlan plan, const int32_t *__restrict__ len, const int64_t *__restrict__ offsets, int64_t capacity) {
    extern __shared__ __align__(16) float xf_smem[];
    __shared__ int64_t s_off[XF_WARPS][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float(*tile)[32][33] = reinterpret_cast<float(*)[32][33]>(xf_smem + (size_t)w * 4 * 32 * 33);
    const int e = blockIdx.x * 32 + lane;
    const int t0 = (blockIdx.y * XF_WARPS + w) * 32;
    const int my_len = e < E ? min(len[e], T_cap) : 0;
    const int max_len = __reduce_max_sync(0xffffffffu, my_len);
    if (t0 >= max_len) return;
    const int64_t my_off = e < E ? offsets[e] : 0;
    const int t = t0 + lane;   // this lane's time step in the write phase
    // once per tile: destination row of (env ee, this lane's t) = s_off[ee] + lane; bit ee of `ok` = that row exists
    s_off[w][lane] = my_off + t0;
    unsigned ok = 0;
#pragma unroll
    for (int ee = 0; ee < 32; ++ee) {
        const int l = __shfl_sync(0xffffffffu, my_len, ee);
        const int64_t off = __shfl_sync(0xffffffffu, my_off, ee);
        ok |= (unsigned)(t < l && off + t < capacity) << ee;
    }
    for (int gi = 0; gi < plan.ngroups; ++gi) {
        const int gn = plan.g[gi].n;
        __syncwarp();          // the previous group's tile has been read
        // float4 loads: lane l fetches envs 4 (l % 8) .. + 3 of time rows l / 8 + 4 j; the scalar stores into the 33-word rows
        // are conflict-free (bank = row + column).  Rows past an env's length hold stale data that is never written out.
        {
            const int lrow = lane >> 3, lcol = (lane & 7) * 4;
            const int rows = min(32, min(max_len, T_cap) - t0);
            const bool col_in = blockIdx.x * 32 + lcol < E;
            for (int cc = 0; cc < gn; ++cc) {
                const int64_t ss = plan.g[gi].sstride[cc];
                const float *sp = plan.g[gi].src[cc] + blockIdx.x * 32 + lcol + (int64_t)(t0 + lrow) * ss;
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (col_in && 4 * j + lrow < rows) v[j] = __ldcs(reinterpret_cast<const float4 *>(sp + (int64_t)(4 * j) * ss));
                float *tp = &tile[cc][lrow][lcol];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (col_in && 4 * j + lrow < rows) { float *q = tp + 4 * j * 33; q[0] = v[j].x; q[1] = v[j].y; q[2] = v[j].z; q[3] = v[j].w; }
            }
        }
        __syncwarp();
        if (plan.g[gi].vec) {
            float4 *d4 = reinterpret_cast<float4 *>(plan.g[gi].dst[0]) + (int64_t)lane * (plan.g[gi].dC[0] >> 2);
            const int64_t rs = plan.g[gi].dC[0] >> 2;   // float4s per destination row
#pragma unroll
            for (int ee = 0; ee < 32; ++ee)
                if (ok >> ee & 1) d4[s_off[w][ee] * rs] = make_float4(tile[0][lane][ee], tile[1][lane][ee], tile[2][lane][ee], tile[3][lane][ee]);
        } else {
            for (int cc = 0; cc < gn; ++cc) {
                const int64_t dC = plan.g[gi].dC[cc];
                float *dp = plan.g[gi].dst[cc] + (int64_t)lane * dC;
#pragma unroll
                for (int ee = 0; ee < 32; ++ee)
                    if (ok >> ee & 1) dp[s_off[w][ee] * dC] = tile[cc][lane][ee];
            }
        }
    }
}

__global__ void k_zero_i32(int32_t *p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

// ------------------------------------------------------------------------------------------------ PPO.get_action
__global__ void __launch_bounds__(EV_THREADS, 2)
k_policy_act(const float *__restrict__ params, PolicyLayout L, float action_scaling, const float *__restrict__ states,
             const int32_t *__restrict__ row_ids, int64_t n, uint64_t seed, uint64_t call_index, void *__restrict__ actions,
             float *__restrict__ dist) {
    extern __shared__ __align__(16) float smem[];
    const EvSmem S = ev_layout(L, L.n_heads - 1);   // policy heads only (get_action does not touch the critic)
    ev_stage_weights(smem, S, params, L);
    const int tid = threadIdx.x, O = L.O;
    const int64_t row0 = (int64_t)blockIdx.x * EV_ROWS;
    const int rows = (int)min((int64_t)EV_ROWS, n - row0);
    float *sX = smem + S.x;
    for (int i = tid; i < EV_ROWS * O; i += EV_THREADS) sX[i] = i < rows * O ? __ldg(states + row0 * O + i) : 0.f;
    __syncthreads();
    ev_forward_tile(smem, S, L);
    __syncthreads();
    if (tid >= rows) return;
    const int64_t row = row0 + tid;
    float *o = smem + S.out + tid * S.so;
    Philox ph(seed);
    const uint32_t rid = row_ids ? (uint32_t)row_ids[row] : (uint32_t)row;
    uint32_t r[4];
    if (!L.cont) {
        ph(rid, (uint32_t)call_index, action_stream_word(0), (uint32_t)(call_index >> 32), r);
        const int a = ev_sample_categorical(o, L.A, u01f(r[0]), dist ? dist + row * L.A : nullptr);
        static_cast<int64_t *>(actions)[row] = a;
    } else {
        float *out = static_cast<float *>(actions) + row * L.A;
        float nrm[4];
        for (int a = 0; a < L.A; ++a) {
            if ((a & 3) == 0) {
                ph(rid, (uint32_t)call_index, action_stream_word(a >> 2), (uint32_t)(call_index >> 32), r);
                normals4(r, nrm);
            }
            const float mu = o[S.col[0] + a];
            const float ls = o[S.col[1] + a];
            const float sd = softplus_t(fminf(fmaxf(ls, -2.f), 2.f));
            const float tril = sqrtf(sd * sd);  // cholesky of diag(std^2)
            out[a] = tanhf(fmaf(tril, nrm[a & 3], mu)) * action_scaling;
            if (dist) { dist[row * 2 * L.A + a] = mu; dist[row * 2 * L.A + L.A + a] = tril; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ fused worker()
// One thread per env for the physics, sampling and bookkeeping; the policy forward of the CTA's 256 envs runs as one
// register-tiled pass per step (tiled_mlp.cuh) between two barriers.  TAPED: actions come from a tape (teacher forcing,
// parity tests and the env-step bandwidth benchmark), no network, 128 threads.
// RPT (tiled_mlp.cuh): rows per thread of the forward; the CTA holds 32 RPT envs.  8 for throughput; 1 when there are so few envs
// that the launch cannot fill the GPU anyway (configs[0]: 32 envs) - the 8 warps then share the CTA's 32 rows and a step is
// several times shorter.  Same bits either way.
template <class ENV, bool TAPED, int RPT = EV_RPT>
__global__ void __launch_bounds__(EV_THREADS, 2)
k_rollout(int E, int T_cap, const float *__restrict__ params, PolicyLayout L, float action_scaling, uint64_t seed,
          uint64_t episode, const void *__restrict__ tape, double *__restrict__ state, int32_t *__restrict__ elapsed,
          uint8_t *__restrict__ terminal, float *__restrict__ bs, float *__restrict__ ba, float *__restrict__ br,
          float *__restrict__ bd, int32_t *__restrict__ lengths, double *__restrict__ scores, float *__restrict__ blp,
          float *__restrict__ bv, int horizon, double *__restrict__ score_ws) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NT = TAPED ? TPB : 32 * RPT;   // envs per CTA
    EvSmem S{};
    // blp != nullptr: the old-policy evaluation of PPO.learn (PPO.py:134-154: log-prob of the stored action and V(s) under the
    // acting policy) is taken here, where the network outputs of the step already exist: the critic head joins the forward and two
    // more [T][E] planes are written.  Same forward, same epilogue arithmetic as k_policy_evaluate: the same bits.
    const bool with_eval = !TAPED && blp != nullptr;
    if constexpr (!TAPED) {
        S = ev_layout(L, with_eval ? L.n_heads : L.n_heads - 1);
        ev_stage_weights(smem, S, params, L);   // made visible by the first barrier of the step loop
    }
    __shared__ double red[32];
    const int e = blockIdx.x * NT + threadIdx.x;
    const bool valid = (int)threadIdx.x < NT && e < E;
    bool alive = valid;
    double s[ENV::S];
    if (valid) load_state<ENV>(state, E, e, s);
    double rsum = 0.0;
    int len = 0;
    // horizon > 0: opt-in AUTO-RESET (not the reference's worker, which lets finished envs drop out - AsyncPPO.py:118,143-146): an
    // env whose episode ends (terminated, or `horizon` = the TimeLimit steps into the episode) is reset in place and goes on, so
    // every env fills all T_cap slots; the slot at T_cap - 1 closes the last episode (done = 1) like a truncation.  The k-th
    // in-rollout reset of an env draws what reset() would draw in episode `episode | k << 40`.
    int ep_t = 0;
    uint32_t resets = 0;
    Philox ph(seed);
    for (int t = 0; t < T_cap; ++t) {
        float o[ENV::O];
        if (alive) ENV::obs(s, o);
        if constexpr (TAPED) {
            if (!__any_sync(0xffffffffu, alive)) break;
        } else {
            // the whole CTA runs the forward while any of its envs is alive (finished envs feed zeros, results unused)
            if ((int)threadIdx.x < NT) {
                float *sX = smem + S.x + threadIdx.x * ENV::O;
#pragma unroll
                for (int c = 0; c < ENV::O; ++c) sX[c] = alive ? o[c] : 0.f;
            }
            if (!__syncthreads_or(alive)) break;
            ev_forward_tile<RPT>(smem, S, L);
            __syncthreads();
        }
        if (alive) {
            typename ENV::Action a;
            float a_store;
            if constexpr (TAPED) {
                if constexpr (ENV::CONT) a = static_cast<const float *>(tape)[((size_t)t * E + e) * ENV::A];
                else a = static_cast<const int32_t *>(tape)[(size_t)t * E + e];
                a_store = (float)a;
            } else {
                float *out = smem + S.out + threadIdx.x * S.so;
                uint32_t r[4];
                ph((uint32_t)e, (uint32_t)t, action_stream_word(0), (uint32_t)episode, r);
                if constexpr (ENV::CONT) {
                    float nrm[4];
                    normals4(r, nrm);
                    const float mu = out[S.col[0]];
                    const float ls = out[S.col[1]];
                    const float sd = softplus_t(fminf(fmaxf(ls, -2.f), 2.f));
                    const float tril = sqrtf(sd * sd);
                    a = tanhf(fmaf(tril, nrm[0], mu)) * action_scaling;
                    a_store = a;
                    if (with_eval) {   // k_policy_evaluate's continuous epilogue at A = 1 (log-prob of the STORED action)
                        const float zt = (a_store - mu) / tril;
                        const float q = fmaf(zt, zt, 0.f), hld = 0.f + logf(tril);
                        blp[(size_t)t * E + e] = -0.5f * (1 * LOG_2PI + q) - hld;
                    }
                } else {
                    a = ev_sample_categorical(out, ENV::A, u01f(r[0]), nullptr);
                    a_store = (float)a;
                    if (with_eval) {
                        // ev_sample_categorical left e_a / S in the scratch row; k_policy_evaluate's epilogue from there on
                        float P = 0.f;
#pragma unroll
                        for (int k = 0; k < ENV::A; ++k) P += out[k];
                        const float pa = out[a] / P;
                        blp[(size_t)t * E + e] = logf(fminf(fmaxf(pa, F32_EPS), 1.0f - F32_EPS));
                    }
                }
                if (with_eval) bv[(size_t)t * E + e] = out[S.col[L.n_heads - 1]];
            }
            double r64;
            const bool term = ENV::step(s, a, r64);
            ++ep_t;
            const bool over = horizon > 0 && (term || ep_t >= horizon);   // auto-reset: this episode is over, the env is not
            const bool fin = horizon > 0 ? (over || t + 1 >= T_cap) : (term || t + 1 >= T_cap);
#pragma unroll
            for (int c = 0; c < ENV::O; ++c) bs[((size_t)t * ENV::O + c) * E + e] = o[c];
            ba[(size_t)t * E + e] = a_store;
            br[(size_t)t * E + e] = (float)r64;
            bd[(size_t)t * E + e] = fin ? 1.f : 0.f;
            rsum += r64;
            len = t + 1;
            alive = horizon > 0 ? (t + 1 < T_cap) : !fin;
            if (over && alive) {
                ++resets;
                draw_reset<ENV>(ph, e, episode | ((uint64_t)resets << 40), s);
                ep_t = 0;
            }
        }
    }
    if (valid) {
        store_state<ENV>(state, E, e, s);
        lengths[e] = len;
        elapsed[e] = len;
        terminal[e] = 1;
    }
    const double bsum = block_sum<double>(rsum, red);
    const double blen = block_sum<double>((double)len, red);
    if (threadIdx.x == 0) {
        if (score_ws == nullptr) {   // order-dependent in the last bits of the reward sum (the step count is exact either way)
            atomicAdd(scores + 0, bsum);
            atomicAdd(scores + 1, blen);
        } else {
            // bit-reproducible: every CTA leaves its partial sums, the one that draws the last ticket adds them in CTA order
            // score_ws = {ticket (8 bytes), partial[grid][2]}, zero ticket on entry and on exit
            score_ws[2 + 2 * blockIdx.x] = bsum;
            score_ws[3 + 2 * blockIdx.x] = blen;
            __threadfence();
            unsigned int *ticket = reinterpret_cast<unsigned int *>(score_ws);
            if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
                __threadfence();
                double a = 0.0, b = 0.0;
                for (unsigned int c = 0; c < gridDim.x; ++c) { a += __ldcg(score_ws + 2 + 2 * c); b += __ldcg(score_ws + 3 + 2 * c); }
                scores[0] += a;
                scores[1] += b;
                *ticket = 0u;
            }
        }
    }
}

}  // namespace prl

// =================================================================================================== C ABI
using namespace prl;

extern "C" {

int prl_env_reset(int env_id, int E, uint64_t seed, uint64_t episode, double *state, int32_t *elapsed,
                  uint8_t *terminal, float *obs, void *stream) {
    PRL_REQUIRE(E > 0 && state && elapsed && terminal && obs, "prl_env_reset: bad arguments (E=%d)", E);
    return dispatch_env(env_id, [&](auto env) -> int {
        using ENV = decltype(env);
        k_env_reset<ENV><<<cdiv(E, TPB), TPB, 0, (cudaStream_t)stream>>>(E, seed, episode, state, elapsed, terminal, obs);
        return check_launch("k_env_reset");
    });
}

int prl_env_reset_numpy(int env_id, int E, const uint64_t *seeds, uint64_t *rng, double *state, int32_t *elapsed,
                        uint8_t *terminal, float *obs, void *stream) {
    PRL_REQUIRE(E > 0 && rng && state && elapsed && terminal && obs, "prl_env_reset_numpy: bad arguments (E=%d)", E);
    return dispatch_env(env_id, [&](auto env) -> int {
        using ENV = decltype(env);
        k_env_reset_numpy<ENV><<<cdiv(E, TPB), TPB, 0, (cudaStream_t)stream>>>(E, seeds, rng, state, elapsed, terminal, obs);
        return check_launch("k_env_reset_numpy");
    });
}

int prl_env_set_state(int env_id, int E, const double *state_aos, double *state, int32_t *elapsed, uint8_t *terminal,
                      float *obs, void *stream) {
    PRL_REQUIRE(E > 0 && state_aos && state && elapsed && terminal && obs, "prl_env_set_state: bad arguments");
    return dispatch_env(env_id, [&](auto env) -> int {
        using ENV = decltype(env);
        k_env_set_state<ENV><<<cdiv(E, TPB), TPB, 0, (cudaStream_t)stream>>>(E, state_aos, state, elapsed, terminal, obs);
        return check_launch("k_env_set_state");
    });
}

int prl_env_get_state(int env_id, int E, const double *state, double *state_aos, void *stream) {
    PRL_REQUIRE(E > 0 && state && state_aos, "prl_env_get_state: bad arguments");
    return dispatch_env(env_id, [&](auto env) -> int {
        using ENV = decltype(env);
        k_env_get_state<ENV><<<cdiv(E, TPB), TPB, 0, (cudaStream_t)stream>>>(E, state, state_aos);
        return check_launch("k_env_get_state");
    });
}

int prl_env_step(int env_id, int E, int n, const int32_t *active_idx, const void *actions, int action_dtype,
                 double *state, int32_t *elapsed, int max_episode_steps, float *obs, double *rewards, uint8_t *dones,
                 uint8_t *truncs, void *stream) {
    PRL_REQUIRE(E > 0 && n >= 0 && n <= E, "prl_env_step: n=%d out of range for E=%d", n, E);
    if (n == 0) return PRL_OK;
    PRL_REQUIRE(active_idx && actions && state && elapsed && obs && rewards && dones && truncs, "prl_env_step: null pointer");
    return dispatch_env(env_id, [&](auto env) -> int {
        using ENV = decltype(env);
        cudaStream_t st = (cudaStream_t)stream;
        if (ENV::CONT) {
            PRL_REQUIRE(action_dtype == PRL_ACT_F32, "prl_env_step: continuous env needs float32 actions");
            k_env_step<ENV, PRL_ACT_F32><<<cdiv(n, TPB), TPB, 0, st>>>(E, n, active_idx, actions, state, elapsed, max_episode_steps, obs, rewards, dones, truncs);
        } else if (action_dtype == PRL_ACT_I64) {
            k_env_step<ENV, PRL_ACT_I64><<<cdiv(n, TPB), TPB, 0, st>>>(E, n, active_idx, actions, state, elapsed, max_episode_steps, obs, rewards, dones, truncs);
        } else {
            PRL_REQUIRE(action_dtype == PRL_ACT_I32, "prl_env_step: discrete env needs int32/int64 actions");
            k_env_step<ENV, PRL_ACT_I32><<<cdiv(n, TPB), TPB, 0, st>>>(E, n, active_idx, actions, state, elapsed, max_episode_steps, obs, rewards, dones, truncs);
        }
        return check_launch("k_env_step");
    });
}

size_t prl_scan_ws_bytes(int64_t n) { return (size_t)(cdiv(n, SCAN_TPB) + 1) * sizeof(int32_t) + (size_t)(n + 1) * sizeof(int64_t); }

int prl_compact_indices(const uint8_t *flags, int64_t n, int want, int32_t *idx, int32_t *count, void *ws, size_t ws_bytes,
                        void *stream) {
    PRL_REQUIRE(n >= 0 && count, "prl_compact_indices: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { PRL_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st)); return PRL_OK; }
    const int nb = cdiv(n, SCAN_TPB);
    PRL_REQUIRE(flags && idx && ws && ws_bytes >= (size_t)nb * sizeof(int32_t), "prl_compact_indices: workspace too small");
    int32_t *counts = static_cast<int32_t *>(ws);
    k_flag_counts<<<nb, SCAN_TPB, 0, st>>>(flags, n, want, counts);
    k_flag_compact<<<nb, SCAN_TPB, 0, st>>>(flags, n, want, counts, idx, count);
    return check_launch("k_flag_compact");
}

int prl_gather_rows(const float *rows, const int32_t *idx, const int32_t *count, int64_t max_rows, int width, float *out,
                    void *stream) {
    PRL_REQUIRE(width > 0 && max_rows >= 0 && count, "prl_gather_rows: bad arguments");
    if (max_rows == 0) return PRL_OK;
    const int64_t total = max_rows * width;
    const int nb = (int)min((int64_t)148 * 8, (total + 255) / 256);
    k_gather_rows<<<nb, 256, 0, (cudaStream_t)stream>>>(rows, idx, count, width, out);
    return check_launch("k_gather_rows");
}

int prl_mask_update(uint8_t *terminal, const int32_t *active_idx, const uint8_t *dones, int n, void *stream) {
    if (n <= 0) return PRL_OK;
    PRL_REQUIRE(terminal && active_idx && dones, "prl_mask_update: null pointer");
    k_mask_update<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(terminal, active_idx, dones, n);
    return check_launch("k_mask_update");
}

int prl_buffer_append(int E, int T_cap, int n, const int32_t *active_idx, const float *states, int obs_dim,
                      const float *actions, int act_width, const float *rewards, const float *dones, float *buf_states,
                      float *buf_actions, float *buf_rewards, float *buf_dones, int32_t *lengths, int32_t *overflow,
                      void *stream) {
    if (n <= 0) return PRL_OK;
    PRL_REQUIRE(E > 0 && T_cap > 0 && n <= E && obs_dim > 0 && act_width > 0, "prl_buffer_append: bad sizes");
    k_buffer_append<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(E, T_cap, n, active_idx, states, obs_dim, actions, act_width,
                                                                rewards, dones, buf_states, buf_actions, buf_rewards,
                                                                buf_dones, lengths, overflow);
    return check_launch("k_buffer_append");
}

int prl_buffer_transfer(int E, int T_cap, int obs_dim, int act_width, const float *buf_states, const float *buf_actions,
                        const float *buf_rewards, const float *buf_dones, int32_t *lengths, int64_t base, int64_t capacity,
                        float *mem_states, float *mem_actions, float *mem_rewards, float *mem_dones, int64_t *total,
                        void *ws, size_t ws_bytes, void *stream) {
    return prl_buffer_transfer_ex(E, T_cap, obs_dim, act_width, buf_states, buf_actions, buf_rewards, buf_dones, 0, nullptr, nullptr, lengths, base,
                                  capacity, mem_states, mem_actions, mem_rewards, mem_dones, total, ws, ws_bytes, stream);
}

int prl_buffer_transfer_ex(int E, int T_cap, int obs_dim, int act_width, const float *buf_states, const float *buf_actions,
                           const float *buf_rewards, const float *buf_dones, int n_extra, const float *const *extra_planes,
                           float *const *extra_rows, int32_t *lengths, int64_t base, int64_t capacity, float *mem_states,
                           float *mem_actions, float *mem_rewards, float *mem_dones, int64_t *total, void *ws, size_t ws_bytes,
                           void *stream) {
    PRL_REQUIRE(E > 0 && T_cap > 0 && total && ws && ws_bytes >= prl_scan_ws_bytes(E), "prl_buffer_transfer: bad arguments / workspace");
    PRL_REQUIRE(n_extra >= 0 && n_extra <= 4 && (n_extra == 0 || (extra_planes && extra_rows)), "prl_buffer_transfer_ex: 0..4 extra [T][E] planes");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = cdiv(E, SCAN_TPB);
    int32_t *sums = static_cast<int32_t *>(ws);
    int64_t *offsets = reinterpret_cast<int64_t *>(static_cast<char *>(ws) + (((size_t)(nb + 1) * sizeof(int32_t) + 7) & ~(size_t)7));
    k_len_block_sums<<<nb, SCAN_TPB, 0, st>>>(lengths, E, sums);
    k_len_offsets<<<nb, SCAN_TPB, 0, st>>>(lengths, E, sums, base, offsets, total);
    // pack the channels of the four fields into groups of <= 4 (a group of 4 aligned channels of one field is stored as float4)
    XferPlan plan{};
    struct Field { const float *src; float *dst; int C; } fields[8] = {
        {buf_states, mem_states, obs_dim}, {buf_actions, mem_actions, act_width}, {buf_rewards, mem_rewards, 1}, {buf_dones, mem_dones, 1}};
    const int n_fields = 4 + n_extra;
    for (int k = 0; k < n_extra; ++k) fields[4 + k] = Field{extra_planes[k], extra_rows[k], 1};
    bool fits = true;
    XferGroup cur{};
    auto flush = [&]() {
        if (cur.n == 0) return;
        if (plan.ngroups == XF_MAX_GROUPS) { fits = false; return; }
        plan.g[plan.ngroups++] = cur;
        cur = XferGroup{};
    };
    for (int fi = 0; fi < n_fields; ++fi) {
        const Field &f = fields[fi];
        PRL_REQUIRE(f.src && f.dst && f.C > 0, "prl_buffer_transfer: null field");
        int c = 0;
        if (f.C % 4 == 0 && ((uintptr_t)f.dst & 15) == 0) {
            for (; c + 4 <= f.C; c += 4) {
                flush();
                for (int k = 0; k < 4; ++k) { cur.src[k] = f.src + (size_t)(c + k) * E; cur.dst[k] = f.dst + c + k; cur.sstride[k] = (int64_t)f.C * E; cur.dC[k] = f.C; }
                cur.n = 4; cur.vec = 1;
                flush();
            }
        }
        for (; c < f.C; ++c) {
            if (cur.n == 4) flush();
            const int k = cur.n++;
            cur.src[k] = f.src + (size_t)c * E; cur.dst[k] = f.dst + c; cur.sstride[k] = (int64_t)f.C * E; cur.dC[k] = f.C;
        }
    }
    flush();
    for (int fi = 0; fi < n_fields; ++fi) fits = fits && ((uintptr_t)fields[fi].src & 15) == 0;
    if (fits && E % 4 == 0) {
        const size_t smem = (size_t)XF_WARPS * 4 * 32 * 33 * sizeof(float);
        PRL_CUDA(cudaFuncSetAttribute(k_transfer_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_transfer_tiles<<<dim3(cdiv(E, 32), cdiv(cdiv(T_cap, 32), XF_WARPS)), XF_WARPS * 32, smem, st>>>(E, T_cap, plan, lengths, offsets, capacity);
    } else {   // very wide observations / actions: one field at a time
        const int tb = cdiv(E, 128);
        k_transfer<<<tb, 128, 0, st>>>(E, T_cap, obs_dim, buf_states, lengths, offsets, mem_states, capacity);
        k_transfer<<<tb, 128, 0, st>>>(E, T_cap, act_width, buf_actions, lengths, offsets, mem_actions, capacity);
        k_transfer<<<tb, 128, 0, st>>>(E, T_cap, 1, buf_rewards, lengths, offsets, mem_rewards, capacity);
        k_transfer<<<tb, 128, 0, st>>>(E, T_cap, 1, buf_dones, lengths, offsets, mem_dones, capacity);
        for (int k = 0; k < n_extra; ++k) k_transfer<<<tb, 128, 0, st>>>(E, T_cap, 1, extra_planes[k], lengths, offsets, extra_rows[k], capacity);
    }
    k_zero_i32<<<cdiv(E, 256), 256, 0, st>>>(lengths, E);
    return check_launch("k_transfer");
}

int prl_policy_act(const float *params, int is_continuous, int obs_dim, int action_dim, float action_scaling,
                   const float *states, const int32_t *row_ids, int64_t n, uint64_t seed, uint64_t call_index,
                   void *actions, float *dist, void *stream) {
    PRL_REQUIRE(params && obs_dim > 0 && action_dim > 0 && n >= 0, "prl_policy_act: bad arguments");
    if (n == 0) return PRL_OK;
    PRL_REQUIRE(states && actions, "prl_policy_act: null pointer");
    const PolicyLayout L = make_policy_layout(is_continuous, obs_dim, action_dim);
    const size_t smem = (size_t)ev_layout(L, L.n_heads - 1).total * sizeof(float);
    PRL_REQUIRE(smem <= 227 * 1024, "prl_policy_act: observ_dim=%d action_dim=%d needs %zu B shared memory (> 227 KB)", obs_dim, action_dim, smem);
    PRL_CUDA(cudaFuncSetAttribute(k_policy_act, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_policy_act<<<cdiv(n, EV_ROWS), EV_THREADS, smem, (cudaStream_t)stream>>>(params, L, action_scaling, states, row_ids, n, seed, call_index, actions, dist);
    return check_launch("k_policy_act");
}

int prl_rollout(int env_id, int E, int T_cap, const float *params, float action_scaling, uint64_t seed, uint64_t episode,
                const void *tape, double *state, int32_t *elapsed, uint8_t *terminal, float *buf_states, float *buf_actions,
                float *buf_rewards, float *buf_dones, int32_t *lengths, double *scores, void *stream) {
    return prl_rollout_eval(env_id, E, T_cap, params, action_scaling, seed, episode, tape, state, elapsed, terminal, buf_states, buf_actions,
                            buf_rewards, buf_dones, nullptr, nullptr, lengths, scores, 0, nullptr, stream);
}

size_t prl_rollout_score_ws_doubles(int E) { return 2 + 2 * (size_t)cdiv(E > 0 ? E : 1, 32); }   // (a CTA holds at least 32 envs)

int prl_rollout_eval(int env_id, int E, int T_cap, const float *params, float action_scaling, uint64_t seed, uint64_t episode,
                     const void *tape, double *state, int32_t *elapsed, uint8_t *terminal, float *buf_states, float *buf_actions,
                     float *buf_rewards, float *buf_dones, float *buf_logp, float *buf_values, int32_t *lengths, double *scores,
                     int auto_reset_horizon, double *score_ws, void *stream) {
    PRL_REQUIRE(E > 0 && T_cap > 0 && state && elapsed && terminal && buf_states && buf_actions && buf_rewards && buf_dones &&
                    lengths && scores, "prl_rollout: bad arguments");
    PRL_REQUIRE(tape || params, "prl_rollout: need policy parameters or an action tape");
    PRL_REQUIRE(auto_reset_horizon >= 0 && episode < (1ull << 40), "prl_rollout_eval: bad auto_reset_horizon / episode");
    PRL_REQUIRE((buf_logp == nullptr) == (buf_values == nullptr) && !(tape && buf_logp),
                "prl_rollout_eval: buf_logp and buf_values go together and need the policy (no action tape)");
    return dispatch_env(env_id, [&](auto env) -> int {
        using ENV = decltype(env);
        cudaStream_t st = (cudaStream_t)stream;
        const PolicyLayout L = make_policy_layout(ENV::CONT, ENV::O, ENV::A);
        if (tape) {
            k_rollout<ENV, true><<<cdiv(E, TPB), TPB, 0, st>>>(E, T_cap, params, L, action_scaling, seed, episode, tape, state, elapsed,
                                                              terminal, buf_states, buf_actions, buf_rewards, buf_dones, lengths, scores,
                                                              nullptr, nullptr, auto_reset_horizon, score_ws);
        } else {
            PRL_REQUIRE(!(buf_logp && ENV::CONT && ENV::A != 1), "prl_rollout_eval: continuous envs with action_dim > 1 are evaluated by prl_policy_evaluate");
            const size_t smem = (size_t)ev_layout(L, buf_logp ? L.n_heads : L.n_heads - 1).total * sizeof(float);
            PRL_REQUIRE(smem <= 227 * 1024, "prl_rollout_eval: %zu B of shared memory needed (> 227 KB)", smem);
            // rows per thread of the forward = envs per CTA / 32 (see the kernel's header):
            //   1  few envs (at most 32 per CTA slot of the GPU): the step is latency-bound, all 8 warps share the CTA's 32 rows;
            //   7  when 224-env CTAs fill the CTA slots (2 per SM) more evenly than 256-env ones: 65 536 envs are 256 CTAs of 256 -
            //      108 SMs carry two of them, 40 carry one - or 293 CTAs of 224, two on every SM with 7/8 of the work each;
            //   8  otherwise.
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int slots = 2 * sms;   // (two resident CTAs per SM: the discrete layouts; the three-head continuous layout of
                                         //  121 KB fits once, where 7 and 8 rows tie at the BASELINE sizes - a model, not a measurement)
            auto rounds = [&](int rpt) { return cdiv(cdiv(E, 32 * rpt), slots) * rpt; };   // CTAs an SM slot works through x their size
            int rpt = E <= 32 * slots ? 1 : rounds(7) < rounds(8) ? 7 : 8;
            if (const char *ev = getenv("PRL_ROLLOUT_RPT")) rpt = atoi(ev) == 1 || atoi(ev) == 7 ? atoi(ev) : 8;   // (A/B between the forms)
            auto go = [&](auto kernel, int envs_per_cta) -> int {
                PRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kernel<<<cdiv(E, envs_per_cta), EV_THREADS, smem, st>>>(E, T_cap, params, L, action_scaling, seed, episode, nullptr, state,
                                                                       elapsed, terminal, buf_states, buf_actions, buf_rewards, buf_dones,
                                                                       lengths, scores, buf_logp, buf_values, auto_reset_horizon, score_ws);
                return PRL_OK;
            };
            const int rc = rpt == 1 ? go(k_rollout<ENV, false, 1>, 32) : rpt == 7 ? go(k_rollout<ENV, false, 7>, 224) : go(k_rollout<ENV, false, 8>, 256);
            if (rc != PRL_OK) return rc;
        }
        return check_launch("k_rollout");
    });
}

}  // extern "C"
