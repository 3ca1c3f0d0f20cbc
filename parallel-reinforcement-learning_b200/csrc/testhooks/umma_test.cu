// Parity-test hook for the tensor-core building blocks (csrc/umma.cuh): one CTA stages fp32 matrices in the
// row-per-thread shared-memory layout, issues tcgen05.mma kind::f16 (bf16) in each of the operand-major combinations the
// fused update kernel uses, and returns the accumulator.  Test infrastructure behind prl_test_umma.
#include "../common.cuh"
#include "../umma.cuh"

namespace prl {
using namespace umma;

// stage Y[rows][cols] (row-major fp32, global) as bf16 into the row-per-thread layout; rows == 128 == blockDim.x
__device__ __forceinline__ void stage_rows(unsigned char *dst, const float *__restrict__ src, int cols) {
    const int r = threadIdx.x;
    for (int c = 0; c < cols; c += 8) {
        const float4 a = *reinterpret_cast<const float4 *>(src + (size_t)r * cols + c);
        const float4 b = *reinterpret_cast<const float4 *>(src + (size_t)r * cols + c + 4);
        *reinterpret_cast<uint4 *>(dst + (c / 8) * CHUNK + r * 16) =
            make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
}

// stage as bf16 into 128-byte rows with the 128B swizzle: byte(r, c) = r * 128 + (((c / 8) ^ (r % 8)) * 16) + (c % 8) * 2 within
// each [128][64] block (blocks of 64 columns are 16 KB apart); the base must be 1024-byte aligned
__device__ __forceinline__ void stage_rows_sw128(unsigned char *dst, const float *__restrict__ src, int cols) {
    const int r = threadIdx.x;
    for (int c = 0; c < cols; c += 8) {
        const float4 a = *reinterpret_cast<const float4 *>(src + (size_t)r * cols + c);
        const float4 b = *reinterpret_cast<const float4 *>(src + (size_t)r * cols + c + 4);
        const int blk = c / 64, ch = (c / 8) & 7;
        *reinterpret_cast<uint4 *>(dst + blk * 16384 + r * 128 + ((ch ^ (r & 7)) << 4)) =
            make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
}

// Generic driver: A is staged as [128][a_cols], B as [128][b_cols] (layout 0 = interleaved / no swizzle, 2 = 128-byte swizzled
// rows); K-step s uses descriptors at off + s * step; the whole sequence is issued `reps` times (accumulating) and timed.
struct UmmaTestCfg {
    int a_cols, b_cols, n_out;
    uint32_t idesc;
    int nsteps, a_off, a_step, a_lbo, a_sbo, b_off, b_step, b_lbo, b_sbo;
    int a_layout, b_layout, reps;   // reps < 0: |reps| repetitions of the weight-stationary form (tcgen05.mma.ws)
    int d_lane;                     // lane offset of the accumulator address (0 for the product's layouts)
};

__global__ void __launch_bounds__(128, 1)
k_test_umma(UmmaTestCfg c, const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, int *__restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sA = smem_raw;               // up to 128 x 128 bf16 = 32 KB
    unsigned char *sB = smem_raw + 32 * 1024;   // up to 128 x 64 bf16 = 16 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5;

    if (c.a_layout == 2) stage_rows_sw128(sA, A, c.a_cols); else stage_rows(sA, A, c.a_cols);
    if (c.b_layout == 2) stage_rows_sw128(sB, B, c.b_cols); else stage_rows(sB, B, c.b_cols);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, 128);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_slot;

    long long t0 = 0;
    if (warp == 0) {   // warp-uniform control flow around the single issuing lane: descriptors stay in uniform registers
        const uint32_t a0 = smem_u32(sA) + c.a_off, b0 = smem_u32(sB) + c.b_off;
        uint64_t da[8], db[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            da[s] = smem_desc_sw(a0 + s * c.a_step, c.a_lbo, c.a_sbo, c.a_layout);
            db[s] = smem_desc_sw(b0 + s * c.b_step, c.b_lbo, c.b_sbo, c.b_layout);
        }
        const uint32_t idesc = c.idesc;
        const int nsteps = c.nsteps;
        const bool ws = c.reps < 0;
        const int reps = ws ? -c.reps : c.reps;
        const uint32_t dst = tmem + ((uint32_t)c.d_lane << 16);
        t0 = clock64();
        if (elect_one()) {
            for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    if (s < nsteps) {
                        if (ws) mma_bf16_ss_ws(dst, da[s], db[s], idesc, (rep | s) != 0);
                        else mma_bf16_ss(dst, da[s], db[s], idesc, (rep | s) != 0);
                    }
            }
            mma_commit(&bar);
        }
        __syncwarp();
    }
    const bool ok = mbar_wait(&bar, 0);
    if (threadIdx.x == 0) {
        status[1] = (int)(clock64() - t0);
        if (!ok) status[0] = 1;
    }
    fence_after_sync();
    for (int col = 0; col < c.n_out; col += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + col, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) D[(size_t)threadIdx.x * c.n_out + col + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace prl

using namespace prl;

extern "C" int prl_test_umma(int mode, const float *A, const float *B, float *D, int *status, const int32_t *cfg_host, void *stream) {
    PRL_REQUIRE(mode >= -1 && mode <= 3 && A && B && D && status, "prl_test_umma: bad arguments");
    UmmaTestCfg c;
    constexpr int CH = umma::CHUNK;
    switch (mode) {
        // forward, two heads stacked along N: A, B K-major, 4 K-steps of 16 features
        case 0: c = {64, 64, 128, idesc_bf16(128, 128, 0, 0), 4, 0, 2 * CH, CH, 128, 0, 2 * CH, CH, 128, 0, 0, 1, 0}; break;
        // dgrad of head 1: A K-major; B'[n' = k][K' = j] = W[64 + j][k] is the MN-major view of the same weight buffer
        case 1: c = {64, 64, 64, idesc_bf16(128, 64, 0, 1), 4, 0, 2 * CH, CH, 128, 64 * 16, 256, 128, CH, 0, 0, 1, 0}; break;
        // weight gradient: contraction over the 128 rows, both operands MN-major, 8 K-steps of 16 rows
        case 2: c = {128, 64, 64, idesc_bf16(128, 64, 1, 1), 8, 0, 256, 128, CH, 0, 256, 128, CH, 0, 0, 1, 0}; break;
        case 3: c = {128, 16, 16, idesc_bf16(128, 16, 1, 1), 8, 0, 256, 128, CH, 0, 256, 128, CH, 0, 0, 1, 0}; break;
        default:
            PRL_REQUIRE(cfg_host, "prl_test_umma: mode -1 needs cfg_host[17]");
            c = {cfg_host[0], cfg_host[1], cfg_host[2], (uint32_t)cfg_host[3], cfg_host[4], cfg_host[5], cfg_host[6], cfg_host[7],
                 cfg_host[8], cfg_host[9], cfg_host[10], cfg_host[11], cfg_host[12], cfg_host[13], cfg_host[14], cfg_host[15], cfg_host[16]};
    }
    PRL_REQUIRE(c.a_cols % 8 == 0 && c.a_cols <= 128 && c.b_cols % 8 == 0 && c.b_cols <= 64 && c.n_out % 16 == 0 && c.n_out <= 128 &&
                    c.nsteps >= 1 && c.nsteps <= 8 && c.reps != 0 && c.reps >= -4096 && c.reps <= 4096 && (c.a_layout | c.b_layout | 2) == 2 &&
                    c.d_lane >= 0 && c.d_lane < 128,
                "prl_test_umma: configuration out of range");
    const size_t smem = 48 * 1024 + 1024;
    PRL_CUDA(cudaFuncSetAttribute(k_test_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_test_umma<<<1, 128, smem, (cudaStream_t)stream>>>(c, A, B, D, status);
    return check_launch("k_test_umma");
}
