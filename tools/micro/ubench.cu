// Micro-benchmarks behind the design of the lean update kernel (run on the GPU box: tools/micro/run.sh).
//  (1) issue economics of the packed fp32 instructions (FFMA2 / FADD2 / FMUL2) against scalar FFMA, alone and mixed
//      with MUFU / ALU work, at 4 and 16 warps per SM sub-partition
//  (2) column sums of a [32 rows][16 features] per-warp block: shuffle butterfly vs shared-memory transposition
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void k_issue(float *out, int iters) {
    float2 a[8], b = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    uint32_t u = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }        // 16 FFMA
            if (MODE == 1) a[i] = __ffma2_rn(a[i], b, c);                                               // 8 FFMA2
            if (MODE == 2) { a[i] = __ffma2_rn(a[i], b, c); u = (u << 3) ^ (u >> 5); u += 0x9e3779b9u; }  // 8 FFMA2 + ALU
            if (MODE == 3) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); u = (u << 3) ^ (u >> 5); u += 0x9e3779b9u; }
            if (MODE == 4) { a[i] = __ffma2_rn(a[i], b, c); float t; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(a[i].x)); a[i].y += t * 1e-9f; }
            if (MODE == 5) { a[i] = __fadd2_rn(a[i], c); a[i] = __fmul2_rn(a[i], b); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + u;
}

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

template <int MODE>
__global__ void k_colsum(float *out, int iters) {
    __shared__ __align__(16) float tbs[16][32 * 20];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *tb = tbs[warp];
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 1e-3f + i;
    float2 acc = make_float2(0.f, 0.f);
    float sacc0 = 0.f, sacc1 = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            sacc0 += colsum8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            sacc1 += colsum8(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
        } else {
            float4 *row = reinterpret_cast<float4 *>(tb + lane * 20);
#pragma unroll
            for (int c = 0; c < 4; ++c) row[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            __syncwarp();
            const int jp = lane & 7, rg = lane >> 3, rot = (rg & 1) << 2;
            const float *lo = tb + (rg * 8 + rot) * 20 + 2 * jp, *hi = tb + (rg * 8 + (rot ^ 4)) * 20 + 2 * jp;
            float2 s = acc;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = __fadd2_rn(s, *reinterpret_cast<const float2 *>(lo + k * 20));
#pragma unroll
            for (int k = 0; k < 4; ++k) s = __fadd2_rn(s, *reinterpret_cast<const float2 *>(hi + k * 20));
            acc = s;
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] * 1.0001f;   // 16 FMUL of "real work" between the sums
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + sacc0 + sacc1 + v[3];
}

template <typename K>
static float time_kernel(K kern, int grid, int block, float *out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, block>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<grid, block>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float *out; CK(cudaMalloc(&out, 148 * 1024 * 4 * 4));
    const int iters = 4096;
    const char *names[6] = {"16 FFMA", "8 FFMA2", "8 FFMA2 + 24 ALU", "16 FFMA + 24 ALU", "8 FFMA2 + 8 MUFU + 8 FMA", "8 FADD2 + 8 FMUL2"};
    for (int block : {512, 1024}) {
        float ms[6];
        ms[0] = time_kernel(k_issue<0>, 148, block, out, iters); ms[1] = time_kernel(k_issue<1>, 148, block, out, iters);
        ms[2] = time_kernel(k_issue<2>, 148, block, out, iters); ms[3] = time_kernel(k_issue<3>, 148, block, out, iters);
        ms[4] = time_kernel(k_issue<4>, 148, block, out, iters); ms[5] = time_kernel(k_issue<5>, 148, block, out, iters);
        for (int m = 0; m < 6; ++m) {
            // cycles per loop body per warp-scheduler at 1.9 GHz: ms * 1.9e6 / iters / (warps per scheduler)
            const double cyc = ms[m] * 1.9e6 / iters / (block / 32 / 4);
            printf("issue  block=%4d  %-26s %8.3f ms  ~%6.1f cycles per body per warp (1.9 GHz assumed)\n", block, names[m], ms[m], cyc);
        }
    }
    for (int block : {512}) {
        const float a = time_kernel(k_colsum<0>, 148, block, out, iters), b = time_kernel(k_colsum<1>, 148, block, out, iters);
        printf("colsum block=%4d  shuffle butterfly %8.3f ms   shared-memory transposition %8.3f ms  (per 16 values x 32 rows per warp, + 16 FMUL)\n", block, a, b);
    }
    CK(cudaDeviceSynchronize());
    printf("ok\n");
    return 0;
}
