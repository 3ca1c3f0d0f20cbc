// Which ingredient of the update kernel's launch makes its first global loads take ~4.5k cycles (tools/micro/ldlat.cu: ~1k)?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__global__ void k_write(float *p, int n, float v) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v + i; }
template <int ALLOC>
__global__ void __launch_bounds__(544, 1) k_read(const float *p, long long *out, float *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    long long t0 = clock64();
    if (ALLOC && warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    float s = 0.f;
    long long t1 = 0;
    if (warp < 16) {
        const float4 *src = reinterpret_cast<const float4 *>(p + 384) + tid * 4;
        float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
        s = a.x + b.y + c.z + d.w;
        if (s == 1234.5f) sink[0] = s;
        t1 = clock64();
        smem[tid * 16] = (unsigned char)s;
    }
    __syncthreads();
    long long t2 = clock64();
    if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    if (ALLOC && warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
    sink[4 + tid % 64] = s + smem[(tid * 7) % 1000];
}
template <typename K>
void run(const char *name, K kern, int grid, size_t smem, bool coop, float *p, long long *out, float *sink) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long h[2];
    for (int rep = 0; rep < 50; ++rep) {
        k_write<<<16, 256>>>(p, 9216, (float)rep);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(544); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = coop ? 1 : 0;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kern, (const float *)p, out, sink);
    }
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-60s grid %3d: W1-like row arrived after %6lld cycles, barrier passed after %6lld   (%s)\n", name, grid, h[0], h[1], cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float *p, *sink; long long *out;
    cudaMalloc(&p, 36 * 1024 + 4096); cudaMalloc(&sink, 4096); cudaMalloc(&out, 64);
    for (int grid : {1, 147}) {
        run("544 threads, 16 KB dynamic smem", k_read<0>, grid, 16 * 1024, false, p, out, sink);
        run("544 threads, 190 KB dynamic smem", k_read<0>, grid, 190 * 1024, false, p, out, sink);
        run("544 threads, 190 KB, tcgen05.alloc of 512 columns by warp 16", k_read<1>, grid, 190 * 1024, false, p, out, sink);
        run("544 threads, 190 KB, tcgen05.alloc, cooperative launch", k_read<1>, grid, 190 * 1024, true, p, out, sink);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
