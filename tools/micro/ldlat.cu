// How long does the FIRST global load of a kernel take (the parameter staging of the update kernel waits ~5k cycles for it)?
// kernel A rewrites a 36 KB buffer (like the previous launch's AdamW), kernel B times loads from it.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k_write(float *p, int n, float v) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v + i; }
template <int MODE>
__global__ void k_read(const float *p, const float *other, long long *out, float *sink) {
    const int tid = threadIdx.x;
    long long t0 = clock64();
    float4 a;
    const float4 *src = reinterpret_cast<const float4 *>(p) + tid * 4;
    if (MODE == 0) a = __ldg(src);
    if (MODE == 1) a = __ldcg(src);
    if (MODE == 2) asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(src));
    float s = a.x + a.y + a.z + a.w;
    if (s == 1234.5f) sink[0] = s;
    long long t1 = clock64();
    float4 b = __ldg(reinterpret_cast<const float4 *>(p) + tid * 4 + 1);     // neighbouring 16 bytes (same line)
    s += b.x + b.y;
    if (s == 1234.5f) sink[1] = s;
    long long t2 = clock64();
    float4 c = __ldg(reinterpret_cast<const float4 *>(other) + tid * 4);    // another buffer, not written recently
    s += c.x + c.y;
    if (s == 1234.5f) sink[2] = s;
    long long t3 = clock64();
    float4 d = __ldg(reinterpret_cast<const float4 *>(other) + 1048576 + tid * 4);   // 16 MB further on: another page
    s += d.x + d.y;
    if (s == 1234.5f) sink[3] = s;
    long long t4 = clock64();
    if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; }
    sink[4 + (blockIdx.x * blockDim.x + tid) % 64] = s;
}
int main() {
    float *p, *other, *sink; long long *out, h[4];
    cudaMalloc(&p, 36 * 1024); cudaMalloc(&other, 64 << 20); cudaMalloc(&sink, 4096); cudaMalloc(&out, 64);
    cudaMemset(other, 0, 64 << 20);
    const char *names[3] = {"ld.global.nc", "ld.global.cg", "ld.relaxed.gpu"};
    for (int grid : {1, 148}) for (int mode = 0; mode < 3; ++mode) for (int rep = 0; rep < 3; ++rep) {
        k_write<<<16, 256>>>(p, 9216, (float)rep);
        if (mode == 0) k_read<0><<<grid, 512>>>(p, other, out, sink);
        if (mode == 1) k_read<1><<<grid, 512>>>(p, other, out, sink);
        if (mode == 2) k_read<2><<<grid, 512>>>(p, other, out, sink);
        cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
        if (rep == 2) printf("grid %3d %-15s first load %6lld cycles | same line again %5lld | other buffer %5lld | other page %5lld\n", grid, names[mode], h[0], h[1], h[2], h[3]);
    }
    // the same inside a 1000-launch stream of back-to-back launches (clocks up)
    for (int i = 0; i < 2000; ++i) { k_write<<<16, 256>>>(p, 9216, (float)i); k_read<0><<<148, 512>>>(p, other, out, sink); }
    cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    printf("after 2000 back-to-back pairs: first load %6lld cycles | same line again %5lld | other buffer %5lld | other page %5lld\n", h[0], h[1], h[2], h[3]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
