// Parity-test hooks: device builds of the bit-exact math headers (sin / cos / pow, Philox, numpy's PCG64) exposed so that the
// tests can compare them with libm / numpy element by element.  Test infrastructure: built into libprl_b200_test.so, NOT
// into the product library (include/prl_b200_test.h).
#include <stdarg.h>

#include "../envs.cuh"
#include "../np_rng.cuh"

namespace prl {

static thread_local char g_test_err[512] = "";

// (the test library carries its own copy of the error plumbing that csrc/api.cu gives the product library)
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_test_err, sizeof g_test_err, fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return PRL_ERR_CUDA;
    }
    return PRL_OK;
}

__global__ void k_test_sincos(const double *__restrict__ x, double *__restrict__ s, double *__restrict__ c, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        s[i] = prl_trig::sin_glibc(x[i]);
        c[i] = prl_trig::cos_glibc(x[i]);
    }
}

__global__ void k_test_pow2(const double *__restrict__ x, double *__restrict__ out, const float *__restrict__ xf,
                            float *__restrict__ outf, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (x) out[i] = pow2_glibc(x[i]);
        if (xf) outf[i] = powf2_glibc(xf[i]);
    }
}

__global__ void k_test_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out) {
    Philox ph(seed);
    uint32_t r[4];
    ph(c0, c1, c2, c3, r);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
}

// out[i][k] = k-th 64-bit output of PCG64(SeedSequence(seeds[i]))
__global__ void k_test_pcg64(const uint64_t *__restrict__ seeds, int n, int draws, uint64_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pcg64 g = Pcg64::from_seed(seeds[i]);
    for (int k = 0; k < draws; ++k) out[(size_t)i * draws + k] = g.next64();
}

}  // namespace prl

using namespace prl;

extern "C" {

const char *prl_test_last_error(void) { return g_test_err; }

int prl_test_sincos(const double *x, double *s, double *c, int64_t n, void *stream) {
    if (n <= 0) return PRL_OK;
    k_test_sincos<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, s, c, n);
    return check_launch("k_test_sincos");
}
int prl_test_pow2(const double *x, double *out, const float *xf, float *outf, int64_t n, void *stream) {
    if (n <= 0) return PRL_OK;
    k_test_pow2<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, out, xf, outf, n);
    return check_launch("k_test_pow2");
}
int prl_test_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out4, void *stream) {
    k_test_philox<<<1, 1, 0, (cudaStream_t)stream>>>(seed, c0, c1, c2, c3, out4);
    return check_launch("k_test_philox");
}
int prl_test_pcg64(const uint64_t *seeds, int n, int draws, uint64_t *out, void *stream) {
    PRL_REQUIRE(seeds && out && n > 0 && draws > 0, "prl_test_pcg64: bad arguments");
    k_test_pcg64<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(seeds, n, draws, out);
    return check_launch("k_test_pcg64");
}

}  // extern "C"
