// Error plumbing, introspection and the parity-test hooks of the C ABI.
#include <stdarg.h>

#include "envs.cuh"
#include "mlp.cuh"

namespace prl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
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

}  // namespace prl

using namespace prl;

extern "C" {

const char *prl_last_error(void) { return g_err; }
int prl_version(void) { return PRL_VERSION; }

int prl_env_info(int env_id, int *S, int *O, int *A, int *cont, int *max_steps) {
    return dispatch_env(env_id, [&](auto env) -> int {
        using ENV = decltype(env);
        if (S) *S = ENV::S;
        if (O) *O = ENV::O;
        if (A) *A = ENV::A;
        if (cont) *cont = ENV::CONT;
        if (max_steps) *max_steps = ENV::MAX_STEPS;
        return PRL_OK;
    });
}

int64_t prl_policy_param_count(int is_continuous, int obs_dim, int action_dim) {
    return make_policy_layout(is_continuous, obs_dim, action_dim).total;
}
int64_t prl_rnd_param_count(int in_features, int out_features) { return make_rnd_layout(in_features, out_features).total; }

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

}  // extern "C"
