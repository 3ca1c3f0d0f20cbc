// Error plumbing and introspection of the C ABI (the parity-test hooks live in csrc/testhooks/, a separate library).
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

}  // extern "C"
