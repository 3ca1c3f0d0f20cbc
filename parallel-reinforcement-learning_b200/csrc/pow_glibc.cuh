// x**2 as numpy evaluates it for float64 / float32 SCALARS: np.float64.__pow__ calls libm pow(x, 2.0) and
// np.float32.__pow__ calls powf(x, 2.0f); neither is always the correctly rounded x*x (about 0.09% / 0.07% of
// arguments differ by one ulp).  gymnasium's Pendulum cost and Acrobot dynamics use `**2` on scalars.
// PLACEHOLDER: plain products until the libm-identical ports land (tracked in DESIGN.md, "known deviations").
#pragma once
namespace prl {
__device__ __forceinline__ double pow2_glibc(double x) { return __dmul_rn(x, x); }
__device__ __forceinline__ float powf2_glibc(float x) { return __fmul_rn(x, x); }
}  // namespace prl
