// x**2 with the SAME BITS numpy produces for float64 / float32 SCALARS.
//
// Why: np.float64.__pow__ calls libm pow(x, 2.0) and np.float32.__pow__ calls powf(x, 2.0f); neither is the
// correctly rounded x*x (about 0.09% / 0.07% of arguments differ by one ulp).  gymnasium's Pendulum cost
// (`angle_normalize(th) ** 2 + 0.1 * thdot**2 + 0.001 * (u**2)`) and Acrobot dynamics (`dtheta2**2`, `d2**2`, ...)
// use `**2` on scalars, and Acrobot is chaotic, so a one-ulp difference in the fp64 state becomes a visible
// difference in the float32 observations a few hundred steps later.
//
// This is a restatement of the published algorithm glibc >= 2.28 uses (Arm Optimized Routines pow/powf:
// sysdeps/ieee754/dbl-64/e_pow.c log_inline + exp_inline, flt-32/e_powf.c log2_inline + exp2_inline) specialised
// to y == 2, with every rounding spelled out: each PRL_FMA below is an FMA in libm's x86-64 `-mfma` multiarch
// variant (__pow_fma/__powf_fma, selected by ifunc on any CPU with FMA+AVX2), every other operation rounds on its
// own.  The operation sequence was established by reading that variant's machine code; tables: tools/gen_powtab.py.
//
// Attribution: the algorithm and its constants are those of the GNU C Library's pow / powf, which were contributed from the Arm
// Optimized Routines (Copyright (C) Arm Limited; MIT OR Apache-2.0 WITH LLVM-exception upstream, LGPL-2.1-or-later in glibc).  No
// upstream source text is included here: this file is an independent restatement of that algorithm.
//
// Outside the algorithm's normal-result range (x zero / subnormal / inf / nan, x*x overflowing or subnormal) the
// result is the plain product x*x; classic-control states never get there.
//
// The same header compiles on the host with -DPRL_TRIG_HOST (tests/host/pow_check.cpp) so the exact source can be
// checked against libm on the CPU; the product only ever uses the device build.
#pragma once
#include <stdint.h>

#include "powtab.inc"
#include "trig_glibc.cuh"  // PRL_FMA / PRL_MUL / PRL_ADD / PRL_SUB and the host/device switch

#ifdef PRL_TRIG_HOST
#define PRL_POW_FN static inline
static const double prl_pow_log_tab[128 * 3] = {PRL_POW_LOG_TAB};
static const uint64_t prl_pow_exp_tab[256] = {PRL_POW_EXP_TAB};
static const double prl_powf_log2_tab[32] = {PRL_POWF_LOG2_TAB};
static const uint64_t prl_powf_exp2_tab[32] = {PRL_POWF_EXP2_TAB};
#define PRL_PTAB(t, i) t[i]
static inline uint64_t prl_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double prl_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
static inline uint32_t prl_f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static inline float prl_u2f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
#define PRL_FMULF(a, b) ((a) * (b))
#else
#define PRL_POW_FN __device__ __forceinline__
__device__ const double prl_pow_log_tab[128 * 3] = {PRL_POW_LOG_TAB};
__device__ const uint64_t prl_pow_exp_tab[256] = {PRL_POW_EXP_TAB};
__device__ const double prl_powf_log2_tab[32] = {PRL_POWF_LOG2_TAB};
__device__ const uint64_t prl_powf_exp2_tab[32] = {PRL_POWF_EXP2_TAB};
#define PRL_PTAB(t, i) __ldg(&t[i])
__device__ __forceinline__ uint64_t prl_d2u(double x) { return (uint64_t)__double_as_longlong(x); }
__device__ __forceinline__ double prl_u2d(uint64_t u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ uint32_t prl_f2u(float x) { return __float_as_uint(x); }
__device__ __forceinline__ float prl_u2f(uint32_t u) { return __uint_as_float(u); }
#define PRL_FMULF(a, b) __fmul_rn((a), (b))
#endif

namespace prl {

// libm pow(x, 2.0)
PRL_POW_FN double pow2_glibc(double x) {
    constexpr double LH[9] = {PRL_POW_LOG_HDR};  // ln2hi, ln2lo, A[0..6]
    constexpr double EH[8] = {PRL_POW_EXP_HDR};  // InvLn2N, Shift, NegLn2hiN, NegLn2loN, C2..C5
    const uint64_t ix = prl_d2u(x) & 0x7fffffffffffffffull;  // y = 2 is an even integer: sign_bias = 0, |x| is used
    const uint32_t topx = (uint32_t)(ix >> 52);
    if (topx - 1u >= 0x7feu) return PRL_MUL(x, x);
    // ---- log_inline: log(|x|) = hi + tail
    const uint64_t tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((int64_t)tmp >> 52);
    const double z = prl_u2d(ix - (tmp & 0xfff0000000000000ull));
    const double kd = (double)k;
    const double invc = PRL_PTAB(prl_pow_log_tab, 3 * i), logc = PRL_PTAB(prl_pow_log_tab, 3 * i + 1),
                 logctail = PRL_PTAB(prl_pow_log_tab, 3 * i + 2);
    const double t1 = PRL_FMA(kd, LH[0], logc);
    const double lo1 = PRL_FMA(kd, LH[1], logctail);
    const double r = PRL_FMA(z, invc, -1.0);
    const double ar = PRL_MUL(r, LH[2]);
    const double q1 = PRL_FMA(r, LH[4], LH[3]);
    const double q2 = PRL_FMA(r, LH[6], LH[5]);
    const double t2 = PRL_ADD(r, t1);
    const double lo2 = PRL_ADD(PRL_SUB(t1, t2), r);
    const double ar2 = PRL_MUL(r, ar);
    const double ar3 = PRL_MUL(r, ar2);
    const double lo3 = PRL_FMA(ar, r, -ar2);
    const double hi = PRL_ADD(t2, ar2);
    const double q3 = PRL_FMA(r, LH[8], LH[7]);
    const double q23 = PRL_FMA(q3, ar2, q2);
    const double lo4 = PRL_ADD(PRL_SUB(t2, hi), ar2);
    const double poly = PRL_FMA(ar2, q23, q1);
    const double lo = PRL_FMA(ar3, poly, PRL_ADD(PRL_ADD(PRL_ADD(lo1, lo2), lo3), lo4));
    const double lhi = PRL_ADD(hi, lo);
    const double ltail = PRL_ADD(PRL_SUB(hi, lhi), lo);
    // ---- y * log(x) as ehi + elo (y = 2)
    const double ehi = PRL_MUL(2.0, lhi);
    const double elo = PRL_FMA(2.0, ltail, PRL_FMA(lhi, 2.0, -ehi));
    // ---- exp_inline
    const uint32_t abstop = (uint32_t)(prl_d2u(ehi) >> 52) & 0x7ff;
    bool special = false;  // 512 <= |2 ln x| < 1024: the scale 2^(k/N) alone would over/underflow (libm's specialcase())
    if (abstop - 0x3c9u > 0x3eu) {
        if (abstop < 0x3c9u) return PRL_ADD(1.0, ehi);  // |2 ln x| < 2^-54
        if (abstop > 0x408u) return PRL_MUL(x, x);       // overflows / underflows to 0 either way
        special = true;
    }
    const double ks = PRL_FMA(ehi, EH[0], EH[1]);
    const uint64_t ki = prl_d2u(ks);
    const double kd2 = PRL_SUB(ks, EH[1]);
    const double r0 = PRL_FMA(kd2, EH[3], PRL_FMA(kd2, EH[2], ehi));
    const double rr = PRL_ADD(elo, r0);
    const int idx = 2 * (int)(ki & 127);
    const double tail = prl_u2d(PRL_PTAB(prl_pow_exp_tab, idx));
    const uint64_t sbits = PRL_PTAB(prl_pow_exp_tab, idx + 1) + (ki << 45);
    const double c23 = PRL_FMA(rr, EH[5], EH[4]);
    const double tr = PRL_ADD(rr, tail);
    const double r2 = PRL_MUL(rr, rr);
    const double c45 = PRL_FMA(rr, EH[7], EH[6]);
    const double p1 = PRL_FMA(c23, r2, tr);
    const double r4 = PRL_MUL(r2, r2);
    const double tmpv = PRL_FMA(c45, r4, p1);
    if (special) {
        if ((ki & 0x80000000ull) == 0) {  // k > 0: scale by 2^-1009 first
            const double sc = prl_u2d(sbits - (1009ull << 52));
            return PRL_MUL(PRL_FMA(sc, tmpv, sc), 0x1p1009);
        }
        const double sc = prl_u2d(sbits + (1022ull << 52));  // k < 0
        const double yv = PRL_ADD(sc, PRL_MUL(sc, tmpv));
        if (fabs(yv) >= 1.0) return PRL_MUL(0x1p-1022, yv);
        return PRL_MUL(x, x);  // subnormal result: libm rounds it specially; never reached by env states
    }
    const double scale = prl_u2d(sbits);
    return PRL_FMA(tmpv, scale, scale);
}

// libm powf(x, 2.0f)
PRL_POW_FN float powf2_glibc(float x) {
    constexpr double A[5] = {PRL_POWF_LOG2_POLY};
    constexpr double EH[4] = {PRL_POWF_EXP2_HDR};  // SHIFT (scaled), C0, C1, C2
    const uint32_t ix = prl_f2u(x) & 0x7fffffffu;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) return PRL_FMULF(x, x);
    // ---- log2_inline
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15);
    const uint32_t top = tmp & 0xff800000u;
    const double z = (double)prl_u2f(ix - top);
    const int k = (int32_t)top >> 23;
    const double invc = PRL_PTAB(prl_powf_log2_tab, 2 * i), logc = PRL_PTAB(prl_powf_log2_tab, 2 * i + 1);
    const double r = PRL_FMA(z, invc, -1.0);
    const double y0 = PRL_ADD(logc, (double)k);
    const double y = PRL_FMA(r, A[0], A[1]);
    const double p = PRL_FMA(r, A[2], A[3]);
    const double r2 = PRL_MUL(r, r);
    double q = PRL_FMA(r, A[4], y0);
    const double r4 = PRL_MUL(r2, r2);
    q = PRL_FMA(r2, p, q);
    const double logx = PRL_FMA(y, r4, q);
    const double ylogx = PRL_MUL(2.0, logx);
    if (((prl_d2u(ylogx) >> 47) & 0xffff) > 0x80be) return PRL_FMULF(x, x);  // |y log2 x| >= 126
    // ---- exp2_inline
    const double ks = PRL_ADD(ylogx, EH[0]);
    const uint64_t ki = prl_d2u(ks);
    const double kd = PRL_SUB(ks, EH[0]);
    const double rr = PRL_SUB(ylogx, kd);
    const uint64_t t = PRL_PTAB(prl_powf_exp2_tab, (int)(ki & 31)) + (ki << 47);
    const double zz = PRL_FMA(rr, EH[1], EH[2]);
    const double rr2 = PRL_MUL(rr, rr);
    const double yy = PRL_FMA(rr, EH[3], 1.0);
    const double res = PRL_MUL(PRL_FMA(zz, rr2, yy), prl_u2d(t));
    return (float)res;
}

}  // namespace prl
