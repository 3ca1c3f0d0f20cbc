// fp64 sin/cos that return the SAME BITS as glibc's libm (>= 2.28, x86-64 FMA build) for |x| < 105414350.
//
// Why: the reference steps gymnasium classic-control envs in numpy fp64 (np.sin/np.cos == libm here,
// SURVEY.md 7.3 H1); CUDA's sin()/cos() round differently, so observations would not be reproducible.
// This is a restatement of the published algorithm glibc uses (IBM Accurate Mathematical Library,
// sysdeps/ieee754/dbl-64/s_sin.c: do_sin / do_cos / TAYLOR_SIN / reduce_sincos) with every rounding
// spelled out: each PRL_FMA below is an FMA in libm's x86-64 `-mfma` multiarch variant
// (__sin_fma/__cos_fma, selected by ifunc on any CPU with FMA+AVX2), every other operation rounds on its own.
// The operation sequence was established by reading that variant's machine code; the constants are the
// algorithm's published ones; the table is produced by tools/gen_sincostab.py.
//
// Attribution: the algorithm and its constants are those of the GNU C Library's sin / cos, which derive from the IBM Accurate
// Mathematical Library (Copyright (C) 2001 Free Software Foundation, Inc., written by International Business Machines Corp.;
// LGPL-2.1-or-later).  No glibc source text is included here: this file is an independent restatement of that algorithm.
//
// |x| >= 105414350 (glibc's __branred path) is not implemented: classic-control angles never get there
// (Pendulum |theta| <= pi + 8*0.05*T).  Such inputs return NaN so that a mismatch is loud, not silent.
//
// The same header compiles on the host with -DPRL_TRIG_HOST (tests/host/trig_check.cpp) so that the exact
// source can be checked against libm on the CPU; the product only ever uses the device build.
#pragma once
#include <stdint.h>

#ifdef PRL_TRIG_HOST
#include <math.h>
#include <string.h>
#define PRL_TRIG_FN static inline
#define PRL_FMA(a, b, c) fma((a), (b), (c))
#define PRL_MUL(a, b) ((a) * (b))
#define PRL_ADD(a, b) ((a) + (b))
#define PRL_SUB(a, b) ((a) - (b))
static inline int32_t prl_lo32(double x) { uint64_t u; memcpy(&u, &x, 8); return (int32_t)(uint32_t)u; }
static inline int32_t prl_hi32(double x) { uint64_t u; memcpy(&u, &x, 8); return (int32_t)(u >> 32); }
static const double prl_sincostab[440] = {
#include "sincostab.inc"
};
#define PRL_TAB(i) prl_sincostab[i]
#define PRL_NAN (0.0 / 0.0)
#else
#define PRL_TRIG_FN __device__ __forceinline__
#define PRL_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PRL_MUL(a, b) __dmul_rn((a), (b))
#define PRL_ADD(a, b) __dadd_rn((a), (b))
#define PRL_SUB(a, b) __dsub_rn((a), (b))
__device__ __forceinline__ int32_t prl_lo32(double x) { return __double2loint(x); }
__device__ __forceinline__ int32_t prl_hi32(double x) { return __double2hiint(x); }
__device__ const double prl_sincostab[440] = {
#include "sincostab.inc"
};
#define PRL_TAB(i) __ldg(&prl_sincostab[i])
#define PRL_NAN __longlong_as_double(0x7ff8000000000000LL)
#endif

namespace prl_trig {

// Taylor coefficients of sin (|x| < 0.126) and the short sin/cos polynomials used around table points
constexpr double S1 = -0x1.5555555555555p-3, S2 = 0x1.1111111110ecep-7, S3 = -0x1.a01a019db08b8p-13,
                 S4 = 0x1.71de27b9a7ed9p-19, S5 = -0x1.addffc2fcdf59p-26;
constexpr double SN3 = -0x1.5555555555515p-3, SN5 = 0x1.11110e829872fp-7;
constexpr double CS2 = 0.5, CS4 = -0x1.5555555555535p-5, CS6 = 0x1.6c16bedd9e239p-10;
constexpr double BIG = 0x1.8p45;    // ulp(BIG) = 2^-7: BIG + |x| rounds |x| to the nearest k/128
constexpr double TOINT = 0x1.8p52;
constexpr double HPINV = 0x1.45f306dc9c883p-1;  // 2/pi
constexpr double HP0 = 0x1.921fb54442d18p+0, HP1 = 0x1.1a62633145c07p-54;  // pi/2 = HP0 + HP1
constexpr double MP1 = 0x1.921fb58000000p+0, MP2 = -0x1.dde973c000000p-27;
constexpr double PP3 = -0x1.cb3b398000000p-55, PP4 = -0x1.d747f23e32ed7p-83;

PRL_TRIG_FN double taylor_sin(double xx, double x, double dx) {
    double p = PRL_FMA(xx, S5, S4);
    p = PRL_FMA(xx, p, S3);
    p = PRL_FMA(xx, p, S2);
    p = PRL_FMA(xx, p, S1);
    double t = PRL_FMA(PRL_FMA(p, x, -PRL_MUL(dx, 0.5)), xx, dx);
    return PRL_ADD(x, t);
}

// sin(x + dx), |x| < ~0.86, |dx| tiny
PRL_TRIG_FN double do_sin(double x, double dx) {
    const double ax = fabs(x);
    if (ax < 0.126) return taylor_sin(PRL_MUL(x, x), x, dx);
    if (x <= 0) dx = -dx;
    const double u = PRL_ADD(BIG, ax);
    const double xr = PRL_SUB(ax, PRL_SUB(u, BIG));
    const double xx = PRL_MUL(xr, xr);
    const double s = PRL_ADD(xr, PRL_FMA(PRL_MUL(xr, xx), PRL_FMA(xx, SN5, SN3), dx));
    const double c = PRL_FMA(xr, dx, PRL_MUL(xx, PRL_FMA(xx, PRL_FMA(xx, CS6, CS4), CS2)));
    const int k = prl_lo32(u) << 2;
    const double sn = PRL_TAB(k), ssn = PRL_TAB(k + 1), cs = PRL_TAB(k + 2), ccs = PRL_TAB(k + 3);
    const double cor = PRL_FMA(s, cs, PRL_FMA(-c, sn, PRL_FMA(s, ccs, ssn)));
    return copysign(PRL_ADD(sn, cor), x);
}

// cos(x + dx)
PRL_TRIG_FN double do_cos(double x, double dx) {
    if (x < 0) dx = -dx;
    const double ax = fabs(x);
    const double u = PRL_ADD(BIG, ax);
    const double xr = PRL_ADD(PRL_SUB(ax, PRL_SUB(u, BIG)), dx);
    const double xx = PRL_MUL(xr, xr);
    const double s = PRL_FMA(PRL_MUL(xr, xx), PRL_FMA(xx, SN5, SN3), xr);
    const double c = PRL_MUL(xx, PRL_FMA(xx, PRL_FMA(xx, CS6, CS4), CS2));
    const int k = prl_lo32(u) << 2;
    const double sn = PRL_TAB(k), ssn = PRL_TAB(k + 1), cs = PRL_TAB(k + 2), ccs = PRL_TAB(k + 3);
    const double cor = PRL_FMA(-s, sn, PRL_FMA(-c, cs, PRL_FMA(-s, ssn, ccs)));
    return PRL_ADD(cs, cor);
}

// x = n*(pi/2) + (a + da), returns n & 3
PRL_TRIG_FN int reduce_sincos(double x, double &a, double &da) {
    const double t = PRL_FMA(x, HPINV, TOINT);
    const double xn = PRL_SUB(t, TOINT);
    const int n = prl_lo32(t) & 3;
    const double y = PRL_FMA(-xn, MP2, PRL_FMA(-xn, MP1, x));
    const double t2 = PRL_FMA(-xn, PP3, y);
    double db = PRL_FMA(-xn, PP3, PRL_SUB(y, t2));
    const double b = PRL_FMA(-xn, PP4, t2);
    db = PRL_ADD(db, PRL_FMA(-xn, PP4, PRL_SUB(t2, b)));
    a = b;
    da = db;
    return n;
}

PRL_TRIG_FN double do_sincos(double a, double da, int n) {
    double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
    return (n & 2) ? -r : r;
}

PRL_TRIG_FN double sin_glibc(double x) {
    const int k = prl_hi32(x) & 0x7fffffff;
    if (k < 0x3e500000) return x;                       // |x| < 2^-26
    if (k < 0x3feb6000) return do_sin(x, 0.0);          // |x| < 0.855469
    if (k < 0x400368fd) {                               // |x| < 2.426265
        const double t = PRL_SUB(HP0, fabs(x));
        return copysign(do_cos(t, HP1), x);
    }
    if (k < 0x419921FB) {                               // |x| < 105414350
        double a, da;
        const int n = reduce_sincos(x, a, da);
        return do_sincos(a, da, n);
    }
    return PRL_NAN;
}

PRL_TRIG_FN double cos_glibc(double x) {
    const int k = prl_hi32(x) & 0x7fffffff;
    if (k < 0x3e400000) return 1.0;                     // |x| < 2^-27
    if (k < 0x3feb6000) return do_cos(x, 0.0);
    if (k < 0x400368fd) {
        const double y = PRL_SUB(HP0, fabs(x));
        const double a = PRL_ADD(y, HP1);
        const double da = PRL_ADD(PRL_SUB(y, a), HP1);
        return do_sin(a, da);
    }
    if (k < 0x419921FB) {
        double a, da;
        const int n = reduce_sincos(x, a, da);
        return do_sincos(a, da, n + 1);
    }
    return PRL_NAN;
}

}  // namespace prl_trig
