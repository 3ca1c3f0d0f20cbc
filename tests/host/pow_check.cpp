// TEST HARNESS (not product code): compiles csrc/pow_glibc.cuh for the host and compares it with this box's libm
// pow(x, 2.0) / powf(x, 2.0f) bit for bit over the argument ranges the classic-control envs reach.
//   g++ -O2 -ffp-contract=off -fno-builtin -DPRL_TRIG_HOST -I<csrc> pow_check.cpp -o pow_check -lm
// prints "<range> n=<samples> pow_mismatch=<k> powf_mismatch=<k> (x*x differs: <k> / <k>)"; exit code = any mismatch.
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "pow_glibc.cuh"

static uint64_t s_rng = 0x9E3779B97F4A7C15ull;
static inline double urand() {
    s_rng ^= s_rng >> 12; s_rng ^= s_rng << 25; s_rng ^= s_rng >> 27;
    return (double)((s_rng * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 2000000;
    volatile double two = 2.0; volatile float twof = 2.0f;  // keep the compiler from folding pow(x, 2) into x*x
    struct { const char *name; double lo, hi; } ranges[] = {
        {"tiny      [-1e-7,1e-7]", -1e-7, 1e-7}, {"small     [-0.1,0.1]", -0.1, 0.1}, {"unit      [-1,1]", -1.0, 1.0},
        {"near one  [0.99,1.01]", 0.99, 1.01},    {"pi        [-3.2,3.2]", -3.2, 3.2},  {"speed     [-8,8]", -8.0, 8.0},
        {"acrobot   [-30,30]", -30.0, 30.0},      {"wide      [-1e6,1e6]", -1e6, 1e6},  {"huge      [1e100,1e150]", 1e100, 1e150}, {"minute    [1e-150,1e-100]", 1e-150, 1e-100},
    };
    int bad_total = 0;
    for (auto &r : ranges) {
        long bd = 0, bf = 0, dd = 0, df = 0;
        for (long i = 0; i < n; ++i) {
            double x = r.lo + (r.hi - r.lo) * urand();
            double a = prl::pow2_glibc(x), b = pow(x, two), c = x * x;
            if (memcmp(&a, &b, 8)) { if (!bd) printf("  first pow mismatch x=%a got=%a want=%a\n", x, a, b); ++bd; }
            if (memcmp(&c, &b, 8)) ++dd;
            float xf = (float)x;
            if (fabsf(xf) < 1e18f) {
                float af = prl::powf2_glibc(xf), bf_ = powf(xf, twof), cf = xf * xf;
                if (memcmp(&af, &bf_, 4)) { if (!bf) printf("  first powf mismatch x=%a got=%a want=%a\n", xf, af, bf_); ++bf; }
                if (memcmp(&cf, &bf_, 4)) ++df;
            }
        }
        printf("%-26s n=%ld pow_mismatch=%ld powf_mismatch=%ld (x*x differs from libm: %ld / %ld)\n", r.name, n, bd, bf, dd, df);
        bad_total += (bd || bf);
    }
    const double pts[] = {0.0, -0.0, 1.0, -1.0, 0.5, 2.0, 3.0, 1e-300, 1e-310, 1e200, 0x1.fffffffffffffp-1, 0x1.0000000000001p+0};
    for (double x : pts) {
        double a = prl::pow2_glibc(x), b = pow(x, two);
        float xf = (float)x, af = prl::powf2_glibc(xf), bf = powf(xf, twof);
        if (memcmp(&a, &b, 8)) { printf("special point pow mismatch x=%a got=%a want=%a\n", x, a, b); bad_total++; }
        if (memcmp(&af, &bf, 4)) { printf("special point powf mismatch x=%a got=%a want=%a\n", xf, af, bf); bad_total++; }
    }
    return bad_total ? 1 : 0;
}
