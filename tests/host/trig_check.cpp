// TEST HARNESS (not product code): compiles csrc/trig_glibc.cuh for the host and compares it with this
// box's libm sin/cos bit for bit over the argument ranges the classic-control envs reach.
//   g++ -O2 -ffp-contract=off -DPRL_TRIG_HOST -I<csrc> trig_check.cpp -o trig_check -lm
// prints "<range> n=<samples> sin_mismatch=<k> cos_mismatch=<k>" per range; exit code = any mismatch.
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "trig_glibc.cuh"

static uint64_t s_rng = 0x9E3779B97F4A7C15ull;
static inline double urand() {  // xorshift64*, 53-bit uniform in [0,1)
    s_rng ^= s_rng >> 12; s_rng ^= s_rng << 25; s_rng ^= s_rng >> 27;
    return (double)((s_rng * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 2000000;
    struct { const char *name; double lo, hi; } ranges[] = {
        {"tiny      [-1e-7,1e-7]", -1e-7, 1e-7},    {"taylor    [-0.126,0.126]", -0.126, 0.126},
        {"cartpole  [-0.21,0.21]", -0.21, 0.21},    {"table     [-0.8555,0.8555]", -0.8555, 0.8555},
        {"hp        [0.85,2.43]", 0.85, 2.43},      {"-hp       [-2.43,-0.85]", -2.43, -0.85},
        {"acrobot   [-8,8]", -8.0, 8.0},            {"pendulum  [-90,90]", -90.0, 90.0},
        {"wide      [-1e5,1e5]", -1e5, 1e5},        {"near-max  [-1.05e8,1.05e8]", -1.05e8, 1.05e8},
    };
    int bad_total = 0;
    for (auto &r : ranges) {
        long bs = 0, bc = 0;
        for (long i = 0; i < n; ++i) {
            double x = r.lo + (r.hi - r.lo) * urand();
            double a = prl_trig::sin_glibc(x), b = sin(x), c = prl_trig::cos_glibc(x), d = cos(x);
            if (memcmp(&a, &b, 8)) { if (!bs) printf("  first sin mismatch x=%a got=%a want=%a\n", x, a, b); ++bs; }
            if (memcmp(&c, &d, 8)) { if (!bc) printf("  first cos mismatch x=%a got=%a want=%a\n", x, c, d); ++bc; }
        }
        printf("%-28s n=%ld sin_mismatch=%ld cos_mismatch=%ld\n", r.name, n, bs, bc);
        bad_total += (bs || bc);
    }
    // exact multiples / special points
    const double pts[] = {0.0, -0.0, 0.126, 0.855469, 2.426265, 0x1.921fb54442d18p+0, 0x1.921fb54442d18p+1, 1.0, -1.0, 0.5, 3.0, 6.0, 12.566370614359172};
    for (double x : pts) {
        double a = prl_trig::sin_glibc(x), b = sin(x), c = prl_trig::cos_glibc(x), d = cos(x);
        if (memcmp(&a, &b, 8) || memcmp(&c, &d, 8)) { printf("special point mismatch x=%a\n", x); bad_total++; }
    }
    return bad_total ? 1 : 0;
}
