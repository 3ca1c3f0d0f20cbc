// Host build of csrc/np_rng.cuh (the header is __host__ __device__): lets the CPU suite check the very code the GPU
// runs against numpy without a device.  g++ -O2 -shared -fPIC -D__host__= -D__device__= ... (tests/test_host.py).
#include <cstddef>
#include "np_rng.cuh"

extern "C" void np_rng_check(const uint64_t *seeds, int n, int draws, uint64_t *out_raw, double *out_uniform) {
    for (int i = 0; i < n; ++i) {
        prl::Pcg64 g = prl::Pcg64::from_seed(seeds[i]);
        for (int k = 0; k < draws; ++k) out_raw[(size_t)i * draws + k] = g.next64();
        for (int k = 0; k < draws; ++k) out_uniform[(size_t)i * draws + k] = g.next_double();
    }
}
