/* prl_b200_test.h - C ABI of libprl_b200_test.so: parity-test hooks, kept OUT of the product library (libprl_b200.so,
 * include/prl_b200.h).  Device builds of the bit-exact math headers (csrc/trig_glibc.cuh, pow_glibc.cuh, np_rng.cuh, Philox) and
 * of the tensor-core building blocks (csrc/umma.cuh), callable element by element so that tests/ can compare them with libm,
 * numpy and exact integer arithmetic.  Same conventions as prl_b200.h: plain device pointers, int status (0 = ok),
 * prl_test_last_error() for the message. */
#ifndef PRL_B200_TEST_H
#define PRL_B200_TEST_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char *prl_test_last_error(void);
int prl_test_sincos(const double *x, double *sin_out, double *cos_out, int64_t n, void *stream);
int prl_test_pow2(const double *x, double *out, const float *xf, float *outf, int64_t n, void *stream);
/* out[i][k] = k-th 64-bit output of np.random.PCG64(np.random.SeedSequence(seeds[i])) */
int prl_test_pcg64(const uint64_t *seeds, int n, int draws, uint64_t *out, void *stream);
int prl_test_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out4, void *stream);
/* tensor-core building blocks (csrc/umma.cuh): one 128-row tile, tcgen05.mma kind::tf32 from shared memory.
 * mode 0: D[128][128] = A[128][64] B[128][64]^T; 1: D[128][64] = A[128][64] B[64:128][0:64]; 2: D[128][64] =
 * A[128][128]^T B[128][64]; 3: D[128][16] = A[128][128]^T B[128][16]; -1: raw descriptor parameters in cfg_host[17]
 * (see csrc/umma_test.cu).  *status != 0: the MMA never completed. */
int prl_test_umma(int mode, const float *A, const float *B, float *D, int *status, const int32_t *cfg_host, void *stream);


#ifdef __cplusplus
}
#endif
#endif
