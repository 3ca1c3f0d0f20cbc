// Thin wrappers over the Blackwell (sm_100a) tensor-core instructions used by the fused update kernel:
// tcgen05.mma (kind::f16 on bf16 operands in shared memory, fp32 accumulator in tensor memory), tensor-memory
// allocation and loads, mbarrier completion tracking.  Inline PTX only; no library.
//
// Precision: an fp32 value x is split into three bf16 pieces x = p0 + p1 + p2 (8 mantissa bits each, exact), and a
// product a*b is accumulated as p0p0 + p0p1 + p1p0 + p1p1 + p0p2 + p2p0 in the fp32 accumulator - fp32-grade results
// (dropped terms <= 2^-24 |ab|) at six bf16 MMAs per GEMM.  (kind::tf32 would need 3 MMAs of half the K, the same
// tensor time, but tcgen05 accepts MN-major tf32 operands only in a 128B-swizzled layout that cannot double as a
// K-major operand, which would double the shared-memory footprint; measured: the no-swizzle MN-major tf32 MMA
// returns zeros.)
//
// Shared-memory operand layout used throughout (no swizzle, "interleaved" canonical layout of 8 x 16-byte core
// matrices).  A row-per-thread activation matrix Y[r][c] (r = 0..127 rows of the tile, c = 0..C-1 features, bf16) is
// stored as
//        byte(r, c) = (c / 8) * CHUNK + r * 16 + (c % 8) * 2,            CHUNK = 128 rows * 16 B = 2048
// so thread r writes its row with 16-byte stores that are contiguous across the warp (conflict-free).  The SAME
// bytes are a valid tcgen05 operand in two ways:
//   * K-major   with MN = r, K = c :  stride between 8-row groups (SBO) = 128, between 16-byte K chunks (LBO) = CHUNK;
//                                     one instruction covers K = 16 (two chunks), K-step s starts at + s * 2 * CHUNK
//   * MN-major  with MN = c, K = r :  stride between 8-element MN groups (SBO) = CHUNK, between 8-row K groups (LBO)
//                                     = 128; K-step s (rows 16s..16s+15) starts at + s * 256
// which is what lets one staging of the activations feed both the forward/dgrad GEMMs (contraction over features) and
// the weight-gradient GEMM (contraction over rows).  Weight matrices W[n][k] use the same scheme with n in the role
// of r:  byte(n, k) = (k / 8) * (N * 16) + n * 16 + (k % 8) * 2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prl {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- descriptors -------------------------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor: start address, leading / stride byte offsets (all >> 4), version 1, no swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// same with a swizzle mode in bits [61,64): 0 none, 2 = 128-byte swizzle, 4 = 64-byte, 6 = 32-byte
__device__ __forceinline__ uint64_t smem_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)layout_type << 61);
}
// 32-bit instruction descriptor, kind::f16 with bf16 operands, fp32 accumulate.  a_mn / b_mn: 1 = MN-major, 0 = K-major.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// ---- tensor memory -----------------------------------------------------------------------------------------------
// one full warp; writes the base address (lane 0, first column) to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 consecutive 32-bit columns of this thread's lane (warp w reads lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- mbarrier ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Bounded wait: returns false if the phase did not complete within ~2^22 polls (a mis-programmed MMA must not hang
// the GPU); the caller records the failure and carries on.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
#pragma unroll 1
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

// one lane of the (converged) warp; the rest of the warp skips the guarded block
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- MMA ---------------------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]; one thread issues.  accumulate = 0 overwrites D.
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the weight-stationary form (UTCHMMA.WS): same operands, a different placement of the accumulator rows in tensor memory for
// M < 128 (probed by tools/probe_m64.py; not used by the product kernels)
__device__ __forceinline__ void mma_bf16_ss_ws(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.ws.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on `bar` when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- fp32 -> three bf16 pieces ------------------------------------------------------------------------------------
// (x0, x1) -> packed bf16x2 pieces q0, q1, q2 with x = p0 + p1 + p2 up to 2^-24 |x| (low half = x0, high half = x1)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void split_bf16x3(float x0, float x1, uint32_t &q0, uint32_t &q1, uint32_t &q2) {
    q0 = pack_bf16x2(x0, x1);
    const float r0 = x0 - __uint_as_float(q0 << 16), r1 = x1 - __uint_as_float(q0 & 0xffff0000u);
    q1 = pack_bf16x2(r0, r1);
    const float s0 = r0 - __uint_as_float(q1 << 16), s1 = r1 - __uint_as_float(q1 & 0xffff0000u);
    q2 = pack_bf16x2(s0, s1);
}

constexpr int TILE_ROWS = 128;
constexpr int CHUNK = TILE_ROWS * 16;  // bytes between 4-feature chunks of a row-per-thread matrix

}  // namespace umma
}  // namespace prl
