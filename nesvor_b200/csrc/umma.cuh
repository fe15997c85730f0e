// umma.cuh -- thin inline-PTX layer over Blackwell's 5th-generation tensor cores (tcgen05) as used
// by kernel A: shared-memory operand descriptors for the un-swizzled canonical layout, the
// kind::f16 instruction descriptor, MMA issue / commit, TMEM allocation and TMEM -> register loads.
//
// Operand tiles.  A [rows][cols] fp16 tile is stored as 8x8 "core matrices" (8 rows x 16 bytes):
//     addr(r, c) = (r / 8) * RG + (c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2,   RG = (cols / 8) * 128
// This one layout is, for the tensor core,
//   * a K-major operand with MN = rows, K = cols   (LBO = 128, SBO = RG)          -> forward, dgrad A
//   * an MN-major operand with MN = cols, K = rows (LBO = RG,  SBO = 128)         -> wgrad A and B,
//     and weights [out][in] as the B operand of dgrad (N = in, K = out)
// (descriptor semantics: SBO = byte stride between 8-wide groups along MN, LBO = byte stride
// between 8-wide groups along K; cf. cute/atom/mma_traits_sm100.hpp make_umma_desc).
#pragma once
#include "nsv_common.cuh"

namespace nsv {
namespace umma {

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, c) inside a canonical tile with `cols` columns
__device__ __forceinline__ uint32_t tile_off(int r, int c, int cols) {
  return (uint32_t)((r >> 3) * (cols >> 3) * 128 + (c >> 3) * 128 + (r & 7) * 16 + (c & 7) * 2);
}

// shared-memory matrix descriptor, SWIZZLE_NONE, sm_100 version bit set
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// kind::f16 instruction descriptor: fp16 A/B, fp32 accumulate
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues on behalf of the CTA
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// arrive on an mbarrier when every previously issued MMA of this thread has completed
__device__ __forceinline__ void commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(mbar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* mbar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// try_wait parks the thread in hardware until the phase completes or the suspend-time hint (ns) expires: a long
// wait costs a handful of issue slots instead of a spin loop's
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}\n" ::"r"(saddr(mbar)),
      "r"(parity)
      : "memory");
}

// ---- TMA engine, non-tensor form: one bulk copy global -> shared, completion counted in bytes on an mbarrier ----
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(mbar)), "r"(bytes) : "memory");
}
// `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(saddr(mbar))
               : "memory");
}

// TMEM allocation (one warp, all lanes); the base address is written to *slot in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, int ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, int ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}

// TMEM -> registers: each lane reads N consecutive 32-bit columns of its own TMEM lane
// (lane = 32 * (warp % 4) + laneid must be encoded in bits [16,32) of taddr by the caller)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: 16 consecutive 32-bit columns of the caller's own TMEM lane, all set to `v`
__device__ __forceinline__ void tmem_st16_fill(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// cooperative copy: row-major global [rows][cols] fp16 -> canonical shared tile (16-byte chunks)
__device__ __forceinline__ void stage_tile(unsigned char* tile, const __half* __restrict__ src, int rows, int cols, int tid, int nthreads) {
  const int cpr = cols >> 3;
  for (int i = tid; i < rows * cpr; i += nthreads) {
    const int r = i / cpr, cg = i % cpr;
    *reinterpret_cast<uint4*>(tile + tile_off(r, cg * 8, cols)) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * cols) + cg);
  }
}

}  // namespace umma
}  // namespace nsv
