// umma_selftest.cu -- exercises every tcgen05 operand configuration kernel A relies on, in isolation:
// K-major A x K-major B (forward), K-major A x MN-major B (dgrad), MN-major A x MN-major B with
// M = 64 accumulators at TMEM lane offsets 0 and 16 and accumulation over K = 128 (wgrad), N = 16 /
// N = 32 shapes.  tests/test_gpu_umma.py compares the outputs with torch matmuls.
#include "umma.cuh"

namespace nsv {
namespace {

constexpr int kRows = 128;

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ W,
                                                               const __half* __restrict__ G, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* tA = smem;                    // [128][64]  16 KB
  unsigned char* tW = tA + 128 * 64 * 2;       // [64][64]    8 KB
  unsigned char* tG = tW + 64 * 64 * 2;        // [128][16]   4 KB
  unsigned char* tA32 = tG + 128 * 16 * 2;     // [128][32]   8 KB  (first 32 columns of A)
  unsigned char* tW32 = tA32 + 128 * 32 * 2;   // [64][32]    4 KB  (first 32 columns of W)
  unsigned char* tWo = tW32 + 64 * 32 * 2;     // [16][64]    2 KB  (first 16 rows of W)
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  umma::stage_tile(tA, A, 128, 64, tid, 128);
  umma::stage_tile(tW, W, 64, 64, tid, 128);
  umma::stage_tile(tG, G, 128, 16, tid, 128);
  umma::stage_tile(tWo, W, 16, 64, tid, 128);
  for (int i = tid; i < 128 * 4; i += 128) {  // column slices need a strided source
    const int r = i / 4, cg = i % 4;
    *reinterpret_cast<uint4*>(tA32 + umma::tile_off(r, cg * 8, 32)) = __ldg(reinterpret_cast<const uint4*>(A + (size_t)r * 64) + cg);
  }
  for (int i = tid; i < 64 * 4; i += 128) {
    const int r = i / 4, cg = i % 4;
    *reinterpret_cast<uint4*>(tW32 + umma::tile_off(r, cg * 8, 32)) = __ldg(reinterpret_cast<const uint4*>(W + (size_t)r * 64) + cg);
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_slot, 512);
  if (tid == 0) {
    umma::mbar_init(&mbar, 1);
    umma::mbar_fence_init();
  }
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_base_slot;

  if (tid == 0) {
    const uint32_t RG64 = 8 * 128, RG32 = 4 * 128, RG16 = 2 * 128;
    // T1 = A W^T : M128 N64 K64, A K-major, B K-major
    for (int k = 0; k < 4; ++k)
      umma::mma_f16(tm + 0, umma::smem_desc(umma::saddr(tA) + k * 256, 128, RG64), umma::smem_desc(umma::saddr(tW) + k * 256, 128, RG64),
                    umma::instr_desc(128, 64, false, false), k > 0);
    // T2 = A W : B[n][k] = W[k][n] -> W tile as MN-major B (LBO = row-group stride, SBO = 128)
    for (int k = 0; k < 4; ++k)
      umma::mma_f16(tm + 64, umma::smem_desc(umma::saddr(tA) + k * 256, 128, RG64),
                    umma::smem_desc(umma::saddr(tW) + k * 2 * RG64, RG64, 128), umma::instr_desc(128, 64, false, true), k > 0);
    // T3 = A^T A : M64 N64 K128, both MN-major; at lane offset 0 once, at lane offset 16 twice (accumulate)
    for (int rep = 0; rep < 3; ++rep) {
      const uint32_t d = tm + 128 + (rep == 0 ? 0u : (16u << 16));
      for (int k = 0; k < 8; ++k)
        umma::mma_f16(d, umma::smem_desc(umma::saddr(tA) + k * 2 * RG64, RG64, 128), umma::smem_desc(umma::saddr(tA) + k * 2 * RG64, RG64, 128),
                      umma::instr_desc(64, 64, true, true), (rep == 2) || k > 0);
    }
    // T4 = A Wo^T : M128 N16 K64
    for (int k = 0; k < 4; ++k)
      umma::mma_f16(tm + 192, umma::smem_desc(umma::saddr(tA) + k * 256, 128, RG64), umma::smem_desc(umma::saddr(tWo) + k * 256, 128, RG64),
                    umma::instr_desc(128, 16, false, false), k > 0);
    // T5 = G Wo : M128 N64 K16, A = G tile K-major (cols 16), B = Wo [16][64] MN-major
    umma::mma_f16(tm + 256, umma::smem_desc(umma::saddr(tG), 128, RG16), umma::smem_desc(umma::saddr(tWo), RG64, 128),
                  umma::instr_desc(128, 64, false, true), 0);
    // T6 = A^T G : M64 N16 K128, A = A tile MN-major, B = G tile MN-major
    for (int k = 0; k < 8; ++k)
      umma::mma_f16(tm + 320, umma::smem_desc(umma::saddr(tA) + k * 2 * RG64, RG64, 128), umma::smem_desc(umma::saddr(tG) + k * 2 * RG16, RG16, 128),
                    umma::instr_desc(64, 16, true, true), k > 0);
    // T7 = A32 W32^T : M128 N64 K32
    for (int k = 0; k < 2; ++k)
      umma::mma_f16(tm + 336, umma::smem_desc(umma::saddr(tA32) + k * 256, 128, RG32), umma::smem_desc(umma::saddr(tW32) + k * 256, 128, RG32),
                    umma::instr_desc(128, 64, false, false), k > 0);
    // T8 = A W32 : M128 N32 K64, B = W32 [64][32] MN-major
    for (int k = 0; k < 4; ++k)
      umma::mma_f16(tm + 400, umma::smem_desc(umma::saddr(tA) + k * 256, 128, RG64), umma::smem_desc(umma::saddr(tW32) + k * 2 * RG32, RG32, 128),
                    umma::instr_desc(128, 32, false, true), k > 0);
    // T9 = A32^T A : M32?? not used.  Instead wgrad of the first layer: dW0 = A^T A32 : M64 N32 K128
    for (int k = 0; k < 8; ++k)
      umma::mma_f16(tm + 432, umma::smem_desc(umma::saddr(tA) + k * 2 * RG64, RG64, 128), umma::smem_desc(umma::saddr(tA32) + k * 2 * RG32, RG32, 128),
                    umma::instr_desc(64, 32, true, true), k > 0);
    umma::commit(&mbar);
  }
  umma::mbar_wait(&mbar, 0);
  umma::fence_after_sync();

  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  // outputs, all row-major fp32: T1 [128,64] | T2 [128,64] | T3a [64,64] | T3b [64,64] | T4 [128,16] | T5 [128,64] | T6 [64,16] | T7 [128,64] | T8 [128,32] | T9 [64,32]
  float* o1 = out;
  float* o2 = o1 + 128 * 64;
  float* o3a = o2 + 128 * 64;
  float* o3b = o3a + 64 * 64;
  float* o4 = o3b + 64 * 64;
  float* o5 = o4 + 128 * 16;
  float* o6 = o5 + 128 * 64;
  float* o7 = o6 + 64 * 16;
  float* o8 = o7 + 128 * 64;
  float* o9 = o8 + 128 * 32;
  uint32_t r[32];
  auto dump128 = [&](float* dst, uint32_t col, int ncols) {
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      if (ncols - c0 >= 32) {
        umma::tmem_ld32(tm + lane_addr + col + c0, r);
        umma::tmem_ld_wait();
        for (int c = 0; c < 32; ++c) dst[(size_t)tid * ncols + c0 + c] = __uint_as_float(r[c]);
      } else {
        uint32_t q[16];
        umma::tmem_ld16(tm + lane_addr + col + c0, q);
        umma::tmem_ld_wait();
        for (int c = 0; c < 16; ++c) dst[(size_t)tid * ncols + c0 + c] = __uint_as_float(q[c]);
      }
    }
  };
  // M = 64 accumulators: row = 16 * warp + (lane % 16); lanes >= 16 belong to the accumulator at lane offset 16
  auto dump64 = [&](float* dst_lo, float* dst_hi, uint32_t col, int ncols) {
    const int row = 16 * warp + (lane & 15);
    float* dst = lane < 16 ? dst_lo : dst_hi;
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      if (ncols - c0 >= 32) {
        umma::tmem_ld32(tm + lane_addr + col + c0, r);
        umma::tmem_ld_wait();
        if (dst)
          for (int c = 0; c < 32; ++c) dst[(size_t)row * ncols + c0 + c] = __uint_as_float(r[c]);
      } else {
        uint32_t q[16];
        umma::tmem_ld16(tm + lane_addr + col + c0, q);
        umma::tmem_ld_wait();
        if (dst)
          for (int c = 0; c < 16; ++c) dst[(size_t)row * ncols + c0 + c] = __uint_as_float(q[c]);
      }
    }
  };
  dump128(o1, 0, 64);
  dump128(o2, 64, 64);
  dump64(o3a, o3b, 128, 64);
  dump128(o4, 192, 16);
  dump128(o5, 256, 64);
  dump64(o6, nullptr, 320, 16);
  dump128(o7, 336, 64);
  dump128(o8, 400, 32);
  dump64(o9, nullptr, 432, 32);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 512);
}

// Do accumulating MMAs issued by DIFFERENT threads into the SAME TMEM accumulator compose?  `n_issuers` warps' lane 0 each
// issue `reps` x (A^T A, M64 N64 K128 = 8 instructions, accumulate on) into one zero-initialised D without any ordering
// between them; every issuer commits to its own mbarrier.  out [64,64] must equal n_issuers * reps * A^T A.
__global__ void __launch_bounds__(128, 1) umma_shared_acc_kernel(const __half* __restrict__ A, float* __restrict__ out, int n_issuers, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* tA = smem;  // [128][64]
  __shared__ uint64_t mbar[4];
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  umma::stage_tile(tA, A, 128, 64, tid, 128);
  if (warp == 0) umma::tmem_alloc(&tmem_base_slot, 64);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) umma::mbar_init(&mbar[i], 1);
    umma::mbar_fence_init();
  }
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_base_slot;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < 64; c0 += 16) umma::tmem_st16_fill(tm + lane_addr + c0, 0u);
  umma::tmem_st_wait();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (lane == 0 && warp < n_issuers) {
    const uint32_t RG64 = 8 * 128;
    for (int rep = 0; rep < reps; ++rep)
      for (int k = 0; k < 8; ++k)
        umma::mma_f16(tm, umma::smem_desc(umma::saddr(tA) + k * 2 * RG64, RG64, 128), umma::smem_desc(umma::saddr(tA) + k * 2 * RG64, RG64, 128),
                      umma::instr_desc(64, 64, true, true), 1u);
    umma::commit(&mbar[warp]);
  }
  for (int i = 0; i < n_issuers; ++i) umma::mbar_wait(&mbar[i], 0);
  umma::fence_after_sync();
  uint32_t r[32];
  const int row = 16 * warp + (lane & 15);
  for (int c0 = 0; c0 < 64; c0 += 32) {
    umma::tmem_ld32(tm + lane_addr + c0, r);
    umma::tmem_ld_wait();
    if (lane < 16)
      for (int c = 0; c < 32; ++c) out[(size_t)row * 64 + c0 + c] = __uint_as_float(r[c]);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 64);
}

}  // namespace
}  // namespace nsv

extern "C" int nsv_umma_shared_accumulator_test(const void* A, float* out, int n_issuers, int reps, void* stream) {
  using namespace nsv;
  NSV_REQUIRE(A && out && n_issuers >= 1 && n_issuers <= 4 && reps >= 1, "nsv_umma_shared_accumulator_test: bad arguments");
  umma_shared_acc_kernel<<<1, 128, 128 * 64 * 2 + 128, (cudaStream_t)stream>>>((const __half*)A, out, n_issuers, reps);
  return check_launch("nsv_umma_shared_accumulator_test");
}

// A [128,64], W [64,64], G [128,16] fp16 row-major; out: 128*64*4 + 64*64*2 + 128*16 + 64*16 + 128*32 + 64*32 floats
extern "C" int nsv_umma_selftest(const void* A, const void* W, const void* G, float* out, void* stream) {
  using namespace nsv;
  NSV_REQUIRE(A && W && G && out, "nsv_umma_selftest: NULL pointer");
  const size_t smem = (128 * 64 + 64 * 64 + 128 * 16 + 128 * 32 + 64 * 32 + 16 * 64) * 2 + 256;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("nsv_umma_selftest: %s", cudaGetErrorString(e));
    return (int)e;
  }
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __half*)A, (const __half*)W, (const __half*)G, out);
  return check_launch("nsv_umma_selftest");
}
