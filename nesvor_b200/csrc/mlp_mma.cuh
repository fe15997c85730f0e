// mlp_mma.cuh -- warp-level building blocks of the fully fused small MLP (fp16 operands, fp32
// accumulate) shared by mlp.cu (standalone op behind build_network's fp16 branch,
// nesvor/nesvor/models.py:28-41) and inr_fused.cu (kernel A).
//
// A warp owns 32 rows (= 32 PSF samples) of a CTA tile: two m16 row tiles.  Layer inputs live in
// registers as mma A-fragments; weights live in shared memory, row-major [out][in], rows padded by
// 8 halves so that ldmatrix / fragment stores are bank-conflict free.  Between layers the fp32
// accumulator fragment is ReLU'd, rounded to fp16 and re-packed in registers into the next
// layer's A-fragment (accumulator cols {2t,2t+1} of n-tiles 2k,2k+1 == A cols of k-tile k), so
// activations never leave the register file on the forward critical path; a copy is parked in
// shared memory for the backward pass (ReLU mask, wgrad operand).
//
//   forward  : C[32 x out]  = A[32 x in]  * W^T          (B fragment = ldmatrix of W)
//   dgrad    : dA[32 x in]  = dC[32 x out] * W            (B fragment = ldmatrix.trans of W)
//   wgrad    : dW[out x in] += dC^T[out x rows] * A[rows x in]   (both operands ldmatrix.trans of
//              the activation tiles; K runs over all rows of the CTA tile; each warp owns an
//              (m16 x n-tiles) block of dW and keeps it in registers across tiles)
#pragma once
#include "nsv_common.cuh"

namespace nsv {

constexpr int kPad = 8;  // halves of padding per shared-memory row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

// ---- A fragments of the warp's 32 rows from a [rows][ld] fp16 shared tile (cols k0..k0+15) ----
template <int KT, int MTL>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[MTL][KT][4], const __half* tile, int ld, int row0) {
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int m = 0; m < MTL; ++m)
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
      ldsm_x4(a[m][kt], tile + (size_t)(row0 + m * 16 + r + ((mi & 1) ? 8 : 0)) * ld + kt * 16 + ((mi >> 1) ? 8 : 0));
}

// ---- C[32 x 8*NT] = A * W^T, W row-major [8*NT][16*KT (+pad)] in shared memory ----
template <int KT, int NT, int MTL>
__device__ __forceinline__ void warp_gemm_fwd(float (&c)[MTL][NT][4], const uint32_t (&a)[MTL][KT][4], const __half* w, int ldw) {
  static_assert(NT % 2 == 0, "n-tiles come in pairs");
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int m = 0; m < MTL; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int k = 0; k < 4; ++k) c[m][n][k] = 0.f;
#pragma unroll
  for (int kt = 0; kt < KT; ++kt)
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldsm_x4(b, w + (size_t)(np * 16 + r + ((mi >> 1) ? 8 : 0)) * ldw + kt * 16 + ((mi & 1) ? 8 : 0));
#pragma unroll
      for (int m = 0; m < MTL; ++m) {
        mma16816(c[m][2 * np], a[m][kt], b[0], b[1]);
        mma16816(c[m][2 * np + 1], a[m][kt], b[2], b[3]);
      }
    }
}

// ---- dA[32 x 8*NT] = dC[32 x 16*KT] * W, W row-major [16*KT][8*NT (+pad)] ----
template <int KT, int NT, int MTL>
__device__ __forceinline__ void warp_gemm_dgrad(float (&c)[MTL][NT][4], const uint32_t (&a)[MTL][KT][4], const __half* w, int ldw) {
  static_assert(NT % 2 == 0, "n-tiles come in pairs");
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int m = 0; m < MTL; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int k = 0; k < 4; ++k) c[m][n][k] = 0.f;
#pragma unroll
  for (int kt = 0; kt < KT; ++kt)
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, w + (size_t)(kt * 16 + r + ((mi & 1) ? 8 : 0)) * ldw + np * 16 + ((mi >> 1) ? 8 : 0));
#pragma unroll
      for (int m = 0; m < MTL; ++m) {
        mma16816(c[m][2 * np], a[m][kt], b[0], b[1]);
        mma16816(c[m][2 * np + 1], a[m][kt], b[2], b[3]);
      }
    }
}

// ---- accumulator fragment -> next layer's A fragment (optionally ReLU), all in registers ----
template <int NT, bool kRelu, int MTL>
__device__ __forceinline__ void acc_to_a(uint32_t (&a)[MTL][NT / 2][4], const float (&c)[MTL][NT][4]) {
#pragma unroll
  for (int m = 0; m < MTL; ++m)
#pragma unroll
    for (int kt = 0; kt < NT / 2; ++kt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v0 = c[m][2 * kt + h][0], v1 = c[m][2 * kt + h][1], v2 = c[m][2 * kt + h][2], v3 = c[m][2 * kt + h][3];
        if (kRelu) {
          v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
        }
        a[m][kt][2 * h] = pack_half2(v0, v1);
        a[m][kt][2 * h + 1] = pack_half2(v2, v3);
      }
}

// ---- ReLU bookkeeping in registers: bit (m*NT + n)*4 + k set <=> accumulator element > 0 ----
template <int NT, int MTL>
__device__ __forceinline__ uint64_t relu_bits(const float (&c)[MTL][NT][4]) {
  static_assert(MTL * NT * 4 <= 64, "mask must fit 64 bits");
  uint64_t bits = 0;
#pragma unroll
  for (int m = 0; m < MTL; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c[m][n][k] > 0.f) bits |= 1ull << ((m * NT + n) * 4 + k);
  return bits;
}
template <int NT, int MTL>
__device__ __forceinline__ void apply_relu_bits(float (&c)[MTL][NT][4], uint64_t bits) {
#pragma unroll
  for (int m = 0; m < MTL; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (!((bits >> ((m * NT + n) * 4 + k)) & 1ull)) c[m][n][k] = 0.f;
}

// ---- park an A-fragment set (fp16) into a [rows][ld] shared tile, cols 0..16*KT-1 ----
template <int KT, int MTL>
__device__ __forceinline__ void store_a_frags(const uint32_t (&a)[MTL][KT][4], __half* tile, int ld, int row0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < MTL; ++m)
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      __half* p = tile + (size_t)(row0 + m * 16 + g) * ld + kt * 16 + 2 * t;
      *reinterpret_cast<uint32_t*>(p) = a[m][kt][0];
      *reinterpret_cast<uint32_t*>(p + 8 * ld) = a[m][kt][1];
      *reinterpret_cast<uint32_t*>(p + 8) = a[m][kt][2];
      *reinterpret_cast<uint32_t*>(p + 8 * ld + 8) = a[m][kt][3];
    }
}

// ---- ReLU backward on fragments: zero dA where the parked activation (fp16, same layout) is <= 0 ----
template <int NT, int MTL>
__device__ __forceinline__ void relu_mask_acc(float (&c)[MTL][NT][4], const __half* act_tile, int ld, int row0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < MTL; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const __half* p = act_tile + (size_t)(row0 + m * 16 + g) * ld + n * 8 + 2 * t;
      const float2 lo = unpack_half2(*reinterpret_cast<const uint32_t*>(p));
      const float2 hi = unpack_half2(*reinterpret_cast<const uint32_t*>(p + 8 * ld));
      if (!(lo.x > 0.f)) c[m][n][0] = 0.f;
      if (!(lo.y > 0.f)) c[m][n][1] = 0.f;
      if (!(hi.x > 0.f)) c[m][n][2] = 0.f;
      if (!(hi.y > 0.f)) c[m][n][3] = 0.f;
    }
}

// ---- wgrad: this warp's block of dW[out][in] += sum over `rows` of dC[row][out] * A[row][in] ----
// Work split over the CTA's NW warps: m-tile (16 outputs) = warp % MT; the remaining NW/MT warps
// share the n-tiles (8 inputs each).  NTW = n-tiles owned by one warp.
template <int OUT, int IN, int NW>
struct WgradSplit {
  static constexpr int MT = OUT / 16, NTL = IN / 8;
  static constexpr int PARTS = (NW / MT) < 1 ? 1 : (NW / MT);
  static constexpr int NTW = (NTL + PARTS - 1) / PARTS;
  static_assert(OUT % 16 == 0 && IN % 8 == 0, "padded dims");
  static_assert(MT <= NW, "more m-tiles than warps is not instantiated");
};

template <int OUT, int IN, int NW>
__device__ __forceinline__ void warp_wgrad(float (&acc)[WgradSplit<OUT, IN, NW>::NTW][4], const __half* dc_tile, int ld_dc,
                                           const __half* a_tile, int ld_a, int rows, int warp = -1) {
  using S = WgradSplit<OUT, IN, NW>;
  if (warp < 0) warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
  const int mt = warp % S::MT, part = warp / S::MT;
  if (part >= S::PARTS) return;
  const int nt0 = part * S::NTW;
  for (int k0 = 0; k0 < rows; k0 += 16) {
    uint32_t a[4];
    ldsm_x4_t(a, dc_tile + (size_t)(k0 + r + ((mi >> 1) ? 8 : 0)) * ld_dc + mt * 16 + ((mi & 1) ? 8 : 0));
#pragma unroll
    for (int j = 0; j < S::NTW; ++j) {
      const int nt = nt0 + j;
      if (nt < S::NTL) {
        uint32_t b[2];
        ldsm_x2_t(b, a_tile + (size_t)(k0 + (lane & 7) + ((lane & 8) ? 8 : 0)) * ld_a + nt * 8);
        mma16816(acc[j], a, b[0], b[1]);
      }
    }
  }
}

// flush a warp's dW block to the global fp32 gradient (row-major [OUT][ld_g], logical cols < n_in)
template <int OUT, int IN, int NW>
__device__ __forceinline__ void flush_wgrad(const float (&acc)[WgradSplit<OUT, IN, NW>::NTW][4], float* __restrict__ g, int ld_g,
                                            float scale) {
  using S = WgradSplit<OUT, IN, NW>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int mt = warp % S::MT, part = warp / S::MT;
  if (part >= S::PARTS) return;
#pragma unroll
  for (int j = 0; j < S::NTW; ++j) {
    const int nt = part * S::NTW + j;
    if (nt >= S::NTL) continue;
    float* p = g + (size_t)(mt * 16 + gq) * ld_g + nt * 8 + 2 * t;
    red_add_v2(p, acc[j][0] * scale, acc[j][1] * scale);
    red_add_v2(p + 8 * (size_t)ld_g, acc[j][2] * scale, acc[j][3] * scale);
  }
}

// cooperative copy of a row-major fp16 weight matrix [out][in] from global into padded shared rows
__device__ __forceinline__ void stage_weights(__half* dst, int ld, const __half* __restrict__ src, int out, int in) {
  const int vec_per_row = in / 8;  // 16-byte chunks
  for (int i = threadIdx.x; i < out * vec_per_row; i += blockDim.x) {
    const int row = i / vec_per_row, v = i % vec_per_row;
    *reinterpret_cast<uint4*>(dst + (size_t)row * ld + v * 8) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)row * in) + v);
  }
}

}  // namespace nsv
