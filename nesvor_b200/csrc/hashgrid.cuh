// hashgrid.cuh -- multiresolution hash-grid geometry shared by the standalone encoding op and the
// fused INR kernel.  Semantics = tiny-cuda-nn's HashGrid as the reference uses it
// (nesvor/nesvor/models.py:22-25,102-111,146): per level
//   pos = fmaf(scale_l, x, 0.5);  g = (uint32)(int)floor(pos);  w = pos - floor(pos)
//   dense : idx = (gx + gy*res + gz*res^2) mod T_l          (uint32 wrap-around)
//   hashed: idx = (gx ^ gy*2654435761 ^ gz*805459861) mod T_l
// 8 corners, corner bit d set -> coordinate g_d + 1 and weight w_d, else 1 - w_d.
#pragma once
#include "nsv_common.cuh"

namespace nsv {

struct LevelGeom {
  float scale;
  uint32_t res, size, offset, hashed;
};

__device__ __forceinline__ LevelGeom level_geom(const nsv_grid_meta& m, int l) {
  LevelGeom g;
  g.scale = m.scale[l];
  g.res = m.res[l];
  g.size = m.size[l];
  g.offset = m.offset[l];
  g.hashed = m.hashed[l];
  return g;
}

__device__ __forceinline__ void level_pos(const float x[3], float scale, uint32_t g[3], float w[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float p = fmaf(scale, x[d], 0.5f);
    const float fl = floorf(p);
    g[d] = (uint32_t)(int)fl;
    w[d] = p - fl;
  }
}

// index of grid vertex (cx,cy,cz) inside its level (without the level offset)
__device__ __forceinline__ uint32_t vertex_index(const LevelGeom& lv, uint32_t cx, uint32_t cy, uint32_t cz) {
  uint32_t idx;
  if (lv.hashed) {
    idx = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
    // hashed levels are capped at 2^log2_T entries -> power of two unless the cap is not a power of two
    if ((lv.size & (lv.size - 1)) == 0) return idx & (lv.size - 1);
  } else {
    idx = cx + cy * lv.res + cz * (lv.res * lv.res);
  }
  return idx % lv.size;
}

__device__ __forceinline__ float corner_weight(int c, const float w[3]) {
  const float fx = (c & 1) ? w[0] : 1.f - w[0];
  const float fy = (c & 2) ? w[1] : 1.f - w[1];
  const float fz = (c & 4) ? w[2] : 1.f - w[2];
  return fx * fy * fz;
}

// table element loads: returns the F=2 feature pair of one entry as float2
__device__ __forceinline__ float2 load_pair(const float* __restrict__ table, uint32_t entry) {
  return __ldg(reinterpret_cast<const float2*>(table) + entry);
}
__device__ __forceinline__ float2 load_pair(const __half* __restrict__ table, uint32_t entry) {
  const __half2 h = __ldg(reinterpret_cast<const __half2*>(table) + entry);
  return __half22float2(h);
}

}  // namespace nsv
