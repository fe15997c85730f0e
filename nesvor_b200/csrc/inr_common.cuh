// inr_common.cuh -- device code shared by the two implementations of kernel A (inr_fused.cu: mma.sync
// fragments; inr_fused_tc.cu: tcgen05 / TMEM): argument block, level table, Philox noise, the paired-lane
// hash-grid gather / scatter, loss helpers and the 1-block finalize kernel.
#pragma once
#include "hashgrid.cuh"
#include "pose.cuh"

namespace nsv {
namespace fused {

constexpr int kTile = 256;     // sample rows per CTA
constexpr int kThreads = 512;  // 16 warps; a warp owns 16 rows, lane = (row = lane >> 1, x-corner = lane & 1)
constexpr int kIn = 32, kOutP = 16;
constexpr uint32_t kAggMaxSize = 8192;     // generic loops: levels with at most this many entries try warp-level gradient aggregation
constexpr uint32_t kAggDefault = 1u << 20;  // fast loops: every dense level pre-reduces inside the warp (measured best on B200)
constexpr int kAggMaxCells = 4;      // ... when the warp's 16 samples occupy at most this many cells
constexpr int kLddx = kIn + 1;

// ---- level metadata staged in shared memory (dynamic indexing of kernel params costs an LDC miss) ----
// `fast[l]` is the branch-free view used when the level order is dense-then-hashed and every hashed level
// has a power-of-two size (always true for tcnn-style grids): one 16-byte shared load per level.
struct FastLevel {
  float scale;
  uint32_t sy, sz;  // dense: strides res, res^2
  uint32_t last;    // dense: size - 1 (clamp for the speculative load); hashed: the index mask
};
struct LevelTable {
  float scale[kIn / 2];
  uint32_t res[kIn / 2], size[kIn / 2], offset[kIn / 2], hashed[kIn / 2];
  FastLevel fast[kIn / 2];
  const __half2* tbl[kIn / 2];  // first fp16 entry of the level
  float* grd[kIn / 2];          // first gradient element of the level
  uint32_t n_dense;   // levels [0, n_dense) are dense, [n_dense, n_levels) hashed
  uint32_t fast_ok;   // the branch-free loops apply to this grid
  uint32_t agg_last;  // dense levels with size - 1 <= agg_last pre-reduce their gradient inside the warp
  uint32_t ablate;    // profiling only (NSV_ABLATE): bit 0 skips the table loads, bit 1 the table reductions, bit 2 the MLP chain
};
static_assert(sizeof(FastLevel) == 16 && sizeof(LevelTable) % 16 == 0, "LevelTable layout");

// executed by the first kIn/2 threads of the CTA (followed by a CTA barrier at the call site)
__device__ __forceinline__ void stage_level_table(LevelTable& lt, const nsv_grid_meta& m, int tid, uint32_t agg_max, int fast,
                                                  const __half* table, float* g_table, uint32_t ablate = 0u) {
  if (tid < kIn / 2) {
    lt.scale[tid] = m.scale[tid];
    lt.res[tid] = m.res[tid];
    lt.size[tid] = m.size[tid];
    lt.offset[tid] = m.offset[tid];
    lt.hashed[tid] = m.hashed[tid];
    FastLevel f;
    f.scale = m.scale[tid];
    f.sy = m.res[tid];
    f.sz = m.res[tid] * m.res[tid];
    f.last = m.size[tid] - 1u;
    lt.fast[tid] = f;
    lt.tbl[tid] = reinterpret_cast<const __half2*>(table) + m.offset[tid];
    lt.grd[tid] = g_table ? g_table + 2 * (size_t)m.offset[tid] : nullptr;
  }
  if (tid == 0) {
    int nd = 0;
    while (nd < m.n_levels && !m.hashed[nd]) ++nd;
    bool ok = fast != 0 && (m.n_levels & 3) == 0;
    for (int l = nd; l < m.n_levels; ++l) ok = ok && m.hashed[l] && (m.size[l] & (m.size[l] - 1u)) == 0u;
    for (int l = 0; l < nd; ++l) ok = ok && m.size[l] >= 1u && (uint64_t)m.res[l] * m.res[l] * m.res[l] < (1ull << 31);
    lt.n_dense = (uint32_t)nd;
    lt.fast_ok = ok ? 1u : 0u;
    lt.agg_last = agg_max ? agg_max - 1u : 0u;
    lt.ablate = ablate;
  }
}
struct FusedArgs {
  nsv_inr_config cfg;
  const __half* table;
  const __half* mlp;
  const float* axisangle;
  const float* psf_sigma;
  const float* slice_embedding;
  const float* logit_coef;
  const float* log_var_slice;
  int n_slices;
  float* g_table;
  float* g_mlp;
  float* g_axisangle;
  float* g_se;
  float* g_c;
  float* g_lvs;
  float* losses;
  const float* xyz;
  const float* v;
  const int64_t* slice_idx;
  const float* noise;
  uint64_t seed, offset;
  float* v_out;
  int64_t B;
  int S, log2S;
  int64_t off_density, off_sigma, off_bias;
  uint32_t ablate;   // profiling only: see LevelTable::ablate
  long long* timers; // profiling only: 8 per-phase warp-cycle counters (device memory) or NULL
  uint32_t agg_max;  // tuning: warp-level gradient pre-reduction for dense levels with at most this many entries (0: off)
  int fast;          // tuning: 0 forces the generic (branchy) gather / scatter loops
  int smem_levels;   // tcgen05 kernel: the first `smem_levels` (dense) levels of the fp16 table are bulk-copied (TMA engine,
                     // cp.async.bulk) into shared memory once per CTA and gathered from there; -1: as many as fit, 0: none
  uint32_t smem_table_bytes;  // set by the launcher: bytes of that table prefix
  int tc_groups;     // tcgen05 kernel: 128-sample groups per CTA, 2 (512 threads, 128 registers) or 3 (768 threads, 80 registers;
                     // density-only configurations with n_samples <= 128)
  int tile_order;    // tcgen05 kernel: 0 tiles strided over the CTAs (t, t + 2 grid, ...), 1 a contiguous run of tiles per CTA
                     // (consecutive tiles = consecutive pixels of the batch: pairs with Dataset's locality-aware batch order)
};

// ---- Philox4x32-10 + Box-Muller: three N(0,1) per (seed, sample index) ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ void normal3(uint64_t seed, uint64_t idx, float e[3]) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float u0 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f, u3 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
  const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
  float s, c;
  __sincosf(6.283185307179586f * u1, &s, &c);
  e[0] = r0 * c;
  e[1] = r0 * s;
  e[2] = r1 * __cosf(6.283185307179586f * u3);
}

__device__ __forceinline__ LevelGeom level_from(const LevelTable& t, int l) {
  LevelGeom g;
  g.scale = t.scale[l];
  g.res = t.res[l];
  g.size = t.size[l];
  g.offset = t.offset[l];
  g.hashed = t.hashed[l];
  return g;
}
// The 4 (y,z) corner entries of one lane (x-corner fixed): e[q], q = yb + 2 zb, level offset included.
// The hashed / dense decision is warp-uniform and made once per level; corners are derived
// incrementally (hash: two multiplies + XORs; dense: one base index + strides).
__device__ __forceinline__ void corner_entries(const LevelGeom& lv, uint32_t cx, uint32_t gy, uint32_t gz, uint32_t e[4]) {
  if (lv.hashed) {
    const uint32_t hy0 = gy * 2654435761u, hy1 = hy0 + 2654435761u;
    const uint32_t hz0 = gz * 805459861u, hz1 = hz0 + 805459861u;
    e[0] = cx ^ hy0 ^ hz0;
    e[1] = cx ^ hy1 ^ hz0;
    e[2] = cx ^ hy0 ^ hz1;
    e[3] = cx ^ hy1 ^ hz1;
    if ((lv.size & (lv.size - 1)) == 0) {
      const uint32_t mask = lv.size - 1;
#pragma unroll
      for (int q = 0; q < 4; ++q) e[q] = (e[q] & mask) + lv.offset;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) e[q] = e[q] % lv.size + lv.offset;
    }
  } else {
    const uint32_t sy = lv.res, sz = lv.res * lv.res;
    e[0] = cx + gy * sy + gz * sz;
    e[1] = e[0] + sy;
    e[2] = e[0] + sz;
    e[3] = e[2] + sy;
    if (e[3] >= lv.size || e[0] > e[3]) {  // wrap-around only for out-of-box samples (tcnn semantics: mod T_l)
#pragma unroll
      for (int q = 0; q < 4; ++q) e[q] %= lv.size;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) e[q] += lv.offset;
  }
}

// Lane mapping of the whole kernel: a warp owns 16 samples, lane = (sample s = lane >> 1, x-corner
// xb = lane & 1); both lanes of a pair carry the sample's geometry.  The two x-neighbours of a grid
// cell are adjacent table entries (dense levels always, hashed levels whenever g_x is even), so the
// two lanes of a pair hit the same 128-byte line and a warp-wide LDG / RED touches <= 16 lines
// instead of 32 -- the L1 wavefront count, which bounds the gather / scatter phases, halves.  Each
// lane blends / scatters its 4 (y,z) corners; one shfl_xor(1) combines the pair.

// ---- phase 0: encode the warp's 16 samples into their fp16 rows of the shared tile ----
// `store(l, half2)` parks features (2l, 2l+1) of this lane's sample; called by the xb == 0 lane for l < n_levels
// and by both lanes (interleaved) for the zero padding up to kIn/2
template <typename StoreFn>
__device__ __forceinline__ void encode_warp_generic(const float xn[3], const LevelTable& lt, int n_levels, const __half* __restrict__ table,
                                                    StoreFn store) {
  const int lane = threadIdx.x & 31, xb = lane & 1;
#pragma unroll 4
  for (int l = 0; l < n_levels; ++l) {
    const LevelGeom lv = level_from(lt, l);
    uint32_t g[3], e[4];
    float w[3];
    level_pos(xn, lv.scale, g, w);
    corner_entries(lv, g[0] + xb, g[1], g[2], e);
    float2 f[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) f[q] = load_pair(table, e[q]);
    const float wx = xb ? w[0] : 1.f - w[0];
    const float wy0 = wx * (1.f - w[1]), wy1 = wx * w[1];
    const float wq[4] = {wy0 * (1.f - w[2]), wy1 * (1.f - w[2]), wy0 * w[2], wy1 * w[2]};
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      a0 = fmaf(wq[q], f[q].x, a0);
      a1 = fmaf(wq[q], f[q].y, a1);
    }
    a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    if (xb == 0) store(l, __floats2half2_rn(a0, a1));
  }
  // zero the padding columns (the pair splits them)
  for (int l = n_levels + xb; l < kIn / 2; l += 2) store(l, __floats2half2_rn(0.f, 0.f));
}

// ---- phase 3 tail: scatter dL/d(features) of the warp's 16 samples; optionally dL/dx through the grid ----
// gx: dL/dx_normalised of this lane's sample (both lanes of a pair)
// `fetch(l)` returns dL/d(features 2l, 2l+1) of this lane's sample (both lanes of a pair call it)
template <bool kInputGrad, typename FetchFn>
__device__ __forceinline__ void scatter_warp_generic(const float xn[3], const LevelTable& lt, int n_levels, const __half* __restrict__ table,
                                                     FetchFn fetch, float inv_scale, float* __restrict__ g_table, float gx[3]) {
  const int lane = threadIdx.x & 31, xb = lane & 1;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll 4
  for (int l = 0; l < n_levels; ++l) {
    const LevelGeom lv = level_from(lt, l);
    const float2 gq = fetch(l);
    const float g0 = gq.x * inv_scale, g1 = gq.y * inv_scale;
    uint32_t g[3], e[4];
    float w[3];
    level_pos(xn, lv.scale, g, w);
    corner_entries(lv, g[0] + xb, g[1], g[2], e);
    const float wx = xb ? w[0] : 1.f - w[0];
    if (kInputGrad) {
      float2 f[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) f[q] = load_pair(table, e[q]);
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float dot = fmaf(f[q].x, g0, f[q].y * g1);
        const float fy = (q & 1) ? w[1] : 1.f - w[1], fz = (q >> 1) ? w[2] : 1.f - w[2];
        d0 += (xb ? dot : -dot) * fy * fz;
        d1 += ((q & 1) ? dot : -dot) * wx * fz;
        d2 += ((q >> 1) ? dot : -dot) * wx * fy;
      }
      acc[0] = fmaf(lv.scale, d0, acc[0]);
      acc[1] = fmaf(lv.scale, d1, acc[1]);
      acc[2] = fmaf(lv.scale, d2, acc[2]);
    }
    const float wy0 = wx * (1.f - w[1]), wy1 = wx * w[1];
    const float wt[4] = {wy0 * (1.f - w[2]), wy1 * (1.f - w[2]), wy0 * w[2], wy1 * w[2]};
    bool done = false;
    if (lv.size <= kAggMaxSize) {
      // Coarse levels: the 16 samples of a warp (one pixel's PSF cloud) fall into a handful of cells, and a few
      // thousand entries would receive millions of same-address reductions per iteration, which serialise in a
      // few L2 slices (measured: 0.42 ms of a 1.26 ms kernel for the 3 coarsest levels).  So the warp first finds
      // its distinct cells (match.any), butterfly-reduces every cell's 8 corner contributions across the 16 lanes
      // that share an x-corner, and only the cell's leading lane pair issues reductions.
      const uint32_t key = (g[0] & 0x3ffu) | ((g[1] & 0x3ffu) << 10) | ((g[2] & 0x3ffu) << 20);
      const uint32_t peers = __match_any_sync(0xffffffffu, key);
      const int my_leader = __ffs(peers) - 1;  // even lane: the xb == 0 lane of the first sample in this cell
      uint32_t rem = __ballot_sync(0xffffffffu, lane == my_leader);
      if (__popc(rem) <= kAggMaxCells) {
        while (rem) {
          const int ld = __ffs(rem) - 1;
          rem &= rem - 1;
          const bool mine = (my_leader == ld);
          float v[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v[2 * q] = mine ? wt[q] * g0 : 0.f;
            v[2 * q + 1] = mine ? wt[q] * g1 : 0.f;
          }
#pragma unroll
          for (int off = 2; off < 32; off <<= 1)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
          if ((lane & ~1) == ld) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (v[2 * q] != 0.f || v[2 * q + 1] != 0.f) red_add_v2(g_table + 2 * (size_t)e[q], v[2 * q], v[2 * q + 1]);
          }
        }
        done = true;
      }
    }
    if (!done && (g0 != 0.f || g1 != 0.f)) {
#pragma unroll
      for (int q = 0; q < 4; ++q) red_add_v2(g_table + 2 * (size_t)e[q], wt[q] * g0, wt[q] * g1);
    }
  }
  if (kInputGrad) {
#pragma unroll
    for (int d = 0; d < 3; ++d) gx[d] = acc[d] + __shfl_xor_sync(0xffffffffu, acc[d], 1);
  }
}


// =====================================================================================================
// Branch-free fast path.  The generic loops above decide dense / hashed / wrap-around per level with
// branches, which stops the compiler from overlapping the table loads of different levels (each level
// exposed a full L2 round trip).  Here levels are processed in aligned chunks of 4 (ND leading dense
// levels, the rest hashed, ND a template parameter): all 4 x 4 corner loads of a chunk are issued before
// the first blend, and the chunk's 8 fp16 features leave as one 16-byte store.  Dense levels speculate
// that no index wraps (true unless the sample lies outside the bounding box), clamp the load address and
// raise `bad`; a warp with any bad lane redoes the tile with the generic loops, so results are identical.
// =====================================================================================================
__device__ __forceinline__ uint32_t ldg_u32(const __half2* p) { return __ldg(reinterpret_cast<const unsigned int*>(p)); }
// dense levels: `lt.tbl[l]` is a GENERIC address -- global memory, or the CTA's shared-memory copy of a coarse level that
// the tcgen05 kernel staged with cp.async.bulk (the load resolves the window at run time)
__device__ __forceinline__ uint32_t ldx_u32(const __half2* p) { return *reinterpret_cast<const volatile unsigned int*>(p); }
__device__ __forceinline__ float2 h2f2(uint32_t raw) { return __half22float2(*reinterpret_cast<const __half2*>(&raw)); }

// floor + fraction with full-rate instructions only (FRND / F2I run at quarter rate): adding 1.5 * 2^23 with
// round-down leaves floor(p) in the low mantissa bits; exact for |p| < 2^22 (p <= 2^12 here)
__device__ __forceinline__ void fast_pos(const float xn[3], float scale, uint32_t g[3], float w[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float p = fmaf(scale, xn[d], 0.5f);
    const float t = __fadd_rd(p, 12582912.f);
    g[d] = (uint32_t)(__float_as_int(t) - 0x4B400000);
    w[d] = p - (t - 12582912.f);
  }
}

template <bool HASHED>
__device__ __forceinline__ uint32_t fast_entries(const FastLevel& fl, uint32_t cx, uint32_t gy, uint32_t gz, uint32_t e[4]) {
  if (HASHED) {
    const uint32_t hy0 = gy * 2654435761u, hy1 = hy0 + 2654435761u;
    const uint32_t hz0 = gz * 805459861u, hz1 = hz0 + 805459861u;
    e[0] = (cx ^ hy0 ^ hz0) & fl.last;
    e[1] = (cx ^ hy1 ^ hz0) & fl.last;
    e[2] = (cx ^ hy0 ^ hz1) & fl.last;
    e[3] = (cx ^ hy1 ^ hz1) & fl.last;
    return 0u;
  } else {
    e[0] = cx + gy * fl.sy + gz * fl.sz;
    e[1] = e[0] + fl.sy;
    e[2] = e[0] + fl.sz;
    e[3] = e[2] + fl.sy;
    return (uint32_t)(e[3] > fl.last) | (uint32_t)(e[0] > e[3]);
  }
}

__device__ __forceinline__ void corner_weights(const float w[3], int xb, float wq[4]) {
  const float wx = xb ? w[0] : 1.f - w[0];
  const float wy0 = wx * (1.f - w[1]), wy1 = wx * w[1];
  wq[0] = wy0 * (1.f - w[2]);
  wq[1] = wy1 * (1.f - w[2]);
  wq[2] = wy0 * w[2];
  wq[3] = wy1 * w[2];
}

// levels [4c, 4c+4): the first ND dense, the others hashed.  `store8(c, v)`: v = features 8c .. 8c+7 as 8 halves
template <int ND, typename Store8Fn>
__device__ __forceinline__ uint32_t encode_chunk(const float xn[3], const LevelTable& lt, int c, int xb, Store8Fn store8) {
  uint32_t raw[4][4], bad = 0u;
  float w[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const FastLevel fl = lt.fast[4 * c + i];
    const __half2* tl = lt.tbl[4 * c + i];
    uint32_t g[3], e[4];
    fast_pos(xn, fl.scale, g, w[i]);
    if (i < ND) {
      bad |= fast_entries<false>(fl, g[0] + xb, g[1], g[2], e);
#pragma unroll
      for (int q = 0; q < 4; ++q) e[q] = min(e[q], fl.last);
    } else {
      fast_entries<true>(fl, g[0] + xb, g[1], g[2], e);
    }
    if (lt.ablate & 1u) {
#pragma unroll
      for (int q = 0; q < 4; ++q) raw[i][q] = e[q] & 0x03ff03ffu;
    } else if (i < ND) {
#pragma unroll
      for (int q = 0; q < 4; ++q) raw[i][q] = ldx_u32(tl + e[q]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) raw[i][q] = ldg_u32(tl + e[q]);
    }
  }
  uint32_t packed[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float wq[4];
    corner_weights(w[i], xb, wq);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = h2f2(raw[i][q]);
      a0 = fmaf(wq[q], f.x, a0);
      a1 = fmaf(wq[q], f.y, a1);
    }
    a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    const __half2 h = __floats2half2_rn(a0, a1);
    packed[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  if (xb == 0) store8(c, make_uint4(packed[0], packed[1], packed[2], packed[3]));
  return bad;
}

// Encodes the warp's 16 samples; returns true when the warp had to take the generic loops (some sample wraps
// around a dense level, or the grid has no fast view) -- the caller hands the flag to scatter_warp.
// `store(l, half2)` parks features (2l, 2l+1); `store8(c, uint4)` parks features 8c .. 8c+7 (16-byte aligned).
template <typename StoreFn, typename Store8Fn>
__device__ __forceinline__ bool encode_warp(const float xn[3], const LevelTable& lt, int n_levels, const __half* __restrict__ table,
                                            StoreFn store, Store8Fn store8) {
  const int xb = threadIdx.x & 1;
  bool slow = !lt.fast_ok;
  if (!slow) {
    const int nd = (int)lt.n_dense, nc = n_levels >> 2;
    uint32_t bad = 0u;
    for (int c = 0; c < nc; ++c) {
      const int ndc = nd - 4 * c;
      if (ndc >= 4) bad |= encode_chunk<4>(xn, lt, c, xb, store8);
      else if (ndc <= 0) bad |= encode_chunk<0>(xn, lt, c, xb, store8);
      else if (ndc == 1) bad |= encode_chunk<1>(xn, lt, c, xb, store8);
      else if (ndc == 2) bad |= encode_chunk<2>(xn, lt, c, xb, store8);
      else bad |= encode_chunk<3>(xn, lt, c, xb, store8);
    }
    if (xb == 0)
      for (int c = nc; c < kIn / 8; ++c) store8(c, make_uint4(0u, 0u, 0u, 0u));
    slow = __any_sync(0xffffffffu, bad != 0u);
  }
  if (slow) encode_warp_generic(xn, lt, n_levels, table, store);
  return slow;
}

// ---- warp-level pre-reduction of one coarse dense level's gradient ----
// The 16 samples of a warp belong to one pixel's PSF cloud and mostly share a grid cell on the coarsest levels.
// The warp elects its most populated cell; lanes in it reduce their 8 (corner, feature) contributions with a
// reduce-scatter butterfly over the 16 lanes that share an x-corner (8 shuffles instead of the 32 a plain
// butterfly needs) and 16 lanes issue one scalar reduction each; the other lanes scatter directly.
// e[]: this lane's 4 entries inside the level, p[2q+f] = weight_q * grad_f, gl: the level's gradient base.
__device__ __forceinline__ void scatter_aggregated(const FastLevel& fl, const uint32_t e[4], const float p[8], float* __restrict__ gl,
                                                   int lane, int xb) {
  const uint32_t key = e[0] - (uint32_t)xb;  // entry of the cell's (0,0,0) corner: identical for both lanes of a pair
  const uint32_t peers = __match_any_sync(0xffffffffu, key);
  const uint32_t mine_leader = (uint32_t)(__ffs(peers) - 1);
  const uint32_t best = __reduce_max_sync(0xffffffffu, ((uint32_t)__popc(peers) << 8) | mine_leader);
  const bool mine = mine_leader == (best & 0xffu);
  const bool uniform = (best >> 8) == 32u;
  const uint32_t base = __shfl_sync(0xffffffffu, e[0], (int)(best & 0xffu));  // the elected cell's (0,0,0) corner (leader has xb == 0)
  float r4[4], r2[2], r1;
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lo_v = mine ? p[i] : 0.f, hi_v = mine ? p[i + 4] : 0.f;
      const float send = hi ? lo_v : hi_v, keep = hi ? hi_v : lo_v;
      r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = hi ? r4[i] : r4[i + 2], keep = hi ? r4[i + 2] : r4[i];
      r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
    const float send = hi ? r2[0] : r2[1], keep = hi ? r2[1] : r2[0];
    r1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
  if ((lane & 2) == 0) {
    // this lane now owns corner q = (bit4 -> z, bit3 -> y) and feature f = bit2 of the elected cell, x-corner xb
    const uint32_t ent = base + (uint32_t)xb + ((lane & 8) ? fl.sy : 0u) + ((lane & 16) ? fl.sz : 0u);
    if (r1 != 0.f) red_add(gl + 2 * (size_t)ent + ((lane >> 2) & 1), r1);
  }
  if (!uniform && !mine) {
#pragma unroll
    for (int q = 0; q < 4; ++q) red_add_v2(gl + 2 * (size_t)e[q], p[2 * q], p[2 * q + 1]);
  }
}

// `fetch(l)`: dL/d(features 2l, 2l+1) of this lane's sample (both lanes of a pair call it)
template <int ND, bool kInputGrad, typename FetchFn>
__device__ __forceinline__ void scatter_chunk(const float xn[3], const LevelTable& lt, int c, FetchFn fetch, float inv_scale, float acc[3],
                                              int lane, int xb) {
  uint32_t raw[kInputGrad ? 4 : 1][4], e[4][4];
  float w[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const FastLevel fl = lt.fast[4 * c + i];
    uint32_t g[3];
    fast_pos(xn, fl.scale, g, w[i]);
    if (i < ND) fast_entries<false>(fl, g[0] + xb, g[1], g[2], e[i]);  // the warp is known not to wrap (flag from encode_warp)
    else fast_entries<true>(fl, g[0] + xb, g[1], g[2], e[i]);
    if (kInputGrad) {
      const __half2* tl = lt.tbl[4 * c + i];
#pragma unroll
      for (int q = 0; q < 4; ++q) raw[i][q] = i < ND ? ldx_u32(tl + e[i][q]) : ldg_u32(tl + e[i][q]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int l = 4 * c + i;
    const float2 gq = fetch(l);
    const float g0 = gq.x * inv_scale, g1 = gq.y * inv_scale;
    if (kInputGrad) {
      const float wx = xb ? w[i][0] : 1.f - w[i][0];
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = h2f2(raw[i][q]);
        const float dot = fmaf(f.x, g0, f.y * g1);
        const float fy = (q & 1) ? w[i][1] : 1.f - w[i][1], fz = (q >> 1) ? w[i][2] : 1.f - w[i][2];
        d0 += (xb ? dot : -dot) * fy * fz;
        d1 += ((q & 1) ? dot : -dot) * wx * fz;
        d2 += ((q >> 1) ? dot : -dot) * wx * fy;
      }
      const float sc = lt.fast[l].scale;
      acc[0] = fmaf(sc, d0, acc[0]);
      acc[1] = fmaf(sc, d1, acc[1]);
      acc[2] = fmaf(sc, d2, acc[2]);
    }
    float wt[4], p[8];
    corner_weights(w[i], xb, wt);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      p[2 * q] = wt[q] * g0;
      p[2 * q + 1] = wt[q] * g1;
    }
    float* gl = lt.grd[l];
    if (lt.ablate & 2u) continue;
    if (i < ND && lt.fast[l].last <= lt.agg_last) {
      scatter_aggregated(lt.fast[l], e[i], p, gl, lane, xb);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) red_add_v2(gl + 2 * (size_t)e[i][q], p[2 * q], p[2 * q + 1]);
    }
  }
}

template <bool kInputGrad, typename FetchFn>
__device__ __forceinline__ void scatter_warp(const float xn[3], const LevelTable& lt, int n_levels, const __half* __restrict__ table,
                                             FetchFn fetch, float inv_scale, float* __restrict__ g_table, float gx[3], bool slow) {
  if (slow) {
    scatter_warp_generic<kInputGrad>(xn, lt, n_levels, table, fetch, inv_scale, g_table, gx);
    return;
  }
  const int lane = threadIdx.x & 31, xb = lane & 1;
  float acc[3] = {0.f, 0.f, 0.f};
  const int nd = (int)lt.n_dense, nc = n_levels >> 2;
  for (int c = 0; c < nc; ++c) {
    const int ndc = nd - 4 * c;
    if (ndc >= 4) scatter_chunk<4, kInputGrad>(xn, lt, c, fetch, inv_scale, acc, lane, xb);
    else if (ndc <= 0) scatter_chunk<0, kInputGrad>(xn, lt, c, fetch, inv_scale, acc, lane, xb);
    else if (ndc == 1) scatter_chunk<1, kInputGrad>(xn, lt, c, fetch, inv_scale, acc, lane, xb);
    else if (ndc == 2) scatter_chunk<2, kInputGrad>(xn, lt, c, fetch, inv_scale, acc, lane, xb);
    else scatter_chunk<3, kInputGrad>(xn, lt, c, fetch, inv_scale, acc, lane, xb);
  }
  if (kInputGrad) {
#pragma unroll
    for (int d = 0; d < 3; ++d) gx[d] = acc[d] + __shfl_xor_sync(0xffffffffu, acc[d], 1);
  }
}

__device__ __forceinline__ float softplus_f(float z) { return z > 20.f ? z : log1pf(expf(z)); }
__device__ __forceinline__ float sigmoid_f(float z) { return 1.f / (1.f + expf(-z)); }

__device__ __forceinline__ void group_barrier(int grp, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void red_shared(float* p, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v) : "memory");
}

// softmax chain rule for the slice scale + final loss values (1 block)
static __global__ void __launch_bounds__(256) inr_finalize_kernel(const float* __restrict__ logit_coef, float* __restrict__ g_c,
                                                           float* __restrict__ losses, int n_slices, int slice_scale, int image_reg,
                                                           float delta, int n_levels_bias = 0) {
  __shared__ float red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (slice_scale) {
    float mx = -INFINITY;
    for (int k = tid; k < n_slices; k += 256) mx = fmaxf(mx, logit_coef[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int k = 1; k < 8; ++k) mx = fmaxf(mx, red[k]);
    __syncthreads();
    float se = 0.f;
    for (int k = tid; k < n_slices; k += 256) se += expf(logit_coef[k] - mx);
    se = warp_sum(se);
    if (lane == 0) red[warp] = se;
    __syncthreads();
    se = 0.f;
    for (int k = 0; k < 8; ++k) se += red[k];
    __syncthreads();
    const float lse = mx + logf(se);
    float dot = 0.f;  // sum_k gc_k c_k
    for (int k = tid; k < n_slices; k += 256) dot += g_c[k] * (float)n_slices * expf(logit_coef[k] - lse);
    dot = warp_sum(dot);
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    dot = 0.f;
    for (int k = 0; k < 8; ++k) dot += red[k];
    for (int k = tid; k < n_slices; k += 256) {
      const float c = (float)n_slices * expf(logit_coef[k] - lse);
      g_c[k] = c * (g_c[k] - dot / (float)n_slices);  // in place: dL/dlogit_k
    }
  }
  if (tid == 0 && image_reg == 2) losses[3] = delta * (losses[3] - 1.f);
  if (tid == 0 && n_levels_bias) losses[2] = losses[4] * losses[4];  // biasReg = mean(log_bias)^2 (models.py:323); [4] = the batch mean
  if (tid == 0) losses[6] = losses[0] + losses[1];  // "MSE+logVar" as the reference logs it (models.py:317-319)
}


// implemented in inr_bias.cu: adds mean(log_bias) over the B x S samples to *out_mean (forward only, CUDA cores)
int launch_bias_mean(const FusedArgs& a, float* out_mean, cudaStream_t st);
// implemented in inr_fused_tc.cu; returns NSV_EUNSUPPORTED when the configuration has no tcgen05 instantiation
int launch_train_tc(const FusedArgs& a, cudaStream_t st);
// implemented in inr_fused_ws.cu (warp-specialised tcgen05 kernel); same contract
int launch_train_ws(const FusedArgs& a, cudaStream_t st);

}  // namespace fused
}  // namespace nsv
