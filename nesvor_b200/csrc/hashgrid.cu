// hashgrid.cu -- standalone multiresolution hash-grid encoding: forward, parameter backward
// (scatter) and input backward (dL/dx, needed because x depends on the learnable slice poses).
// C-ABI replacement for tcnn.Encoding(HashGrid) behind build_encoding
// (nesvor/nesvor/models.py:22-25; call site :146).  The fused training kernel (inr_fused.cu) shares
// hashgrid.cuh with this file; this op is what the unfused, autograd-composed path uses.
//
// Layout: thread = (sample, level), blockIdx.y = level, so a warp gathers one level for 32
// consecutive samples (coarse levels: the same few L1 lines).  Nothing is stored for backward:
// both backward kernels recompute indices and weights from x.
#include "hashgrid.cuh"

namespace nsv {
namespace {

constexpr int kThreads = 256;

template <typename TT, int F>
__device__ __forceinline__ void load_feat(const TT* __restrict__ table, uint32_t entry, float f[F]) {
  if constexpr (F == 2) {
    const float2 v = load_pair(table, entry);
    f[0] = v.x;
    f[1] = v.y;
  } else {
#pragma unroll
    for (int k = 0; k < F; ++k) f[k] = (float)table[(size_t)entry * F + k];
  }
}

template <typename TT, typename TO, int F>
__global__ void __launch_bounds__(kThreads)
    fwd_kernel(const float* __restrict__ x, const TT* __restrict__ table, const __grid_constant__ nsv_grid_meta meta,
               TO* __restrict__ out, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= N) return;
  const int l = blockIdx.y;
  const LevelGeom lv = level_geom(meta, l);
  const float xi[3] = {x[i * 3], x[i * 3 + 1], x[i * 3 + 2]};
  uint32_t g[3];
  float w[3];
  level_pos(xi, lv.scale, g, w);
  float acc[F];
#pragma unroll
  for (int k = 0; k < F; ++k) acc[k] = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t e = lv.offset + vertex_index(lv, g[0] + (c & 1), g[1] + ((c >> 1) & 1), g[2] + (c >> 2));
    float f[F];
    load_feat<TT, F>(table, e, f);
    const float wt = corner_weight(c, w);
#pragma unroll
    for (int k = 0; k < F; ++k) acc[k] = fmaf(wt, f[k], acc[k]);
  }
  TO* o = out + i * (int64_t)(meta.n_levels * F) + l * F;
#pragma unroll
  for (int k = 0; k < F; ++k) o[k] = (TO)acc[k];
}

template <typename TG, int F>
__global__ void __launch_bounds__(kThreads)
    bwd_params_kernel(const float* __restrict__ x, const TG* __restrict__ grad_out, const __grid_constant__ nsv_grid_meta meta,
                      float* __restrict__ grad_table, float inv_scale, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= N) return;
  const int l = blockIdx.y;
  const LevelGeom lv = level_geom(meta, l);
  const float xi[3] = {x[i * 3], x[i * 3 + 1], x[i * 3 + 2]};
  uint32_t g[3];
  float w[3];
  level_pos(xi, lv.scale, g, w);
  float go[F];
  bool any = false;
  const TG* gp = grad_out + i * (int64_t)(meta.n_levels * F) + l * F;
#pragma unroll
  for (int k = 0; k < F; ++k) {
    go[k] = (float)gp[k] * inv_scale;
    any |= go[k] != 0.f;
  }
  if (!any) return;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t e = lv.offset + vertex_index(lv, g[0] + (c & 1), g[1] + ((c >> 1) & 1), g[2] + (c >> 2));
    const float wt = corner_weight(c, w);
    float* dst = grad_table + (size_t)e * F;
    if constexpr (F == 2) {
      red_add_v2(dst, wt * go[0], wt * go[1]);
    } else {
#pragma unroll
      for (int k = 0; k < F; ++k) red_add(dst + k, wt * go[k]);
    }
  }
}

// dL/dx_d = sum_l scale_l * sum_{corners} sign_d(c) * prod_{d' != d} w_{d'}(c) * <feat(c), grad_out_l>
template <typename TT, typename TG, int F>
__global__ void __launch_bounds__(kThreads)
    bwd_input_kernel(const float* __restrict__ x, const TT* __restrict__ table, const TG* __restrict__ grad_out,
                     const __grid_constant__ nsv_grid_meta meta, float inv_scale, float* __restrict__ grad_x, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= N) return;
  const float xi[3] = {x[i * 3], x[i * 3 + 1], x[i * 3 + 2]};
  float gx[3] = {0.f, 0.f, 0.f};
  for (int l = 0; l < meta.n_levels; ++l) {
    const LevelGeom lv = level_geom(meta, l);
    uint32_t g[3];
    float w[3];
    level_pos(xi, lv.scale, g, w);
    float go[F];
    const TG* gp = grad_out + i * (int64_t)(meta.n_levels * F) + l * F;
#pragma unroll
    for (int k = 0; k < F; ++k) go[k] = (float)gp[k] * inv_scale;
    float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t e = lv.offset + vertex_index(lv, g[0] + (c & 1), g[1] + ((c >> 1) & 1), g[2] + (c >> 2));
      float f[F];
      load_feat<TT, F>(table, e, f);
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < F; ++k) dot = fmaf(f[k], go[k], dot);
      const float fx = (c & 1) ? w[0] : 1.f - w[0], fy = (c & 2) ? w[1] : 1.f - w[1], fz = (c & 4) ? w[2] : 1.f - w[2];
      d[0] += ((c & 1) ? dot : -dot) * fy * fz;
      d[1] += ((c & 2) ? dot : -dot) * fx * fz;
      d[2] += ((c & 4) ? dot : -dot) * fx * fy;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) gx[k] = fmaf(lv.scale, d[k], gx[k]);
  }
  grad_x[i * 3] = gx[0];
  grad_x[i * 3 + 1] = gx[1];
  grad_x[i * 3 + 2] = gx[2];
}

int check_meta(const char* name, const nsv_grid_meta* m) {
  NSV_REQUIRE(m != nullptr, "%s: NULL meta", name);
  NSV_REQUIRE(m->n_levels >= 1 && m->n_levels <= NSV_MAX_LEVELS, "%s: n_levels %d outside [1,%d]", name, m->n_levels, NSV_MAX_LEVELS);
  NSV_REQUIRE(m->n_features == 1 || m->n_features == 2 || m->n_features == 4 || m->n_features == 8,
              "%s: n_features_per_level must be 1, 2, 4 or 8 (got %d)", name, m->n_features);
  return NSV_OK;
}

#define NSV_DISPATCH_F(F_RT, ...)                     \
  switch (F_RT) {                                     \
    case 1: { constexpr int F = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int F = 2; __VA_ARGS__; } break; \
    case 4: { constexpr int F = 4; __VA_ARGS__; } break; \
    default: { constexpr int F = 8; __VA_ARGS__; } break; \
  }

template <typename TT, typename TO>
int run_fwd(const char* name, const float* x, const TT* table, const nsv_grid_meta* m, TO* out, int64_t N, void* stream) {
  if (int e = check_meta(name, m)) return e;
  NSV_REQUIRE(N >= 0 && (N == 0 || (x && table && out)), "%s: bad arguments", name);
  if (N == 0) return NSV_OK;
  const dim3 grid((unsigned)((N + kThreads - 1) / kThreads), m->n_levels);
  NSV_DISPATCH_F(m->n_features, (fwd_kernel<TT, TO, F><<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, table, *m, out, N)));
  return check_launch(name);
}

template <typename TG>
int run_bwd_params(const char* name, const float* x, const TG* go, const nsv_grid_meta* m, float* gt, float inv_scale, int64_t N,
                   void* stream) {
  if (int e = check_meta(name, m)) return e;
  NSV_REQUIRE(N >= 0 && (N == 0 || (x && go && gt)), "%s: bad arguments", name);
  if (N == 0) return NSV_OK;
  const dim3 grid((unsigned)((N + kThreads - 1) / kThreads), m->n_levels);
  NSV_DISPATCH_F(m->n_features, (bwd_params_kernel<TG, F><<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, go, *m, gt, inv_scale, N)));
  return check_launch(name);
}

template <typename TT, typename TG>
int run_bwd_input(const char* name, const float* x, const TT* table, const TG* go, const nsv_grid_meta* m, float inv_scale,
                  float* gx, int64_t N, void* stream) {
  if (int e = check_meta(name, m)) return e;
  NSV_REQUIRE(N >= 0 && (N == 0 || (x && table && go && gx)), "%s: bad arguments", name);
  if (N == 0) return NSV_OK;
  const unsigned grid = (unsigned)((N + kThreads - 1) / kThreads);
  NSV_DISPATCH_F(m->n_features,
                 (bwd_input_kernel<TT, TG, F><<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, table, go, *m, inv_scale, gx, N)));
  return check_launch(name);
}

}  // namespace
}  // namespace nsv

extern "C" int64_t nsv_grid_meta_init(nsv_grid_meta* m, int n_levels, int n_features, int log2_hashmap_size, int base_resolution,
                                      float per_level_scale) {
  if (!m || n_levels < 1 || n_levels > NSV_MAX_LEVELS || log2_hashmap_size < 1 || log2_hashmap_size > 31 || base_resolution < 1) {
    nsv::set_error("nsv_grid_meta_init: bad arguments");
    return NSV_EINVAL;
  }
  m->n_levels = n_levels;
  m->n_features = n_features;
  // double precision, narrowed once: the float libm entry points differ in the last bit between hosts
  const double log2s = log2((double)per_level_scale);
  uint64_t off = 0;
  for (int l = 0; l < NSV_MAX_LEVELS; ++l) {
    if (l >= n_levels) {
      m->scale[l] = 0.f; m->res[l] = 0; m->size[l] = 0; m->hashed[l] = 0; m->offset[l + 1] = (uint32_t)off;
      continue;
    }
    const float scale = (float)(exp2((double)l * log2s) * (double)base_resolution - 1.0);
    const uint32_t res = (uint32_t)ceilf(scale) + 1u;
    const uint64_t cube = (uint64_t)res * res * res;
    uint64_t dense = cube > 0x7fffffffull ? 0x7fffffffull : cube;
    dense = (dense + 7) / 8 * 8;
    const uint64_t cap = 1ull << log2_hashmap_size;
    const uint64_t size = dense < cap ? dense : cap;
    m->scale[l] = scale;
    m->res[l] = res;
    m->size[l] = (uint32_t)size;
    m->hashed[l] = cube > size ? 1u : 0u;
    m->offset[l] = (uint32_t)off;
    off += size;
    m->offset[l + 1] = (uint32_t)off;
    if (off > 0xffffffffull) {
      nsv::set_error("nsv_grid_meta_init: table exceeds 2^32 entries");
      return NSV_EUNSUPPORTED;
    }
  }
  return (int64_t)off;
}

extern "C" int nsv_hashgrid_fwd_f32(const float* x, const float* table, const nsv_grid_meta* m, float* out, int64_t N, void* stream) {
  return nsv::run_fwd<float, float>("nsv_hashgrid_fwd_f32", x, table, m, out, N, stream);
}
extern "C" int nsv_hashgrid_fwd_f16(const float* x, const void* table, const nsv_grid_meta* m, void* out, int64_t N, void* stream) {
  return nsv::run_fwd<__half, __half>("nsv_hashgrid_fwd_f16", x, (const __half*)table, m, (__half*)out, N, stream);
}
extern "C" int nsv_hashgrid_bwd_params_f32(const float* x, const float* go, const nsv_grid_meta* m, float* gt, int64_t N, void* stream) {
  return nsv::run_bwd_params<float>("nsv_hashgrid_bwd_params_f32", x, go, m, gt, 1.f, N, stream);
}
extern "C" int nsv_hashgrid_bwd_params_f16(const float* x, const void* go, const nsv_grid_meta* m, float* gt, float grad_scale,
                                           int64_t N, void* stream) {
  return nsv::run_bwd_params<__half>("nsv_hashgrid_bwd_params_f16", x, (const __half*)go, m, gt, 1.f / grad_scale, N, stream);
}
extern "C" int nsv_hashgrid_bwd_input_f32(const float* x, const float* table, const float* go, const nsv_grid_meta* m, float* gx,
                                          int64_t N, void* stream) {
  return nsv::run_bwd_input<float, float>("nsv_hashgrid_bwd_input_f32", x, table, go, m, 1.f, gx, N, stream);
}
extern "C" int nsv_hashgrid_bwd_input_f16(const float* x, const void* table, const void* go, const nsv_grid_meta* m, float grad_scale,
                                          float* gx, int64_t N, void* stream) {
  return nsv::run_bwd_input<__half, __half>("nsv_hashgrid_bwd_input_f16", x, (const __half*)table, (const __half*)go, m,
                                            1.f / grad_scale, gx, N, stream);
}
