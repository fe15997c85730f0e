// transform_convert.cu -- axis-angle(6) <-> [R|t](3x4) converters, forward + analytic backward.
// C-ABI replacement for nesvor.transform_convert_cuda (nesvor/transform/transform_convert_cuda.cpp:27-69).
// One thread per pose row; rows are tiny (24/48 B) so a thread moves its row with vector loads.
#include "pose.cuh"

namespace nsv {
namespace {

constexpr int kThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads) axisangle2mat_fwd_kernel(const T* __restrict__ ax, T* __restrict__ mat, int n) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  T a[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) a[k] = ax[(size_t)i * 6 + k];
  T R[9];
  rodrigues<T>(a, R);
  T* m = mat + (size_t)i * 12;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    m[r * 4 + 0] = R[r * 3 + 0];
    m[r * 4 + 1] = R[r * 3 + 1];
    m[r * 4 + 2] = R[r * 3 + 2];
    m[r * 4 + 3] = a[3 + r];
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    axisangle2mat_bwd_kernel(const T* __restrict__ gmat, const T* __restrict__ ax, T* __restrict__ gax, int n) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  T w[3], G[9], gw[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = ax[(size_t)i * 6 + k];
  const T* g = gmat + (size_t)i * 12;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) G[r * 3 + c] = g[r * 4 + c];
  rodrigues_vjp<T>(w, G, gw);
  T* o = gax + (size_t)i * 6;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o[k] = gw[k];
    o[3 + k] = g[k * 4 + 3];
  }
}

// one pose row: [R|t] (row-major 3x4) -> axis-angle(6); reference arithmetic of transform_convert_cuda_kernel.cu:191-264
template <typename T>
__device__ __forceinline__ void mat2axisangle_row(const T m[12], T o[6]) {
  Quat<T> q = quat_from_rot<T>(m, 4);
  if (q.w < 0) {
    q.w = -q.w;
    q.v[0] = -q.v[0]; q.v[1] = -q.v[1]; q.v[2] = -q.v[2];
  }
  const T n2 = q.v[0] * q.v[0] + q.v[1] * q.v[1] + q.v[2] * q.v[2];
  const T si = sqrtf(n2);
  const T theta = 2 * atan2f(si, q.w);
  const T fac = (n2 > kPoseEps) ? (theta / si) : (T)(2.0 / q.w);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o[k] = q.v[k] * fac;
    o[3 + k] = m[k * 4 + 3];
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) mat2axisangle_fwd_kernel(const T* __restrict__ mat, T* __restrict__ ax, int n) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  T m[12], o[6];
#pragma unroll
  for (int k = 0; k < 12; ++k) m[k] = mat[(size_t)i * 12 + k];
  mat2axisangle_row<T>(m, o);
#pragma unroll
  for (int k = 0; k < 6; ++k) ax[(size_t)i * 6 + k] = o[k];
}

// vector-Jacobian product of mat2axisangle_row: ga = dL/d(axis-angle row) -> G = dL/d[R|t] (reference :267-440)
template <typename T>
__device__ __forceinline__ void mat2axisangle_row_vjp(const T m[12], const T ga[6], T G[12]) {
#pragma unroll
  for (int k = 0; k < 12; ++k) G[k] = 0;
  Quat<T> q = quat_from_rot<T>(m, 4);
  const bool neg = q.w < 0;
  if (neg) {
    q.w = -q.w;
    q.v[0] = -q.v[0]; q.v[1] = -q.v[1]; q.v[2] = -q.v[2];
  }
  const T n2 = q.v[0] * q.v[0] + q.v[1] * q.v[1] + q.v[2] * q.v[2];
  const T si = sqrtf(n2);
  const T theta = 2 * atan2f(si, q.w);
  T dw = q.v[0] * ga[0] + q.v[1] * ga[1] + q.v[2] * ga[2];
  T dv[3] = {dw, dw, dw};
  T fac, t;
  if (n2 > kPoseEps) {
    fac = theta / si;
    t = 2 / (q.w * q.w + si * si);
    dw *= -t;
    t = (q.w * t - fac) / si;
#pragma unroll
    for (int k = 0; k < 3; ++k) dv[k] *= t * (q.v[k] / si);
  } else {
    fac = (T)(2.0 / q.w);
    t = 2 / (q.w * q.w + si * si);
    dw *= -t;
    t = (T)((q.w * t - fac) / (si + kPoseEps));
#pragma unroll
    for (int k = 0; k < 3; ++k) dv[k] = (T)(dv[k] * (t * (q.v[k] / (si + kPoseEps))));
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) dv[k] += fac * ga[k];
  if (neg) {
    q.w = -q.w; dw = -dw;
#pragma unroll
    for (int k = 0; k < 3; ++k) { q.v[k] = -q.v[k]; dv[k] = -dv[k]; }
  }
  const T s = q.s;
#define G_(r, c) G[(r) * 4 + (c)]
  if (q.pivot < 0) {
    G_(2, 1) = dv[0] / s; G_(1, 2) = -dv[0] / s;
    G_(0, 2) = dv[1] / s; G_(2, 0) = -dv[1] / s;
    G_(1, 0) = dv[2] / s; G_(0, 1) = -dv[2] / s;
    T ds = (T)(-(q.v[0] * dv[0] + q.v[1] * dv[1] + q.v[2] * dv[2]) / s + 0.25 * dw);
    ds *= 2 / s;
    G_(0, 0) = ds; G_(1, 1) = ds; G_(2, 2) = ds;
  } else {
    const int p = q.pivot, a = (p + 1) % 3, b = (p + 2) % 3;
    G_(b, a) = dw / s;
    G_(a, b) = -dw / s;
    G_(p, a) = dv[a] / s; G_(a, p) = dv[a] / s;
    G_(p, b) = dv[b] / s; G_(b, p) = dv[b] / s;
    T term[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) term[k] = (k == p) ? q.w * dw : q.v[k] * dv[k];
    T ds = (T)(-(term[0] + term[1] + term[2]) / s + 0.25 * dv[p]);
    ds *= 2 / s;
#pragma unroll
    for (int k = 0; k < 3; ++k) G_(k, k) = (k == p) ? ds : -ds;
  }
#undef G_
#pragma unroll
  for (int k = 0; k < 3; ++k) G[k * 4 + 3] = ga[3 + k];
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    mat2axisangle_bwd_kernel(const T* __restrict__ mat, const T* __restrict__ gax, T* __restrict__ gmat, int n) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  T m[12], ga[6], G[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) m[k] = mat[(size_t)i * 12 + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) ga[k] = gax[(size_t)i * 6 + k];
  mat2axisangle_row_vjp<T>(m, ga, G);
  T* o = gmat + (size_t)i * 12;
#pragma unroll
  for (int k = 0; k < 12; ++k) o[k] = G[k];
}

// transReg of the INR training loop (nesvor/nesvor/models.py:357-363) and its gradient in one launch:
//   err = axisangle(T0^-1 o T) per slice (trans_first), loss = mean(err_R^2) + 1e-3 mean(err_T^2) over all slices;
// the reference composes it from inv / compose / mat2axisangle / autograd (~30 small launches per iteration, batch
// independent).  Thread = slice: [R|t] = a2m(ax), [R0|t0] = a2m(ax0); T0^-1 = [R0^T | -R0 t0];
// T0^-1 o T = [R0^T R | t - R^T R0 t0] (transform.py:96-107); then the converters' own row functions, chained by hand.
// grad_ax (+= weight * dloss/dax) and loss (+= loss, unweighted) are accumulated.
template <typename T>
__global__ void __launch_bounds__(kThreads) trans_reg_kernel(const T* __restrict__ ax, const T* __restrict__ ax0, T* __restrict__ gax,
                                                             T* __restrict__ loss, T weight, int n) {
  __shared__ T red[kThreads / 32];
  const int i = blockIdx.x * kThreads + threadIdx.x;
  T li = 0;
  if (i < n) {
    T a[6], a0[6], R[9], R0[9], m[12], err[6], ga[6], G[12];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      a[k] = ax[(size_t)i * 6 + k];
      a0[k] = ax0[(size_t)i * 6 + k];
    }
    rodrigues<T>(a, R);
    rodrigues<T>(a0, R0);
    T u[3];  // R0 t0 (= -t1 of the inverse)
#pragma unroll
    for (int r = 0; r < 3; ++r) u[r] = R0[r * 3] * a0[3] + R0[r * 3 + 1] * a0[4] + R0[r * 3 + 2] * a0[5];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) m[r * 4 + c] = R0[r] * R[c] + R0[3 + r] * R[3 + c] + R0[6 + r] * R[6 + c];  // (R0^T R)_rc
      m[r * 4 + 3] = a[3 + r] - (R[r] * u[0] + R[3 + r] * u[1] + R[6 + r] * u[2]);                              // t - R^T u
    }
    mat2axisangle_row<T>(m, err);
    const T inv = (T)1 / (T)(3 * (int64_t)n);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      li += err[k] * err[k] * inv + (T)1e-3 * err[3 + k] * err[3 + k] * inv;
      ga[k] = 2 * err[k] * inv;
      ga[3 + k] = (T)2e-3 * err[3 + k] * inv;
    }
    mat2axisangle_row_vjp<T>(m, ga, G);
    // dL/dR = R0 dL/dRc  -  u (dL/dtc)^T   (tc_i = t_i - sum_j R_ji u_j),   dL/dt = dL/dtc
    T GR[9], gw[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        GR[r * 3 + c] = R0[r * 3] * G[c] + R0[r * 3 + 1] * G[4 + c] + R0[r * 3 + 2] * G[8 + c] - u[r] * G[c * 4 + 3];
    rodrigues_vjp<T>(a, GR, gw);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      gax[(size_t)i * 6 + k] += weight * gw[k];
      gax[(size_t)i * 6 + 3 + k] += weight * G[k * 4 + 3];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) li += __shfl_xor_sync(0xffffffffu, li, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = li;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) s += red[k];
    atomicAdd(loss, s);
  }
}

template <typename K, typename... Args>
int launch_rows(const char* name, K kernel, int n, void* stream, Args... args) {
  NSV_REQUIRE(n >= 0, "%s: n must be >= 0 (got %d)", name, n);
  if (n == 0) return NSV_OK;
  kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(args..., n);
  return check_launch(name);
}

}  // namespace
}  // namespace nsv

#define NSV_POSE_API(SUF, T)                                                                                     \
  extern "C" int nsv_axisangle2mat_fwd_##SUF(const T* ax, T* mat, int n, void* stream) {                       \
    NSV_REQUIRE(n == 0 || (ax && mat), "nsv_axisangle2mat_fwd: NULL pointer");                                  \
    return nsv::launch_rows("nsv_axisangle2mat_fwd", nsv::axisangle2mat_fwd_kernel<T>, n, stream, ax, mat);     \
  }                                                                                                              \
  extern "C" int nsv_axisangle2mat_bwd_##SUF(const T* gm, const T* ax, T* gax, int n, void* stream) {          \
    NSV_REQUIRE(n == 0 || (gm && ax && gax), "nsv_axisangle2mat_bwd: NULL pointer");                            \
    return nsv::launch_rows("nsv_axisangle2mat_bwd", nsv::axisangle2mat_bwd_kernel<T>, n, stream, gm, ax, gax); \
  }                                                                                                              \
  extern "C" int nsv_mat2axisangle_fwd_##SUF(const T* mat, T* ax, int n, void* stream) {                       \
    NSV_REQUIRE(n == 0 || (ax && mat), "nsv_mat2axisangle_fwd: NULL pointer");                                  \
    return nsv::launch_rows("nsv_mat2axisangle_fwd", nsv::mat2axisangle_fwd_kernel<T>, n, stream, mat, ax);     \
  }                                                                                                              \
  extern "C" int nsv_mat2axisangle_bwd_##SUF(const T* mat, const T* gax, T* gm, int n, void* stream) {         \
    NSV_REQUIRE(n == 0 || (mat && gax && gm), "nsv_mat2axisangle_bwd: NULL pointer");                           \
    return nsv::launch_rows("nsv_mat2axisangle_bwd", nsv::mat2axisangle_bwd_kernel<T>, n, stream, mat, gax, gm); \
  }

extern "C" int nsv_trans_reg_f32(const float* axisangle, const float* axisangle_init, float* grad_axisangle, float* loss, int n,
                                 float weight, void* stream) {
  NSV_REQUIRE(n >= 0, "nsv_trans_reg: n must be >= 0 (got %d)", n);
  if (n == 0) return NSV_OK;
  NSV_REQUIRE(axisangle && axisangle_init && grad_axisangle && loss, "nsv_trans_reg: NULL pointer");
  nsv::trans_reg_kernel<float><<<(n + nsv::kThreads - 1) / nsv::kThreads, nsv::kThreads, 0, (cudaStream_t)stream>>>(
      axisangle, axisangle_init, grad_axisangle, loss, weight, n);
  return nsv::check_launch("nsv_trans_reg");
}

NSV_POSE_API(f32, float)
NSV_POSE_API(f64, double)
