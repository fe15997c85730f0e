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

template <typename T>
__global__ void __launch_bounds__(kThreads) mat2axisangle_fwd_kernel(const T* __restrict__ mat, T* __restrict__ ax, int n) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  T m[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) m[k] = mat[(size_t)i * 12 + k];
  Quat<T> q = quat_from_rot<T>(m, 4);
  if (q.w < 0) {
    q.w = -q.w;
    q.v[0] = -q.v[0]; q.v[1] = -q.v[1]; q.v[2] = -q.v[2];
  }
  const T n2 = q.v[0] * q.v[0] + q.v[1] * q.v[1] + q.v[2] * q.v[2];
  const T si = sqrtf(n2);
  const T theta = 2 * atan2f(si, q.w);
  const T fac = (n2 > kPoseEps) ? (theta / si) : (T)(2.0 / q.w);
  T* o = ax + (size_t)i * 6;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o[k] = q.v[k] * fac;
    o[3 + k] = m[k * 4 + 3];
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    mat2axisangle_bwd_kernel(const T* __restrict__ mat, const T* __restrict__ gax, T* __restrict__ gmat, int n) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  T m[12], ga[6], G[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) { m[k] = mat[(size_t)i * 12 + k]; G[k] = 0; }
#pragma unroll
  for (int k = 0; k < 6; ++k) ga[k] = gax[(size_t)i * 6 + k];
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
  T* o = gmat + (size_t)i * 12;
#pragma unroll
  for (int k = 0; k < 12; ++k) o[k] = G[k];
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

NSV_POSE_API(f32, float)
NSV_POSE_API(f64, double)
