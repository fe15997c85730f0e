// pose.cuh -- rigid-pose math shared by the standalone converters and the fused INR kernel.
//
// Semantics follow the reference's converters (nesvor/transform/transform_convert_cuda_kernel.cu:
// axisangle2mat fwd :15-65, bwd :69-188, mat2axisangle fwd :191-264, bwd :267-440): Rodrigues with
// a first-order branch for theta^2 <= 1e-6, four-branch rotation->quaternion, and -- a quirk kept on
// purpose -- *single-precision* libm calls (sqrtf/sinf/cosf/atan2f) whatever the scalar type.
#pragma once
#include "nsv_common.cuh"

namespace nsv {

constexpr double kPoseEps = 1e-6;

// R (row-major 3x3) from a rotation vector w.  R_ij = c d_ij + (1-c) u_i u_j - eps_ijk u_k s.
template <typename T>
__device__ __forceinline__ void rodrigues(const T w[3], T R[9]) {
  const T t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (t2 > kPoseEps) {
    const T th = sqrtf(t2);
    const T u0 = w[0] / th, u1 = w[1] / th, u2 = w[2] / th;
    const T s = sinf(th), c = cosf(th), oc = 1 - c;
    R[0] = c + u0 * u0 * oc;       R[1] = u0 * u1 * oc - u2 * s;  R[2] = u1 * s + u0 * u2 * oc;
    R[3] = u2 * s + u0 * u1 * oc;  R[4] = c + u1 * u1 * oc;       R[5] = -u0 * s + u1 * u2 * oc;
    R[6] = -u1 * s + u0 * u2 * oc; R[7] = u0 * s + u1 * u2 * oc;  R[8] = c + u2 * u2 * oc;
  } else {
    R[0] = 1;     R[1] = -w[2]; R[2] = w[1];
    R[3] = w[2];  R[4] = 1;     R[5] = -w[0];
    R[6] = -w[1]; R[7] = w[0];  R[8] = 1;
  }
}

// Vector-Jacobian product of rodrigues(): G = dL/dR (row-major 3x3) -> gw = dL/dw.
template <typename T>
__device__ __forceinline__ void rodrigues_vjp(const T w[3], const T G[9], T gw[3]) {
  const T t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (t2 > kPoseEps) {
    const T th = sqrtf(t2);
    T u[3] = {w[0] / th, w[1] / th, w[2] / th};
    const T s = sinf(th), c = cosf(th), oc = 1 - c;
    T du[3] = {0, 0, 0}, ds = 0, dc = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const T g = G[i * 3 + j];
        if (i == j) {
          dc += (1 - u[i] * u[i]) * g;
          du[i] += 2 * oc * u[i] * g;
        } else {
          const int k = 3 - i - j;
          const bool plus = ((j - i + 3) % 3) != 1;  // sign of the sine term, -eps_ijk
          dc -= u[i < j ? i : j] * u[i < j ? j : i] * g;
          if (plus) { ds += u[k] * g; du[k] += s * g; }
          else      { ds -= u[k] * g; du[k] -= s * g; }
          du[i] += u[j] * oc * g;
          du[j] += u[i] * oc * g;
        }
      }
    const T radial = c * ds - s * dc;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int b = (i == 0) ? 1 : 0, d = (i == 2) ? 1 : 2;
      T g = radial * u[i];
      g += (du[i] * (1 - u[i] * u[i]) - (du[b] * u[b] + du[d] * u[d]) * u[i]) / th;
      gw[i] = g;
    }
  } else {
    gw[0] = G[7] - G[5];
    gw[1] = G[2] - G[6];
    gw[2] = G[3] - G[1];
  }
}

template <typename T>
struct Quat {
  int pivot;  // -1: trace branch; else dominant diagonal index
  T s, w, v[3];
};

template <typename T>
__device__ __forceinline__ Quat<T> quat_from_rot(const T* R /* 3x3, row stride ld */, int ld) {
  Quat<T> q;
#define R_(i, j) R[(i) * ld + (j)]
  const bool d2 = R_(2, 2) < kPoseEps, d0_gt_d1 = R_(0, 0) > R_(1, 1), d0_lt_nd1 = R_(0, 0) < -R_(1, 1);
  if (!d2 && !d0_lt_nd1) q.pivot = -1;
  else if (d2 && d0_gt_d1) q.pivot = 0;
  else if (d2 && !d0_gt_d1) q.pivot = 1;
  else q.pivot = 2;
  if (q.pivot < 0) {
    q.s = 2 * sqrtf(R_(0, 0) + R_(1, 1) + R_(2, 2) + 1);
    q.w = 0.25 * q.s;
    q.v[0] = (R_(2, 1) - R_(1, 2)) / q.s;
    q.v[1] = (R_(0, 2) - R_(2, 0)) / q.s;
    q.v[2] = (R_(1, 0) - R_(0, 1)) / q.s;
  } else {
    const int p = q.pivot, a = (p + 1) % 3, b = (p + 2) % 3;
    const int o0 = (p == 0) ? 1 : 0, o1 = (p == 2) ? 1 : 2;
    q.s = 2 * sqrtf(R_(p, p) - R_(o0, o0) - R_(o1, o1) + 1);
    q.w = (R_(b, a) - R_(a, b)) / q.s;
    q.v[p] = 0.25 * q.s;
    q.v[a] = (R_(p, a) + R_(a, p)) / q.s;
    q.v[b] = (R_(p, b) + R_(b, p)) / q.s;
  }
#undef R_
  return q;
}

}  // namespace nsv
