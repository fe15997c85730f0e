/*
 * TEST INFRASTRUCTURE ONLY -- see oracle/nsv_oracle.c.  Included twice (REAL = float, double).
 *
 * CPU restatement of the reference's rigid-pose converters
 *   axisangle2mat forward   /root/reference/nesvor/transform/transform_convert_cuda_kernel.cu:15-65
 *   axisangle2mat backward  ...:69-188
 *   mat2axisangle forward   ...:191-264
 *   mat2axisangle backward  ...:267-440
 * Rows are [rx ry rz tx ty tz] <-> 3x4 [R | t] row-major.  The reference calls the *float* libm
 * entry points (sqrtf/sinf/cosf/atan2f) whatever scalar_t is; so does this file.  Loops over
 * matrix entries / quaternion pivots replace the reference's unrolled statements, but every
 * accumulator receives its terms in the same order, so -ffp-contract=off builds agree bit for bit
 * with oracle/_ref.
 */

#define NSV_CAT_(a, b) a##b
#define NSV_CAT(a, b) NSV_CAT_(a, b)
#define FN(name) NSV_CAT(name, SUFFIX)

#define NSV_POSE_EPS 1e-6 /* TRANSFORM_EPS, a double literal in the reference too */

/* R_ij = c d_ij + (1-c) u_i u_j - eps_ijk u_k s ;  returns -eps_ijk and writes k */
static int FN(skew_sign)(int i, int j, int* k) {
  *k = 3 - i - j;
  return ((j - i + 3) % 3 == 1) ? -1 : 1;
}

void FN(nsv_oracle_axisangle2mat_forward_)(const REAL* axisangle, REAL* mat, int n) {
  for (int row = 0; row < n; ++row) {
    const REAL* a = axisangle + row * 6;
    REAL* m = mat + row * 12;
    REAL u[3] = {a[0], a[1], a[2]};
    const REAL theta2 = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    if (theta2 > NSV_POSE_EPS) {
      const REAL theta = sqrtf(theta2);
      for (int i = 0; i < 3; ++i) u[i] /= theta;
      const REAL s = sinf(theta), c = cosf(theta), oc = 1 - c;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          if (i == j) {
            m[i * 4 + j] = c + u[i] * u[i] * oc;
          } else {
            int k;
            const int sg = FN(skew_sign)(i, j, &k);
            /* the reference writes the sine term first when it is positive or leads the line */
            const REAL lo = (i < j) ? u[i] : u[j], hi = (i < j) ? u[j] : u[i];
            const REAL sym = lo * hi * oc;
            if (sg > 0)
              m[i * 4 + j] = (i == 0 && j == 2) || (i == 1 && j == 0) || (i == 2 && j == 1)
                                 ? u[k] * s + sym
                                 : sym + u[k] * s;
            else
              m[i * 4 + j] = (i == 0 && j == 1) ? sym - u[k] * s : -u[k] * s + sym;
          }
        }
    } else { /* small angle: I + [w]x */
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          if (i == j) {
            m[i * 4 + j] = 1;
          } else {
            int k;
            const int sg = FN(skew_sign)(i, j, &k);
            m[i * 4 + j] = sg > 0 ? u[k] : -u[k];
          }
        }
    }
    for (int i = 0; i < 3; ++i) m[i * 4 + 3] = a[3 + i];
  }
}

void FN(nsv_oracle_axisangle2mat_backward_)(const REAL* grad_mat, const REAL* axisangle,
                                            REAL* grad_axisangle, int n) {
  for (int row = 0; row < n; ++row) {
    const REAL* a = axisangle + row * 6;
    const REAL* G = grad_mat + row * 12;
    REAL* ga = grad_axisangle + row * 6;
    REAL u[3] = {a[0], a[1], a[2]};
    const REAL theta2 = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    if (theta2 > NSV_POSE_EPS) {
      const REAL theta = sqrtf(theta2);
      for (int i = 0; i < 3; ++i) u[i] /= theta;
      const REAL s = sinf(theta), c = cosf(theta), oc = 1 - c;
      REAL du[3] = {0, 0, 0}, ds = 0, dc = 0;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          const REAL g = G[i * 4 + j];
          if (i == j) {
            dc += (1 - u[i] * u[i]) * g;
            du[i] += 2 * oc * u[i] * g;
          } else {
            int k;
            const int sg = FN(skew_sign)(i, j, &k);
            const REAL lo = (i < j) ? u[i] : u[j], hi = (i < j) ? u[j] : u[i];
            dc -= lo * hi * g;
            if (sg > 0) { ds += u[k] * g; du[k] += s * g; }
            else        { ds -= u[k] * g; du[k] -= s * g; }
            du[i] += u[j] * oc * g;
            du[j] += u[i] * oc * g;
          }
        }
      for (int i = 0; i < 3; ++i) {
        const int b = (i == 0) ? 1 : 0, d = (i == 2) ? 1 : 2; /* the two other axes, ascending */
        REAL g = (c * ds - s * dc) * u[i];
        g += (du[i] * (1 - u[i] * u[i]) - (du[b] * u[b] + du[d] * u[d]) * u[i]) / theta;
        ga[i] = g;
      }
    } else {
      ga[0] = G[9] - G[6];
      ga[1] = G[2] - G[8];
      ga[2] = G[4] - G[1];
    }
    for (int i = 0; i < 3; ++i) ga[3 + i] = G[i * 4 + 3];
  }
}

/* rotation -> unit quaternion (w, v) with the reference's branch selection; pivot = -1 is the
 * trace branch, otherwise the index of the dominant diagonal element. */
typedef struct {
  int pivot;
  REAL s, w, v[3];
} FN(Quat);

static void FN(quat_from_rot)(const REAL* m, FN(Quat) * q) {
  const REAL r00 = m[0], r11 = m[5], r22 = m[10];
  const int d2 = r22 < NSV_POSE_EPS, d0_gt_d1 = r00 > r11, d0_lt_nd1 = r00 < -r11;
  if (!d2 && !d0_lt_nd1) q->pivot = -1;
  else if (d2 && d0_gt_d1) q->pivot = 0;
  else if (d2 && !d0_gt_d1) q->pivot = 1;
  else q->pivot = 2;
#define R_(i, j) m[(i) * 4 + (j)]
  if (q->pivot < 0) {
    q->s = 2 * sqrtf(r00 + r11 + r22 + 1);
    q->w = (REAL)(0.25 * q->s);
    q->v[0] = (R_(2, 1) - R_(1, 2)) / q->s;
    q->v[1] = (R_(0, 2) - R_(2, 0)) / q->s;
    q->v[2] = (R_(1, 0) - R_(0, 1)) / q->s;
  } else {
    const int p = q->pivot, a = (p + 1) % 3, b = (p + 2) % 3;
    const int o0 = (p == 0) ? 1 : 0, o1 = (p == 2) ? 1 : 2; /* other diagonal entries, ascending */
    q->s = 2 * sqrtf(R_(p, p) - R_(o0, o0) - R_(o1, o1) + 1);
    q->w = (R_(b, a) - R_(a, b)) / q->s;
    q->v[p] = (REAL)(0.25 * q->s);
    q->v[a] = (R_(p, a) + R_(a, p)) / q->s;
    q->v[b] = (R_(p, b) + R_(b, p)) / q->s;
  }
#undef R_
}

void FN(nsv_oracle_mat2axisangle_forward_)(const REAL* mat, REAL* axisangle, int n) {
  for (int row = 0; row < n; ++row) {
    const REAL* m = mat + row * 12;
    REAL* a = axisangle + row * 6;
    FN(Quat) q;
    FN(quat_from_rot)(m, &q);
    if (q.w < 0) {
      q.w *= -1;
      for (int i = 0; i < 3; ++i) q.v[i] *= -1;
    }
    const REAL n2 = q.v[0] * q.v[0] + q.v[1] * q.v[1] + q.v[2] * q.v[2];
    const REAL si = sqrtf(n2);
    const REAL theta = 2 * atan2f(si, q.w);
    const REAL fac = (n2 > NSV_POSE_EPS) ? (theta / si) : (REAL)(2.0 / q.w);
    for (int i = 0; i < 3; ++i) {
      a[i] = q.v[i] * fac;
      a[3 + i] = m[i * 4 + 3];
    }
  }
}

void FN(nsv_oracle_mat2axisangle_backward_)(const REAL* mat, const REAL* grad_axisangle,
                                            REAL* grad_mat, int n) {
  for (int row = 0; row < n; ++row) {
    const REAL* m = mat + row * 12;
    const REAL* ga = grad_axisangle + row * 6;
    REAL* G = grad_mat + row * 12;
    for (int k = 0; k < 12; ++k) G[k] = 0;
    FN(Quat) q;
    FN(quat_from_rot)(m, &q);
    const int neg = q.w < 0;
    if (neg) {
      q.w *= -1;
      for (int i = 0; i < 3; ++i) q.v[i] *= -1;
    }
    REAL n2 = q.v[0] * q.v[0] + q.v[1] * q.v[1] + q.v[2] * q.v[2];
    const REAL si = sqrtf(n2);
    const REAL theta = 2 * atan2f(si, q.w);
    REAL dw = q.v[0] * ga[0] + q.v[1] * ga[1] + q.v[2] * ga[2];
    REAL dv[3] = {dw, dw, dw};
    REAL fac, t;
    if (n2 > NSV_POSE_EPS) {
      fac = theta / si;
      t = 2 / (q.w * q.w + si * si);
      dw *= -t;
      t = (q.w * t - fac) / si;
      for (int i = 0; i < 3; ++i) dv[i] *= t * (q.v[i] / si);
    } else {
      fac = (REAL)(2.0 / q.w);
      t = 2 / (q.w * q.w + si * si);
      dw *= -t;
      t = (REAL)((q.w * t - fac) / (si + NSV_POSE_EPS));
      for (int i = 0; i < 3; ++i) dv[i] = (REAL)(dv[i] * (t * (q.v[i] / (si + NSV_POSE_EPS))));
    }
    for (int i = 0; i < 3; ++i) dv[i] += fac * ga[i];
    if (neg) {
      q.w *= -1;
      dw *= -1;
      for (int i = 0; i < 3; ++i) { q.v[i] *= -1; dv[i] *= -1; }
    }
    const REAL s = q.s;
#define G_(i, j) G[(i) * 4 + (j)]
    if (q.pivot < 0) {
      G_(2, 1) = dv[0] / s; G_(1, 2) = -dv[0] / s;
      G_(0, 2) = dv[1] / s; G_(2, 0) = -dv[1] / s;
      G_(1, 0) = dv[2] / s; G_(0, 1) = -dv[2] / s;
      REAL ds = (REAL)(-(q.v[0] * dv[0] + q.v[1] * dv[1] + q.v[2] * dv[2]) / s + 0.25 * dw);
      ds *= 2 / s;
      G_(0, 0) = ds; G_(1, 1) = ds; G_(2, 2) = ds;
    } else {
      const int p = q.pivot, a = (p + 1) % 3, b = (p + 2) % 3;
      G_(b, a) = dw / s;
      G_(a, b) = -dw / s;
      G_(p, a) = dv[a] / s; G_(a, p) = dv[a] / s;
      G_(p, b) = dv[b] / s; G_(b, p) = dv[b] / s;
      REAL term[3];
      for (int i = 0; i < 3; ++i) term[i] = (i == p) ? q.w * dw : q.v[i] * dv[i];
      REAL ds = (REAL)(-(term[0] + term[1] + term[2]) / s + 0.25 * dv[p]);
      ds *= 2 / s;
      for (int i = 0; i < 3; ++i) G_(i, i) = (i == p) ? ds : -ds;
    }
#undef G_
    for (int i = 0; i < 3; ++i) G[i * 4 + 3] = ga[3 + i];
  }
}

#undef NSV_POSE_EPS
#undef FN
#undef NSV_CAT
#undef NSV_CAT_
