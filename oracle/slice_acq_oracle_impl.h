/*
 * TEST INFRASTRUCTURE ONLY -- see oracle/nsv_oracle.c.  Included twice (REAL = float, double).
 *
 * CPU restatement of the reference's slice-acquisition operator family
 *   forward            /root/reference/nesvor/slice_acquisition/slice_acq_cuda_kernel.cu:18-171
 *   backward           ...:174-470
 *   adjoint forward    ...:473-670
 *   equalize           ...:673-693
 *   adjoint backward   ...:696-950
 * and of what their host wrappers do around them (zero-filled outputs, equalize passes,
 * ...:954-1132).  Written from the algorithm (SURVEY.md App. D + quirks Q1-Q9), organised around a
 * shared "pixel frame -> PSF tap -> 8-corner stencil" walker instead of five unrolled kernels.
 * The floating-point expression order of the reference is kept (sum over taps in z,y,x order,
 * corners in the order 000,100,010,001,110,101,011,111, weights as ((fx*fy)*fz)*psf) so that,
 * compiled with -ffp-contract=off, results agree with oracle/_ref to the last bit on one thread.
 */

#define NSV_CAT_(a, b) a##b
#define NSV_CAT(a, b) NSV_CAT_(a, b)
#define FN(name) NSV_CAT(name, SUFFIX)

/* corner c of the trilinear cell: bit0 -> +x, bit1 -> +y, bit2 -> +z; visiting order of the reference */
static const int FN(kCornerOrder)[8] = {0, 1, 2, 4, 3, 5, 6, 7};

typedef struct {
  REAL r[3][3];  /* rotation rows */
  REAL s[3];     /* pixel position in the slice frame, voxel units (translation added) */
  REAL c[3];     /* pixel centre in volume index space (x->W, y->H, z->D) */
} FN(PixelFrame);

/* slice_acq_cuda_kernel.cu:38-56 (same block opens every pixel kernel) */
static void FN(pixel_frame)(const REAL* tf, int ix, int iy, int h, int w, int D, int H, int W,
                            REAL res_slice, FN(PixelFrame) * f) {
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) f->r[a][b] = tf[a * 4 + b];
  /* Q8: centre offset in double, narrowed on assignment */
  f->s[0] = (REAL)((ix - (w - 1) / 2.) * res_slice + tf[3]);
  f->s[1] = (REAL)((iy - (h - 1) / 2.) * res_slice + tf[7]);
  f->s[2] = tf[11];
  const double half[3] = {(W - 1) / 2., (H - 1) / 2., (D - 1) / 2.};
  for (int a = 0; a < 3; ++a) {
    REAL v = f->r[a][0] * f->s[0] + f->r[a][1] * f->s[1] + f->r[a][2] * f->s[2];
    f->c[a] = (REAL)(v + half[a]);
  }
}

/* position of PSF tap (tx,ty,tz) in volume index space; returns 0 if outside [0, dim-1) (Q5) */
static int FN(tap_position)(const FN(PixelFrame) * f, int tx, int ty, int tz, int D, int H, int W,
                            REAL p[3]) {
  for (int a = 0; a < 3; ++a) p[a] = f->c[a] + f->r[a][0] * tx + f->r[a][1] * ty + f->r[a][2] * tz;
  if (p[0] < 0 || p[1] < 0 || p[2] < 0 || p[0] >= W - 1 || p[1] >= H - 1 || p[2] >= D - 1) return 0;
  return 1;
}

typedef struct {
  int base;     /* flat index of corner 000 */
  int off[8];   /* offset of corner c from base */
  REAL wt[8];   /* trilinear weight of corner c (no PSF factor) */
  REAL fx[2], fy[2], fz[2];
} FN(Cell);

static void FN(cell_at)(const REAL p[3], int sy, int sz, FN(Cell) * cell) {
  const int x0 = (int)floor(p[0]), y0 = (int)floor(p[1]), z0 = (int)floor(p[2]);
  const REAL wx = p[0] - x0, wy = p[1] - y0, wz = p[2] - z0;
  cell->fx[0] = 1 - wx; cell->fx[1] = wx;
  cell->fy[0] = 1 - wy; cell->fy[1] = wy;
  cell->fz[0] = 1 - wz; cell->fz[1] = wz;
  cell->base = z0 * sz + y0 * sy + x0;
  for (int c = 0; c < 8; ++c) {
    const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
    cell->off[c] = bx + by * sy + bz * sz;
    cell->wt[c] = cell->fx[bx] * cell->fy[by] * cell->fz[bz];
  }
}

/* d(trilinear)/d(x,y,z) contribution of corner c carrying value v (sign pattern of .cu:394-449) */
static void FN(cell_grad_accum)(const FN(Cell) * cell, int c, REAL v, REAL d[3]) {
  const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
  const REAL gx = cell->fy[by] * cell->fz[bz] * v;
  const REAL gy = cell->fx[bx] * cell->fz[bz] * v;
  const REAL gz = cell->fx[bx] * cell->fy[by] * v;
  d[0] = bx ? d[0] + gx : d[0] - gx;
  d[1] = by ? d[1] + gy : d[1] - gy;
  d[2] = bz ? d[2] + gz : d[2] - gz;
}

/* "interp_psf" mode (Q9): nearest voxel, PSF resampled trilinearly at the voxel's offset from the
 * pixel centre expressed in the slice frame.  Returns 0 when the resampling point leaves the PSF box. */
typedef struct {
  int vox;         /* flat index of the nearest voxel */
  int rx, ry, rz;  /* its integer coordinates */
  FN(Cell) pc;     /* cell inside the PSF array */
} FN(NearestTap);

static int FN(nearest_tap)(const FN(PixelFrame) * f, const REAL p[3], int sy, int sz, int d_p,
                           int h_p, int w_p, FN(NearestTap) * t) {
  t->rx = (int)round(p[0]); /* Q7: half away from zero */
  t->ry = (int)round(p[1]);
  t->rz = (int)round(p[2]);
  t->vox = t->rz * sz + t->ry * sy + t->rx;
  const REAL dx = t->rx - f->c[0], dy = t->ry - f->c[1], dz = t->rz - f->c[2];
  REAL q[3];
  q[0] = (REAL)(f->r[0][0] * dx + f->r[1][0] * dy + f->r[2][0] * dz + (w_p - 1) / 2.);
  q[1] = (REAL)(f->r[0][1] * dx + f->r[1][1] * dy + f->r[2][1] * dz + (h_p - 1) / 2.);
  q[2] = (REAL)(f->r[0][2] * dx + f->r[1][2] * dy + f->r[2][2] * dz + (d_p - 1) / 2.);
  if (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= w_p - 1 || q[1] >= h_p - 1 || q[2] >= d_p - 1)
    return 0;
  FN(cell_at)(q, w_p, w_p * h_p, &t->pc);
  return 1;
}

static REAL FN(psf_resampled)(const REAL* psf, const FN(Cell) * pc) {
  REAL v = 0;
  for (int k = 0; k < 8; ++k) {
    const int c = FN(kCornerOrder)[k];
    v += pc->wt[c] * psf[pc->base + pc->off[c]];
  }
  return v;
}

#define NSV_TAP_LOOP_BEGIN                                               \
  for (int tz = -d_p / 2, ip = 0; tz < (d_p + 1) / 2; ++tz)              \
    for (int ty = -h_p / 2; ty < (h_p + 1) / 2; ++ty)                    \
      for (int tx = -w_p / 2; tx < (w_p + 1) / 2; ++tx, ++ip) {          \
        REAL tap = psf[ip];                                              \
        if (tap == 0) continue;                                          \
        REAL p[3];                                                       \
        if (!FN(tap_position)(&f, tx, ty, tz, D, H, W, p)) continue;
#define NSV_TAP_LOOP_END }

/* Q3: normalisation weight used by backward / adjoint: in-bounds taps, vol_mask ignored */
static REAL FN(unmasked_weight)(const FN(PixelFrame) * fp, const REAL* psf, int D, int H, int W,
                                int d_p, int h_p, int w_p, int interp_psf) {
  const FN(PixelFrame) f = *fp;
  const int sy = W, sz = H * W;
  REAL weight = 0;
  NSV_TAP_LOOP_BEGIN
  if (interp_psf) {
    FN(NearestTap) t;
    if (!FN(nearest_tap)(&f, p, sy, sz, d_p, h_p, w_p, &t)) continue;
    tap = FN(psf_resampled)(psf, &t.pc);
  }
  weight += tap;
  NSV_TAP_LOOP_END
  return weight;
}

/* ---------------------------------------------------------------- forward: A (gather) */
void FN(nsv_oracle_slice_acq_forward_)(const REAL* transforms, const REAL* vol,
                                       const unsigned char* vol_mask,
                                       const unsigned char* slices_mask, const REAL* psf,
                                       REAL* slices, REAL* slices_weight, int D, int H, int W,
                                       int d_p, int h_p, int w_p, int n, int h, int w,
                                       REAL res_slice, int interp_psf) {
  const int sy = W, sz = H * W;
  const long npx = (long)n * h * w;
  memset(slices, 0, (size_t)npx * sizeof(REAL));
  if (slices_weight) memset(slices_weight, 0, (size_t)npx * sizeof(REAL));
#pragma omp parallel for schedule(static)
  for (long idx = 0; idx < npx; ++idx) {
    if (slices_mask && !slices_mask[idx]) continue;
    const int ix = (int)(idx % w), iy = (int)((idx / w) % h), is = (int)(idx / ((long)h * w));
    FN(PixelFrame) f;
    FN(pixel_frame)(transforms + is * 12, ix, iy, h, w, D, H, W, res_slice, &f);
    REAL val = 0, weight = 0;
    NSV_TAP_LOOP_BEGIN
    if (interp_psf) {
      FN(NearestTap) t;
      t.rx = (int)round(p[0]); t.ry = (int)round(p[1]); t.rz = (int)round(p[2]);
      t.vox = t.rz * sz + t.ry * sy + t.rx;
      if (vol_mask && !vol_mask[t.vox]) continue;
      const REAL v = vol[t.vox];
      if (!FN(nearest_tap)(&f, p, sy, sz, d_p, h_p, w_p, &t)) continue;
      tap = FN(psf_resampled)(psf, &t.pc);
      val += tap * v;
      weight += tap;
    } else {
      FN(Cell) cell;
      FN(cell_at)(p, sy, sz, &cell);
      for (int k = 0; k < 8; ++k) {
        const int c = FN(kCornerOrder)[k];
        const int iv = cell.base + cell.off[c];
        if (vol_mask && !vol_mask[iv]) continue;
        const REAL pw = cell.wt[c] * tap;
        val += pw * vol[iv];
        weight += pw;
      }
    }
    NSV_TAP_LOOP_END
    if (weight > 0) { /* Q1 */
      slices[idx] = val / weight;
      if (slices_weight) slices_weight[idx] = weight;
    }
  }
}

/* rigid-transform gradient bookkeeping shared by backward and adjoint-backward */
typedef struct { REAL g[12]; } FN(TfGrad);

static void FN(tfgrad_linear)(FN(TfGrad) * a, const FN(PixelFrame) * f, const REAL d[3], int tx,
                              int ty, int tz) {
  const REAL q[3] = {f->s[0] + tx, f->s[1] + ty, f->s[2] + tz};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) a->g[r * 4 + c] += d[r] * q[c];
  for (int c = 0; c < 3; ++c)
    a->g[c * 4 + 3] += d[0] * f->r[0][c] + d[1] * f->r[1][c] + d[2] * f->r[2][c];
}

static void FN(tfgrad_nearest)(FN(TfGrad) * a, const REAL d[3], const FN(NearestTap) * t, int D,
                               int H, int W) {
  /* the voxel offset is a double in the reference, so the product and the sum round once */
  const double q[3] = {t->rx - (W - 1) / 2., t->ry - (H - 1) / 2., t->rz - (D - 1) / 2.};
  /* note the transposed pattern of the reference (.cu:370-372): row <- voxel coordinate, col <- d */
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) a->g[r * 4 + c] = (REAL)(a->g[r * 4 + c] + d[c] * q[r]);
  for (int c = 0; c < 3; ++c) a->g[c * 4 + 3] -= d[c];
}

/* ---------------------------------------------------------------- backward of A */
void FN(nsv_oracle_slice_acq_backward_)(const REAL* transforms, const REAL* vol,
                                        const unsigned char* vol_mask, const REAL* psf,
                                        const REAL* grad_slices, const unsigned char* slices_mask,
                                        REAL* grad_vol, REAL* grad_transforms, int D, int H, int W,
                                        int d_p, int h_p, int w_p, int n, int h, int w,
                                        REAL res_slice, int interp_psf) {
  const int sy = W, sz = H * W;
  const long npx = (long)n * h * w;
  if (grad_vol) memset(grad_vol, 0, (size_t)D * H * W * sizeof(REAL));
  if (grad_transforms) memset(grad_transforms, 0, (size_t)n * 12 * sizeof(REAL));
#pragma omp parallel for schedule(static)
  for (long idx = 0; idx < npx; ++idx) { /* scatter: atomics; 1 thread == serial order */
    if (slices_mask && !slices_mask[idx]) continue;
    REAL gs = grad_slices[idx];
    if (gs == 0) continue; /* Q2 */
    const int ix = (int)(idx % w), iy = (int)((idx / w) % h), is = (int)(idx / ((long)h * w));
    FN(PixelFrame) f;
    FN(pixel_frame)(transforms + is * 12, ix, iy, h, w, D, H, W, res_slice, &f);
    const REAL weight = FN(unmasked_weight)(&f, psf, D, H, W, d_p, h_p, w_p, interp_psf);
    if (weight == 0) continue;
    gs /= weight;
    FN(TfGrad) acc;
    memset(&acc, 0, sizeof acc);
    NSV_TAP_LOOP_BEGIN
    if (interp_psf) {
      FN(NearestTap) t;
      if (!FN(nearest_tap)(&f, p, sy, sz, d_p, h_p, w_p, &t)) continue;
      if (vol_mask && !vol_mask[t.vox]) continue;
      if (grad_vol) {
        const REAL add = FN(psf_resampled)(psf, &t.pc) * gs;
#pragma omp atomic
        grad_vol[t.vox] += add;
      }
      if (grad_transforms) {
        REAL d[3] = {0, 0, 0};
        for (int k = 0; k < 8; ++k) {
          const int c = FN(kCornerOrder)[k];
          FN(cell_grad_accum)(&t.pc, c, psf[t.pc.base + t.pc.off[c]], d);
        }
        const REAL sc = gs * vol[t.vox];
        d[0] *= sc; d[1] *= sc; d[2] *= sc;
        FN(tfgrad_nearest)(&acc, d, &t, D, H, W);
      }
    } else {
      FN(Cell) cell;
      FN(cell_at)(p, sy, sz, &cell);
      tap *= gs;
      if (grad_vol)
        for (int k = 0; k < 8; ++k) {
          const int c = FN(kCornerOrder)[k];
          const int iv = cell.base + cell.off[c];
          if (vol_mask && !vol_mask[iv]) continue;
          const REAL add = cell.wt[c] * tap;
#pragma omp atomic
          grad_vol[iv] += add;
        }
      if (grad_transforms) {
        REAL d[3] = {0, 0, 0};
        for (int k = 0; k < 8; ++k) {
          const int c = FN(kCornerOrder)[k];
          const int iv = cell.base + cell.off[c];
          if (vol_mask && !vol_mask[iv]) continue;
          FN(cell_grad_accum)(&cell, c, tap * vol[iv], d);
        }
        FN(tfgrad_linear)(&acc, &f, d, tx, ty, tz);
      }
    }
    NSV_TAP_LOOP_END
    if (grad_transforms)
      for (int k = 0; k < 12; ++k) {
#pragma omp atomic
        grad_transforms[is * 12 + k] += acc.g[k];
      }
  }
}

/* ---------------------------------------------------------------- equalize */
void FN(nsv_oracle_equalize_)(REAL* vol, const REAL* vol_weight, int is_grad, long DHW) {
  for (long i = 0; i < DHW; ++i) {
    const REAL wgt = vol_weight[i];
    if (!(wgt > 0)) continue;
    if (is_grad && wgt < 1e-3)
      vol[i] = (REAL)(vol[i] / 1e-3); /* double division, as `vol[idx] /= 1e-3` */
    else
      vol[i] /= wgt;
  }
}

/* ---------------------------------------------------------------- adjoint forward: A^T (scatter) */
void FN(nsv_oracle_slice_acq_adjoint_forward_)(const REAL* transforms, const REAL* psf,
                                               const REAL* slices,
                                               const unsigned char* slices_mask,
                                               const unsigned char* vol_mask, REAL* vol,
                                               REAL* vol_weight, int D, int H, int W, int d_p,
                                               int h_p, int w_p, int n, int h, int w,
                                               REAL res_slice, int interp_psf, int equalize) {
  const int sy = W, sz = H * W;
  const long npx = (long)n * h * w, nvx = (long)D * H * W;
  memset(vol, 0, (size_t)nvx * sizeof(REAL));
  REAL* vw = equalize ? vol_weight : NULL;
  if (vw) memset(vw, 0, (size_t)nvx * sizeof(REAL));
#pragma omp parallel for schedule(static)
  for (long idx = 0; idx < npx; ++idx) {
    if (slices_mask && !slices_mask[idx]) continue;
    const REAL s = slices[idx];
    const int ix = (int)(idx % w), iy = (int)((idx / w) % h), is = (int)(idx / ((long)h * w));
    FN(PixelFrame) f;
    FN(pixel_frame)(transforms + is * 12, ix, iy, h, w, D, H, W, res_slice, &f);
    const REAL weight = FN(unmasked_weight)(&f, psf, D, H, W, d_p, h_p, w_p, interp_psf);
    if (weight < 0.5) continue; /* Q4 */
    NSV_TAP_LOOP_BEGIN
    if (interp_psf) {
      FN(NearestTap) t;
      if (!FN(nearest_tap)(&f, p, sy, sz, d_p, h_p, w_p, &t)) continue;
      tap = FN(psf_resampled)(psf, &t.pc);
      tap /= weight;
      if (vol_mask && !vol_mask[t.vox]) continue;
      {
        const REAL add = tap * s;
#pragma omp atomic
        vol[t.vox] += add;
      }
      if (vw) {
#pragma omp atomic
        vw[t.vox] += tap;
      }
    } else {
      FN(Cell) cell;
      FN(cell_at)(p, sy, sz, &cell);
      tap /= weight;
      for (int k = 0; k < 8; ++k) {
        const int c = FN(kCornerOrder)[k];
        const int iv = cell.base + cell.off[c];
        if (vol_mask && !vol_mask[iv]) continue;
        const REAL pw = cell.wt[c] * tap;
        const REAL add = pw * s;
#pragma omp atomic
        vol[iv] += add;
        if (vw) {
#pragma omp atomic
          vw[iv] += pw;
        }
      }
    }
    NSV_TAP_LOOP_END
  }
  if (equalize) FN(nsv_oracle_equalize_)(vol, vol_weight, 0, nvx);
}

/* ---------------------------------------------------------------- backward of A^T */
void FN(nsv_oracle_slice_acq_adjoint_backward_)(
    const REAL* transforms, REAL* grad_vol, const REAL* vol_weight, const unsigned char* vol_mask,
    const REAL* psf, const REAL* slices, const unsigned char* slices_mask, const REAL* vol,
    REAL* grad_slices, REAL* grad_transforms, int D, int H, int W, int d_p, int h_p, int w_p, int n,
    int h, int w, REAL res_slice, int interp_psf, int equalize) {
  const int sy = W, sz = H * W;
  const long npx = (long)n * h * w;
  if (equalize) FN(nsv_oracle_equalize_)(grad_vol, vol_weight, 1, (long)D * H * W); /* in place */
  const REAL* resid = equalize ? vol : NULL;
  if (grad_slices) memset(grad_slices, 0, (size_t)npx * sizeof(REAL));
  if (grad_transforms) memset(grad_transforms, 0, (size_t)n * 12 * sizeof(REAL));
#pragma omp parallel for schedule(static)
  for (long idx = 0; idx < npx; ++idx) {
    if (slices_mask && !slices_mask[idx]) continue;
    const int ix = (int)(idx % w), iy = (int)((idx / w) % h), is = (int)(idx / ((long)h * w));
    FN(PixelFrame) f;
    FN(pixel_frame)(transforms + is * 12, ix, iy, h, w, D, H, W, res_slice, &f);
    REAL val = 0, weight = 0;
    FN(TfGrad) acc;
    memset(&acc, 0, sizeof acc);
    NSV_TAP_LOOP_BEGIN
    REAL tapval = 0;
    if (interp_psf) {
      FN(NearestTap) t;
      t.rx = (int)round(p[0]); t.ry = (int)round(p[1]); t.rz = (int)round(p[2]);
      t.vox = t.rz * sz + t.ry * sy + t.rx;
      if (vol_mask && !vol_mask[t.vox]) continue;
      tapval = grad_vol[t.vox];
      if (!FN(nearest_tap)(&f, p, sy, sz, d_p, h_p, w_p, &t)) continue;
      tap = FN(psf_resampled)(psf, &t.pc);
      if (grad_transforms) {
        REAL d[3] = {0, 0, 0};
        for (int k = 0; k < 8; ++k) {
          const int c = FN(kCornerOrder)[k];
          FN(cell_grad_accum)(&t.pc, c, psf[t.pc.base + t.pc.off[c]], d);
        }
        const REAL sc = resid ? (slices[idx] - resid[t.vox]) * tapval : slices[idx] * tapval;
        d[0] *= sc; d[1] *= sc; d[2] *= sc;
        FN(tfgrad_nearest)(&acc, d, &t, D, H, W);
      }
    } else {
      FN(Cell) cell;
      FN(cell_at)(p, sy, sz, &cell);
      if (grad_slices)
        for (int k = 0; k < 8; ++k) {
          const int c = FN(kCornerOrder)[k];
          const int iv = cell.base + cell.off[c];
          if (vol_mask && !vol_mask[iv]) continue;
          tapval += cell.wt[c] * grad_vol[iv];
        }
      if (grad_transforms) {
        REAL d[3] = {0, 0, 0};
        for (int k = 0; k < 8; ++k) {
          const int c = FN(kCornerOrder)[k];
          const int iv = cell.base + cell.off[c];
          if (vol_mask && !vol_mask[iv]) continue;
          const REAL sc =
              resid ? (slices[idx] - resid[iv]) * grad_vol[iv] : slices[idx] * grad_vol[iv];
          FN(cell_grad_accum)(&cell, c, sc, d);
        }
        d[0] *= tap; d[1] *= tap; d[2] *= tap;
        FN(tfgrad_linear)(&acc, &f, d, tx, ty, tz);
      }
    }
    val += tap * tapval;
    weight += tap;
    NSV_TAP_LOOP_END
    if (weight > 0) {
      if (grad_slices) grad_slices[idx] = val / weight;
      if (grad_transforms)
        for (int k = 0; k < 12; ++k) {
          const REAL add = acc.g[k] / weight;
#pragma omp atomic
          grad_transforms[is * 12 + k] += add;
        }
    }
  }
}

#undef NSV_TAP_LOOP_BEGIN
#undef NSV_TAP_LOOP_END
#undef FN
#undef NSV_CAT
#undef NSV_CAT_
