// slice_acq_impl.cuh -- kernel B, generic implementation: PSF-weighted slice acquisition A (gather), its adjoint A^T
// (scatter) and the backward passes of both, every mode of the reference (interp_psf, masks, equalize), float and double.
//
// Semantics -- including quirks Q1-Q9 of SURVEY.md s.8c -- follow slice_acq_cuda_kernel.cu:18-950; the organisation does not:
//   * the PSF is staged once per CTA in shared memory and compacted to its non-zero taps (the
//     reference re-reads every tap, zero or not, from global memory per pixel);
//   * a CTA owns a 32x8 pixel patch of ONE slice and a warp an 8x4 sub-patch, so the 8-corner
//     gathers of neighbouring lanes fall into the same L1 lines, and the 12 pose-gradient terms are
//     reduced warp -> CTA before touching global memory (12 atomics per CTA instead of per pixel);
//   * scatter passes use fire-and-forget reductions (RED) on the caller's stream;
//   * the grid is sized to the SM count and strides over patches.
// Included twice, into two translation units with different arithmetic:
//   slice_acq.cu       NSV_SA_NS = sa_exact, compiled with -fmad=false: the gather passes reproduce the reference's C
//                      arithmetic bit for bit (test / verification mode, nsv_set_slice_acq_exact(1));
//   slice_acq_fast.cu  NSV_SA_NS = sa_fma, FMA contraction on: serves double precision and interp_psf beside the
//                      specialised fast kernels of that file (the product default).
#pragma once
#include "nsv_common.cuh"

namespace nsv {
namespace NSV_SA_NS {
namespace {

constexpr int kPatchW = 32, kPatchH = 8, kThreads = kPatchW * kPatchH;
constexpr int kMaxTaps = 4096;

struct Dims {
  int D, H, W, d_p, h_p, w_p, n, h, w;
};

// shared-memory view: raw PSF (for interp_psf resampling) + compacted non-zero taps in tap order
template <typename T>
struct PsfStage {
  T* raw;        // [ntaps]
  T* val;        // [nnz]
  int* xyz;      // [nnz] packed (tx & 0xff) | (ty & 0xff) << 8 | (tz & 0xff) << 16, signed bytes
  int nnz;
};

template <typename T>
__device__ PsfStage<T> stage_psf(const T* __restrict__ psf, const Dims& d, unsigned char* smem) {
  const int ntaps = d.d_p * d.h_p * d.w_p;
  PsfStage<T> st;
  st.raw = reinterpret_cast<T*>(smem);
  st.val = st.raw + ntaps;
  st.xyz = reinterpret_cast<int*>(st.val + ntaps);
  __shared__ int s_nnz;
  for (int i = threadIdx.x; i < ntaps; i += blockDim.x) st.raw[i] = psf[i];
  __syncthreads();
  if (threadIdx.x < 32) {  // ordered compaction by warp 0 (keeps the reference's summation order)
    int base = 0;
    for (int i0 = 0; i0 < ntaps; i0 += 32) {
      const int i = i0 + threadIdx.x;
      const T v = i < ntaps ? st.raw[i] : T(0);
      const bool keep = v != T(0);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int pos = base + __popc(m & ((1u << threadIdx.x) - 1));
        const int tx = i % d.w_p - d.w_p / 2, ty = (i / d.w_p) % d.h_p - d.h_p / 2, tz = i / (d.w_p * d.h_p) - d.d_p / 2;
        st.val[pos] = v;
        st.xyz[pos] = (tx & 0xff) | ((ty & 0xff) << 8) | ((tz & 0xff) << 16);
      }
      base += __popc(m);
    }
    if (threadIdx.x == 0) s_nnz = base;
  }
  __syncthreads();
  st.nnz = s_nnz;
  return st;
}

template <typename T>
struct Frame {
  T r[3][3], s[3], c[3];
};

template <typename T>
__device__ __forceinline__ void make_frame(const T* __restrict__ tf, int ix, int iy, const Dims& d, T res_slice, Frame<T>& f) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) f.r[a][b] = tf[a * 4 + b];
  f.s[0] = (T)((ix - (d.w - 1) / 2.) * res_slice + tf[3]);  // Q8: double, then narrowed
  f.s[1] = (T)((iy - (d.h - 1) / 2.) * res_slice + tf[7]);
  f.s[2] = tf[11];
  const double half[3] = {(d.W - 1) / 2., (d.H - 1) / 2., (d.D - 1) / 2.};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const T v = f.r[a][0] * f.s[0] + f.r[a][1] * f.s[1] + f.r[a][2] * f.s[2];
    f.c[a] = (T)(v + half[a]);
  }
}

template <typename T>
struct Cell {
  int base;
  T fx[2], fy[2], fz[2];
  __device__ __forceinline__ T wt(int c) const { return fx[c & 1] * fy[(c >> 1) & 1] * fz[c >> 2]; }
};

template <typename T>
__device__ __forceinline__ Cell<T> make_cell(const T p[3], int sy, int sz) {
  Cell<T> c;
  const int x0 = (int)floor(p[0]), y0 = (int)floor(p[1]), z0 = (int)floor(p[2]);
  const T wx = p[0] - x0, wy = p[1] - y0, wz = p[2] - z0;
  c.fx[0] = 1 - wx; c.fx[1] = wx;
  c.fy[0] = 1 - wy; c.fy[1] = wy;
  c.fz[0] = 1 - wz; c.fz[1] = wz;
  c.base = z0 * sz + y0 * sy + x0;
  return c;
}

__device__ __forceinline__ int corner_off(int c, int sy, int sz) { return (c & 1) + ((c >> 1) & 1) * sy + (c >> 2) * sz; }

// visiting order of the reference: 000,100,010,001,110,101,011,111 (bit0 = +x), one nibble per step
__device__ __forceinline__ constexpr int corner_at(int k) { return (0x76534210u >> (4 * k)) & 7; }
#define NSV_CORNERS(c) _Pragma("unroll") for (int k_ = 0; k_ < 8; ++k_) if (const int c = corner_at(k_); true)

// pose-gradient side only (never feeds a bit-exact output): evaluated in double, see TfGrad
template <typename T>
__device__ __forceinline__ void cell_grad(const Cell<T>& cell, int c, double v, double d[3]) {
  const int bx = c & 1, by = (c >> 1) & 1, bz = c >> 2;
  const double gx = (double)cell.fy[by] * cell.fz[bz] * v, gy = (double)cell.fx[bx] * cell.fz[bz] * v, gz = (double)cell.fx[bx] * cell.fy[by] * v;
  d[0] = bx ? d[0] + gx : d[0] - gx;
  d[1] = by ? d[1] + gy : d[1] - gy;
  d[2] = bz ? d[2] + gz : d[2] - gz;
}

template <typename T>
__device__ __forceinline__ bool tap_pos(const Frame<T>& f, int packed, const Dims& d, T p[3], int t[3]) {
  t[0] = (int)(signed char)(packed & 0xff);
  t[1] = (int)(signed char)((packed >> 8) & 0xff);
  t[2] = (int)(signed char)((packed >> 16) & 0xff);
#pragma unroll
  for (int a = 0; a < 3; ++a) p[a] = f.c[a] + f.r[a][0] * t[0] + f.r[a][1] * t[1] + f.r[a][2] * t[2];
  return !(p[0] < 0 || p[1] < 0 || p[2] < 0 || p[0] >= d.W - 1 || p[1] >= d.H - 1 || p[2] >= d.D - 1);  // Q5
}

// interp_psf mode (Q9): nearest voxel + PSF resampled at the voxel's offset in the slice frame
template <typename T>
struct Nearest {
  int vox, r[3];
  Cell<T> pc;
};

template <typename T>
__device__ __forceinline__ void nearest_voxel(const T p[3], int sy, int sz, Nearest<T>& t) {
  t.r[0] = (int)round(p[0]);  // Q7: half away from zero
  t.r[1] = (int)round(p[1]);
  t.r[2] = (int)round(p[2]);
  t.vox = t.r[2] * sz + t.r[1] * sy + t.r[0];
}

template <typename T>
__device__ __forceinline__ bool nearest_psf_cell(const Frame<T>& f, const Dims& d, Nearest<T>& t) {
  const T dx = t.r[0] - f.c[0], dy = t.r[1] - f.c[1], dz = t.r[2] - f.c[2];
  T q[3];
  q[0] = (T)(f.r[0][0] * dx + f.r[1][0] * dy + f.r[2][0] * dz + (d.w_p - 1) / 2.);
  q[1] = (T)(f.r[0][1] * dx + f.r[1][1] * dy + f.r[2][1] * dz + (d.h_p - 1) / 2.);
  q[2] = (T)(f.r[0][2] * dx + f.r[1][2] * dy + f.r[2][2] * dz + (d.d_p - 1) / 2.);
  if (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= d.w_p - 1 || q[1] >= d.h_p - 1 || q[2] >= d.d_p - 1) return false;
  t.pc = make_cell<T>(q, d.w_p, d.w_p * d.h_p);
  return true;
}

template <typename T>
__device__ __forceinline__ T psf_resampled(const T* raw, const Cell<T>& pc, const Dims& d) {
  T v = 0;
  NSV_CORNERS(c) v += pc.wt(c) * raw[pc.base + corner_off(c, d.w_p, d.w_p * d.h_p)];
  return v;
}

template <typename T>
__device__ __forceinline__ void psf_cell_grad(const T* raw, const Cell<T>& pc, const Dims& d, double g[3]) {
  g[0] = g[1] = g[2] = 0;
  NSV_CORNERS(c) cell_grad<T>(pc, c, raw[pc.base + corner_off(c, d.w_p, d.w_p * d.h_p)], g);
}

// Q3: weight used by backward / adjoint = in-bounds taps, vol_mask ignored
template <typename T>
__device__ T unmasked_weight(const Frame<T>& f, const PsfStage<T>& st, const Dims& d, bool interp_psf) {
  const int sy = d.W, sz = d.H * d.W;
  T weight = 0;
  for (int i = 0; i < st.nnz; ++i) {
    T p[3];
    int t[3];
    if (!tap_pos<T>(f, st.xyz[i], d, p, t)) continue;
    T tap = st.val[i];
    if (interp_psf) {
      Nearest<T> nt;
      nearest_voxel<T>(p, sy, sz, nt);
      if (!nearest_psf_cell<T>(f, d, nt)) continue;
      tap = psf_resampled<T>(st.raw, nt.pc, d);
    }
    weight += tap;
  }
  return weight;
}

// pose-gradient accumulator (12 terms), reduced over the CTA (one slice per CTA patch).  The 12 sums run over every
// (pixel, tap) of a slice with heavy cancellation; they are accumulated in DOUBLE from the first product on (per thread,
// across the warp, across the CTA) and rounded to T once per CTA: measured against the fp64 operator the fp32-accumulated
// version was 3.7x further off than the reference's per-thread float atomics on one case (profiles/r02_kernelB_*).
template <typename T>
struct TfGrad {
  double g[12];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < 12; ++k) g[k] = 0;
  }
  __device__ __forceinline__ void add_linear(const Frame<T>& f, const double d[3], const int t[3]) {
    const double q[3] = {(double)(f.s[0] + t[0]), (double)(f.s[1] + t[1]), (double)(f.s[2] + t[2])};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) g[r * 4 + c] += (double)d[r] * q[c];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c * 4 + 3] += (double)d[0] * f.r[0][c] + (double)d[1] * f.r[1][c] + (double)d[2] * f.r[2][c];
  }
  __device__ __forceinline__ void add_nearest(const double d[3], const Nearest<T>& nt, const Dims& dm) {
    const double q[3] = {nt.r[0] - (dm.W - 1) / 2., nt.r[1] - (dm.H - 1) / 2., nt.r[2] - (dm.D - 1) / 2.};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) g[r * 4 + c] += (double)d[c] * q[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c * 4 + 3] -= (double)d[c];
  }
  __device__ __forceinline__ void scale(T inv) {
#pragma unroll
    for (int k = 0; k < 12; ++k) g[k] *= (double)inv;
  }
};

template <typename T>
__device__ void block_reduce_tf(const TfGrad<T>& acc, T* __restrict__ grad_tf_slice) {
  __shared__ double s_part[kThreads / 32][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    double v = acc.g[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double v = 0;
#pragma unroll
    for (int wv = 0; wv < kThreads / 32; ++wv) v += s_part[wv][threadIdx.x];
    if (v != 0) atomicAdd(grad_tf_slice + threadIdx.x, (T)v);
  }
  __syncthreads();
}

// patch -> (slice, ix, iy) of this thread; returns false when the thread is outside the slice
struct PatchIter {
  int tiles_x, tiles_y;
  long n_patches;
  __device__ PatchIter(const Dims& d)
      : tiles_x((d.w + kPatchW - 1) / kPatchW), tiles_y((d.h + kPatchH - 1) / kPatchH) {
    n_patches = (long)d.n * tiles_x * tiles_y;
  }
  __device__ __forceinline__ bool locate(long patch, const Dims& d, int& is, int& ix, int& iy) const {
    const int tx = (int)(patch % tiles_x), ty = (int)((patch / tiles_x) % tiles_y);
    is = (int)(patch / ((long)tiles_x * tiles_y));
    // lane -> 8x4 sub-patch of its warp; warps tile the 32x8 patch 4 across, 2 down
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ix = tx * kPatchW + (warp & 3) * 8 + (lane & 7);
    iy = ty * kPatchH + (warp >> 2) * 4 + (lane >> 3);
    return ix < d.w && iy < d.h;
  }
};

// ------------------------------------------------------------------------------------- forward (A)
template <typename T>
__global__ void __launch_bounds__(kThreads)
    forward_kernel(const T* __restrict__ transforms, const T* __restrict__ vol, const uint8_t* __restrict__ vol_mask,
                   const uint8_t* __restrict__ slices_mask, const T* __restrict__ psf, T* __restrict__ slices,
                   T* __restrict__ slices_weight, Dims d, T res_slice, bool interp_psf) {
  extern __shared__ __align__(16) unsigned char smem[];
  const PsfStage<T> st = stage_psf<T>(psf, d, smem);
  const int sy = d.W, sz = d.H * d.W;
  const PatchIter it(d);
  for (long patch = blockIdx.x; patch < it.n_patches; patch += gridDim.x) {
    int is, ix, iy;
    if (!it.locate(patch, d, is, ix, iy)) continue;
    const long idx = ((long)is * d.h + iy) * d.w + ix;
    if (slices_mask && !slices_mask[idx]) continue;
    Frame<T> f;
    make_frame<T>(transforms + is * 12, ix, iy, d, res_slice, f);
    T val = 0, weight = 0;
    for (int i = 0; i < st.nnz; ++i) {
      T p[3];
      int t[3];
      if (!tap_pos<T>(f, st.xyz[i], d, p, t)) continue;
      T tap = st.val[i];
      if (interp_psf) {
        Nearest<T> nt;
        nearest_voxel<T>(p, sy, sz, nt);
        if (vol_mask && !vol_mask[nt.vox]) continue;
        const T v = __ldg(vol + nt.vox);
        if (!nearest_psf_cell<T>(f, d, nt)) continue;
        tap = psf_resampled<T>(st.raw, nt.pc, d);
        val += tap * v;
        weight += tap;
      } else {
        const Cell<T> cell = make_cell<T>(p, sy, sz);
        T vv[8];
        if (!vol_mask) {
#pragma unroll
          for (int c = 0; c < 8; ++c) vv[c] = __ldg(vol + cell.base + corner_off(c, sy, sz));
          NSV_CORNERS(c) {
            const T pw = cell.wt(c) * tap;
            val += pw * vv[c];
            weight += pw;
          }
        } else {
          NSV_CORNERS(c) {
            const int iv = cell.base + corner_off(c, sy, sz);
            if (!vol_mask[iv]) continue;
            const T pw = cell.wt(c) * tap;
            val += pw * __ldg(vol + iv);
            weight += pw;
          }
        }
      }
    }
    if (weight > 0) {  // Q1
      slices[idx] = val / weight;
      if (slices_weight) slices_weight[idx] = weight;
    }
  }
}

// ------------------------------------------------------------------------------ backward of A
template <typename T>
__global__ void __launch_bounds__(kThreads)
    backward_kernel(const T* __restrict__ transforms, const T* __restrict__ vol, const uint8_t* __restrict__ vol_mask,
                    const T* __restrict__ psf, const T* __restrict__ grad_slices, const uint8_t* __restrict__ slices_mask,
                    T* __restrict__ grad_vol, T* __restrict__ grad_tf, Dims d, T res_slice, bool interp_psf) {
  extern __shared__ __align__(16) unsigned char smem[];
  const PsfStage<T> st = stage_psf<T>(psf, d, smem);
  const int sy = d.W, sz = d.H * d.W;
  const PatchIter it(d);
  for (long patch = blockIdx.x; patch < it.n_patches; patch += gridDim.x) {
    int is, ix, iy;
    const bool inside = it.locate(patch, d, is, ix, iy);
    const long idx = ((long)is * d.h + iy) * d.w + ix;
    TfGrad<T> acc;
    acc.clear();
    T gs = 0;
    bool active = inside && !(slices_mask && !slices_mask[idx]);
    if (active) {
      gs = grad_slices[idx];
      active = gs != T(0);  // Q2
    }
    if (active) {
      Frame<T> f;
      make_frame<T>(transforms + is * 12, ix, iy, d, res_slice, f);
      const T weight = unmasked_weight<T>(f, st, d, interp_psf);
      if (weight != T(0)) {
        gs /= weight;
        for (int i = 0; i < st.nnz; ++i) {
          T p[3];
          int t[3];
          if (!tap_pos<T>(f, st.xyz[i], d, p, t)) continue;
          T tap = st.val[i];
          if (interp_psf) {
            Nearest<T> nt;
            nearest_voxel<T>(p, sy, sz, nt);
            if (!nearest_psf_cell<T>(f, d, nt)) continue;
            if (vol_mask && !vol_mask[nt.vox]) continue;
            if (grad_vol) atomicAdd(grad_vol + nt.vox, psf_resampled<T>(st.raw, nt.pc, d) * gs);
            if (grad_tf) {
              double g[3];
              psf_cell_grad<T>(st.raw, nt.pc, d, g);
              const double sc = (double)gs * __ldg(vol + nt.vox);
              g[0] *= sc; g[1] *= sc; g[2] *= sc;
              acc.add_nearest(g, nt, d);
            }
          } else {
            const Cell<T> cell = make_cell<T>(p, sy, sz);
            tap *= gs;
            if (grad_vol) {
              NSV_CORNERS(c) {
                const int iv = cell.base + corner_off(c, sy, sz);
                if (vol_mask && !vol_mask[iv]) continue;
                atomicAdd(grad_vol + iv, cell.wt(c) * tap);
              }
            }
            if (grad_tf) {
              double g[3] = {0, 0, 0};
              NSV_CORNERS(c) {
                const int iv = cell.base + corner_off(c, sy, sz);
                if (vol_mask && !vol_mask[iv]) continue;
                cell_grad<T>(cell, c, (double)tap * __ldg(vol + iv), g);
              }
              acc.add_linear(f, g, t);
            }
          }
        }
      }
    }
    if (grad_tf) block_reduce_tf<T>(acc, grad_tf + is * 12);
  }
}

// ------------------------------------------------------------------------- adjoint forward (A^T)
template <typename T>
__global__ void __launch_bounds__(kThreads)
    adjoint_forward_kernel(const T* __restrict__ transforms, T* __restrict__ vol, T* __restrict__ vol_weight,
                           const uint8_t* __restrict__ vol_mask, const T* __restrict__ psf, const T* __restrict__ slices,
                           const uint8_t* __restrict__ slices_mask, Dims d, T res_slice, bool interp_psf) {
  extern __shared__ __align__(16) unsigned char smem[];
  const PsfStage<T> st = stage_psf<T>(psf, d, smem);
  const int sy = d.W, sz = d.H * d.W;
  const PatchIter it(d);
  for (long patch = blockIdx.x; patch < it.n_patches; patch += gridDim.x) {
    int is, ix, iy;
    if (!it.locate(patch, d, is, ix, iy)) continue;
    const long idx = ((long)is * d.h + iy) * d.w + ix;
    if (slices_mask && !slices_mask[idx]) continue;
    const T s = slices[idx];
    Frame<T> f;
    make_frame<T>(transforms + is * 12, ix, iy, d, res_slice, f);
    const T weight = unmasked_weight<T>(f, st, d, interp_psf);
    if (weight < 0.5) continue;  // Q4
    for (int i = 0; i < st.nnz; ++i) {
      T p[3];
      int t[3];
      if (!tap_pos<T>(f, st.xyz[i], d, p, t)) continue;
      T tap = st.val[i];
      if (interp_psf) {
        Nearest<T> nt;
        nearest_voxel<T>(p, sy, sz, nt);
        if (!nearest_psf_cell<T>(f, d, nt)) continue;
        tap = psf_resampled<T>(st.raw, nt.pc, d);
        tap /= weight;
        if (vol_mask && !vol_mask[nt.vox]) continue;
        atomicAdd(vol + nt.vox, tap * s);
        if (vol_weight) atomicAdd(vol_weight + nt.vox, tap);
      } else {
        const Cell<T> cell = make_cell<T>(p, sy, sz);
        tap /= weight;
        NSV_CORNERS(c) {
          const int iv = cell.base + corner_off(c, sy, sz);
          if (vol_mask && !vol_mask[iv]) continue;
          const T pw = cell.wt(c) * tap;
          atomicAdd(vol + iv, pw * s);
          if (vol_weight) atomicAdd(vol_weight + iv, pw);
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(1024) equalize_kernel(T* __restrict__ vol, const T* __restrict__ vol_weight, bool is_grad, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const T wgt = vol_weight[i];
    if (!(wgt > 0)) continue;
    if (is_grad && wgt < 1e-3)
      vol[i] = (T)(vol[i] / 1e-3);
    else
      vol[i] /= wgt;
  }
}

// ------------------------------------------------------------------------------ backward of A^T
template <typename T>
__global__ void __launch_bounds__(kThreads)
    adjoint_backward_kernel(const T* __restrict__ transforms, const T* __restrict__ grad_vol, const T* __restrict__ psf,
                            const T* __restrict__ slices, const uint8_t* __restrict__ slices_mask, const T* __restrict__ resid,
                            const uint8_t* __restrict__ vol_mask, T* __restrict__ grad_slices, T* __restrict__ grad_tf, Dims d,
                            T res_slice, bool interp_psf) {
  extern __shared__ __align__(16) unsigned char smem[];
  const PsfStage<T> st = stage_psf<T>(psf, d, smem);
  const int sy = d.W, sz = d.H * d.W;
  const PatchIter it(d);
  for (long patch = blockIdx.x; patch < it.n_patches; patch += gridDim.x) {
    int is, ix, iy;
    const bool inside = it.locate(patch, d, is, ix, iy);
    const long idx = ((long)is * d.h + iy) * d.w + ix;
    TfGrad<T> acc;
    acc.clear();
    if (inside && !(slices_mask && !slices_mask[idx])) {
      Frame<T> f;
      make_frame<T>(transforms + is * 12, ix, iy, d, res_slice, f);
      const T sval = slices[idx];
      T val = 0, weight = 0;
      for (int i = 0; i < st.nnz; ++i) {
        T p[3];
        int t[3];
        if (!tap_pos<T>(f, st.xyz[i], d, p, t)) continue;
        T tap = st.val[i];
        T tapval = 0;
        if (interp_psf) {
          Nearest<T> nt;
          nearest_voxel<T>(p, sy, sz, nt);
          if (vol_mask && !vol_mask[nt.vox]) continue;
          tapval = __ldg(grad_vol + nt.vox);
          if (!nearest_psf_cell<T>(f, d, nt)) continue;
          tap = psf_resampled<T>(st.raw, nt.pc, d);
          if (grad_tf) {
            double g[3];
            psf_cell_grad<T>(st.raw, nt.pc, d, g);
            const double sc = resid ? ((double)sval - __ldg(resid + nt.vox)) * tapval : (double)sval * tapval;
            g[0] *= sc; g[1] *= sc; g[2] *= sc;
            acc.add_nearest(g, nt, d);
          }
        } else {
          const Cell<T> cell = make_cell<T>(p, sy, sz);
          if (grad_slices) {
            NSV_CORNERS(c) {
              const int iv = cell.base + corner_off(c, sy, sz);
              if (vol_mask && !vol_mask[iv]) continue;
              tapval += cell.wt(c) * __ldg(grad_vol + iv);
            }
          }
          if (grad_tf) {
            double g[3] = {0, 0, 0};
            NSV_CORNERS(c) {
              const int iv = cell.base + corner_off(c, sy, sz);
              if (vol_mask && !vol_mask[iv]) continue;
              const T gv = __ldg(grad_vol + iv);
              const double sc = resid ? ((double)sval - __ldg(resid + iv)) * gv : (double)sval * gv;
              cell_grad<T>(cell, c, sc, g);
            }
            g[0] *= tap; g[1] *= tap; g[2] *= tap;
            acc.add_linear(f, g, t);
          }
        }
        val += tap * tapval;
        weight += tap;
      }
      if (weight > 0) {
        if (grad_slices) grad_slices[idx] = val / weight;
        acc.scale(T(1) / weight);
      } else {
        acc.clear();
      }
    }
    if (grad_tf) block_reduce_tf<T>(acc, grad_tf + is * 12);
  }
}

template <typename T>
size_t psf_smem_bytes(const Dims& d) {
  const size_t ntaps = (size_t)d.d_p * d.h_p * d.w_p;
  return ntaps * (2 * sizeof(T) + sizeof(int));
}

int check_dims(const char* name, const Dims& d) {
  NSV_REQUIRE(d.D > 0 && d.H > 0 && d.W > 0 && d.d_p > 0 && d.h_p > 0 && d.w_p > 0 && d.n >= 0 && d.h > 0 && d.w > 0,
              "%s: non-positive dimension", name);
  NSV_REQUIRE((long)d.d_p * d.h_p * d.w_p <= kMaxTaps, "%s: PSF larger than %d taps", name, kMaxTaps);
  NSV_REQUIRE(d.d_p <= 255 && d.h_p <= 255 && d.w_p <= 255, "%s: PSF extent above 255", name);
  NSV_REQUIRE((long)d.D * d.H * d.W < (1L << 31) && (long)d.n * d.h * d.w < (1L << 31),
              "%s: int32 flat index overflow (same limit as the reference, slice_acq_cuda_kernel.cu:33-34)", name);
  return NSV_OK;
}

int grid_for(const Dims& d, int ctas_per_sm) {
  const long patches = (long)d.n * ((d.w + kPatchW - 1) / kPatchW) * ((d.h + kPatchH - 1) / kPatchH);
  const long cap = (long)num_sms() * ctas_per_sm;
  return (int)(patches < cap ? (patches > 0 ? patches : 1) : cap);
}

template <typename K>
int prep(K kernel, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  return NSV_OK;
}

template <typename T>
int run_equalize(T* vol, const T* vol_weight, int is_grad, int64_t n, void* stream) {
  NSV_REQUIRE(n >= 0 && (n == 0 || (vol && vol_weight)), "nsv_equalize: bad arguments");
  if (n == 0) return NSV_OK;
  const int64_t blocks = (n + 1023) / 1024;
  const int grid = (int)(blocks < (int64_t)num_sms() * 2 ? blocks : (int64_t)num_sms() * 2);
  equalize_kernel<T><<<grid, 1024, 0, (cudaStream_t)stream>>>(vol, vol_weight, is_grad != 0, n);
  return check_launch("nsv_equalize");
}

template <typename T>
int run_forward(const T* transforms, const T* vol, const uint8_t* vol_mask, const uint8_t* slices_mask, const T* psf,
                T* slices, T* slices_weight, Dims d, T res_slice, int interp_psf, void* stream) {
  if (int e = check_dims("nsv_slice_acq_forward", d)) return e;
  if (d.n == 0) return NSV_OK;
  NSV_REQUIRE(transforms && vol && psf && slices, "nsv_slice_acq_forward: NULL pointer");
  const size_t smem = psf_smem_bytes<T>(d);
  if (int e = prep(forward_kernel<T>, smem)) return e;
  forward_kernel<T><<<grid_for(d, 8), kThreads, smem, (cudaStream_t)stream>>>(transforms, vol, vol_mask, slices_mask, psf, slices,
                                                                             slices_weight, d, res_slice, interp_psf != 0);
  return check_launch("nsv_slice_acq_forward");
}

template <typename T>
int run_backward(const T* transforms, const T* vol, const uint8_t* vol_mask, const T* psf, const T* grad_slices,
                 const uint8_t* slices_mask, T* grad_vol, T* grad_tf, Dims d, T res_slice, int interp_psf, void* stream) {
  if (int e = check_dims("nsv_slice_acq_backward", d)) return e;
  if (d.n == 0 || (!grad_vol && !grad_tf)) return NSV_OK;
  NSV_REQUIRE(transforms && vol && psf && grad_slices, "nsv_slice_acq_backward: NULL pointer");
  const size_t smem = psf_smem_bytes<T>(d);
  if (int e = prep(backward_kernel<T>, smem)) return e;
  backward_kernel<T><<<grid_for(d, 8), kThreads, smem, (cudaStream_t)stream>>>(transforms, vol, vol_mask, psf, grad_slices,
                                                                              slices_mask, grad_vol, grad_tf, d, res_slice,
                                                                              interp_psf != 0);
  return check_launch("nsv_slice_acq_backward");
}

template <typename T>
int run_adjoint_forward(const T* transforms, const T* psf, const T* slices, const uint8_t* slices_mask, const uint8_t* vol_mask,
                        T* vol, T* vol_weight, Dims d, T res_slice, int interp_psf, int equalize, void* stream) {
  if (int e = check_dims("nsv_slice_acq_adjoint_forward", d)) return e;
  if (d.n == 0) return NSV_OK;
  NSV_REQUIRE(transforms && psf && slices && vol, "nsv_slice_acq_adjoint_forward: NULL pointer");
  NSV_REQUIRE(!equalize || vol_weight, "nsv_slice_acq_adjoint_forward: equalize needs vol_weight");
  const size_t smem = psf_smem_bytes<T>(d);
  if (int e = prep(adjoint_forward_kernel<T>, smem)) return e;
  adjoint_forward_kernel<T><<<grid_for(d, 8), kThreads, smem, (cudaStream_t)stream>>>(
      transforms, vol, equalize ? vol_weight : (T*)nullptr, vol_mask, psf, slices, slices_mask, d, res_slice, interp_psf != 0);
  if (int e = check_launch("nsv_slice_acq_adjoint_forward")) return e;
  if (equalize) return run_equalize<T>(vol, vol_weight, 0, (int64_t)d.D * d.H * d.W, stream);
  return NSV_OK;
}

template <typename T>
int run_adjoint_backward(const T* transforms, T* grad_vol, const T* vol_weight, const uint8_t* vol_mask, const T* psf,
                         const T* slices, const uint8_t* slices_mask, const T* vol, T* grad_slices, T* grad_tf, Dims d,
                         T res_slice, int interp_psf, int equalize, void* stream) {
  if (int e = check_dims("nsv_slice_acq_adjoint_backward", d)) return e;
  NSV_REQUIRE(grad_vol && (d.n == 0 || (transforms && psf && slices)), "nsv_slice_acq_adjoint_backward: NULL pointer");
  NSV_REQUIRE(!equalize || (vol_weight && vol), "nsv_slice_acq_adjoint_backward: equalize needs vol and vol_weight");
  if (equalize)
    if (int e = run_equalize<T>(grad_vol, vol_weight, 1, (int64_t)d.D * d.H * d.W, stream)) return e;
  if (d.n == 0 || (!grad_slices && !grad_tf)) return NSV_OK;
  const size_t smem = psf_smem_bytes<T>(d);
  if (int e = prep(adjoint_backward_kernel<T>, smem)) return e;
  adjoint_backward_kernel<T><<<grid_for(d, 8), kThreads, smem, (cudaStream_t)stream>>>(
      transforms, grad_vol, psf, slices, slices_mask, equalize ? vol : (const T*)nullptr, vol_mask, grad_slices, grad_tf, d,
      res_slice, interp_psf != 0);
  return check_launch("nsv_slice_acq_adjoint_backward");
}

}  // namespace


#define NSV_SA_DIMS \
  Dims { D, H, W, d_p, h_p, w_p, n, h, w }

// C++-linkage entry points of this translation unit's flavour (same argument lists as the C ABI)
#define NSV_SA_FLAVOUR_API(T, SUF)                                                                                        \
  int forward_##SUF(const T* transforms, const T* vol, const uint8_t* vol_mask, const uint8_t* slices_mask, const T* psf, \
                    T* slices, T* slices_weight, int D, int H, int W, int d_p, int h_p, int w_p, int n, int h, int w,     \
                    T res_slice, int interp_psf, void* stream) {                                                           \
    return run_forward<T>(transforms, vol, vol_mask, slices_mask, psf, slices, slices_weight, NSV_SA_DIMS, res_slice,     \
                          interp_psf, stream);                                                                             \
  }                                                                                                                        \
  int backward_##SUF(const T* transforms, const T* vol, const uint8_t* vol_mask, const T* psf, const T* grad_slices,      \
                     const uint8_t* slices_mask, T* grad_vol, T* grad_transforms, int D, int H, int W, int d_p, int h_p,  \
                     int w_p, int n, int h, int w, T res_slice, int interp_psf, void* stream) {                            \
    return run_backward<T>(transforms, vol, vol_mask, psf, grad_slices, slices_mask, grad_vol, grad_transforms,           \
                           NSV_SA_DIMS, res_slice, interp_psf, stream);                                                    \
  }                                                                                                                        \
  int adjoint_forward_##SUF(const T* transforms, const T* psf, const T* slices, const uint8_t* slices_mask,               \
                            const uint8_t* vol_mask, T* vol, T* vol_weight, int D, int H, int W, int d_p, int h_p,        \
                            int w_p, int n, int h, int w, T res_slice, int interp_psf, int equalize, void* stream) {       \
    return run_adjoint_forward<T>(transforms, psf, slices, slices_mask, vol_mask, vol, vol_weight, NSV_SA_DIMS,           \
                                  res_slice, interp_psf, equalize, stream);                                                \
  }                                                                                                                        \
  int adjoint_backward_##SUF(const T* transforms, T* grad_vol, const T* vol_weight, const uint8_t* vol_mask,              \
                             const T* psf, const T* slices, const uint8_t* slices_mask, const T* vol, T* grad_slices,     \
                             T* grad_transforms, int D, int H, int W, int d_p, int h_p, int w_p, int n, int h, int w,     \
                             T res_slice, int interp_psf, int equalize, void* stream) {                                    \
    return run_adjoint_backward<T>(transforms, grad_vol, vol_weight, vol_mask, psf, slices, slices_mask, vol,             \
                                   grad_slices, grad_transforms, NSV_SA_DIMS, res_slice, interp_psf, equalize, stream);   \
  }                                                                                                                        \
  int equalize_##SUF(T* vol, const T* vol_weight, int is_grad, int64_t DHW, void* stream) {                               \
    return run_equalize<T>(vol, vol_weight, is_grad, DHW, stream);                                                         \
  }

NSV_SA_FLAVOUR_API(float, f32)
NSV_SA_FLAVOUR_API(double, f64)

}  // namespace NSV_SA_NS
}  // namespace nsv
