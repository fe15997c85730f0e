// slice_acq_fast.cu -- kernel B, product flavour + the C ABI of the slice-acquisition family.
//
// C-ABI replacement for nesvor.slice_acq_cuda.{forward, backward, adjoint_forward, adjoint_backward}
// (nesvor/slice_acquisition/slice_acq_cuda.cpp:61-160; kernels slice_acq_cuda_kernel.cu:18-950).
//
// Three flavours sit behind the entry points:
//   sa_exact  slice_acq.cu, -fmad=false: bit-exact with the reference's C arithmetic (nsv_set_slice_acq_exact(1));
//   sa_fma    the same generic kernels with FMA contraction: double precision and interp_psf = true;
//   sa_fast   this file: fp32, interp_psf = false -- the mode every call site of the reference uses
//             (svort/models.py:62,161, svort/inference.py:388,436, srr.py) -- re-organised for the B200 memory system:
//
//   * pixels are classified against the PSF's bounding box once: all taps outside the volume -> nothing to do (the
//     reference walks every tap of every pixel; on the BASELINE stacks 82 % of the pixels lie outside); all taps inside ->
//     no per-tap bounds tests and the normalisation weight of the scatter / backward passes (their first tap loop, Q3) is
//     the pre-summed PSF; only border pixels take the checked loops;
//   * tap positions are evaluated with exactly the reference's expression (x_center + r11 ix_p + r12 iy_p + r13 iz_p,
//     slice_acq_cuda_kernel.cu:66-68, contracted to the same FMA chain by nvcc), so that taps lying ON a boundary plane of
//     lattice-aligned stacks fall on the same side as in the reference's own build; the compacted taps are one float4
//     {ix_p, iy_p, iz_p, psf} per tap in shared memory;
//   * the volume is gathered from / scattered into the copy whose FASTEST axis is the one a slice's pixel rows run along
//     (x-, y- or z-fastest; two transposed scratch copies are made / merged by a streaming pre- / post-pass), and warps are
//     32 x 1 pixel rows along that axis when the slice is within a few degrees of it (8 x 4 patches otherwise): a warp-wide
//     gather or reduction then touches 4-5 sectors instead of 13-32;
//   * (tuning bit 4, off by default: measured slower) in row mode neighbouring lanes' cells overlap by one voxel column:
//     the +1 corners can be handed to the neighbour with a shuffle and leave as ONE reduction;
//   * A^T without `equalize` skips pixels whose value is exactly zero (they add 0 to every voxel).
// Results agree with sa_exact to fp32 round-off (tests/test_gpu_slice_acq.py), not bit for bit.
#define NSV_SA_NS sa_fma
#include "slice_acq_impl.cuh"

#include <stdlib.h>

namespace nsv {
namespace sa_exact {
#define NSV_SA_DECLARE(T, SUF)                                                                                             \
  int forward_##SUF(const T*, const T*, const uint8_t*, const uint8_t*, const T*, T*, T*, int, int, int, int, int, int,   \
                    int, int, int, T, int, void*);                                                                         \
  int backward_##SUF(const T*, const T*, const uint8_t*, const T*, const T*, const uint8_t*, T*, T*, int, int, int, int,  \
                     int, int, int, int, int, T, int, void*);                                                              \
  int adjoint_forward_##SUF(const T*, const T*, const T*, const uint8_t*, const uint8_t*, T*, T*, int, int, int, int,     \
                            int, int, int, int, int, T, int, int, void*);                                                  \
  int adjoint_backward_##SUF(const T*, T*, const T*, const uint8_t*, const T*, const T*, const uint8_t*, const T*, T*,    \
                             T*, int, int, int, int, int, int, int, int, int, T, int, int, void*);                         \
  int equalize_##SUF(T*, const T*, int, int64_t, void*);
NSV_SA_DECLARE(float, f32)
NSV_SA_DECLARE(double, f64)
}  // namespace sa_exact

namespace sa_fast {
namespace {

constexpr int kThreads = 256, kTileF = 32, kTileS = 8, kMaxTaps = 4096;
enum : unsigned { kTunePerm = 1u, kTuneRow = 2u, kTuneMerge = 4u, kTuneZeroSkip = 8u, kTuneClassify = 16u };
constexpr float kEps = 1e-3f;  // safety margin (voxels) of the inside / outside classification

struct Geo {
  int D, H, W, d_p, h_p, w_p, n, h, w;
  float res;
  int tiles_max;  // tiles per slice, maximum over the two tile orientations
  unsigned tune;
};

struct Vols {  // the three layouts of one volume-shaped buffer: [0] x-fastest (the caller's), [1] y-fastest, [2] z-fastest
  float* p[3];
  __device__ __forceinline__ float* at(int perm) const { return perm == 0 ? p[0] : (perm == 1 ? p[1] : p[2]); }
};

// Everything the tap loops touch is kept in COPY order: axis k of the chosen layout is volume axis (perm + k) % 3, so
// that k = 0 is always the fastest (stride-1) axis and corner bit 0 the neighbour one element further in memory.
struct Ctx {  // per slice, rebuilt in shared memory whenever a CTA moves to another slice
  float R[3][3], T[3];         // rows of the rotation in copy order: R[k] = R_vol[(perm + k) % 3]
  float half[3];               // (dim - 1) / 2 per copy axis
  float top[3];                // dim - 1 per copy axis (in-bounds: 0 <= p < top, Q5)
  float lo_in[3], hi_in[3];    // every tap inside the volume  <=>  lo_in <= c <= hi_in on all axes
  float lo_out[3], hi_out[3];  // every tap outside            <=   c < lo_out or c > hi_out on some axis
  int s[3];                    // element strides per copy axis (s[0] == 1)
  int perm, fast_iy, row_mode, dir;
  int tiles_f, tiles_s;
};

struct Smem {
  Ctx ctx;
  float wsum;
  int nnz;
  int cur_slice;
};

__device__ __forceinline__ int dim_of(const Geo& g, int a) { return a == 0 ? g.W : (a == 1 ? g.H : g.D); }

// thread 0: pose, layout choice and classification bounds of slice `is`
__device__ void build_ctx(const float* __restrict__ tf, const Geo& g, Ctx& cx) {
  float Rv[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) Rv[a][b] = tf[a * 4 + b];
    cx.T[a] = tf[a * 4 + 3];
  }
  // which volume axis do pixel rows (ix) / columns (iy) run along?
  int au = 0, av = 0;
#pragma unroll
  for (int a = 1; a < 3; ++a) {
    if (fabsf(Rv[a][0]) > fabsf(Rv[au][0])) au = a;
    if (fabsf(Rv[a][1]) > fabsf(Rv[av][1])) av = a;
  }
  int perm, fast_iy;
  if (g.tune & kTunePerm) {
    fast_iy = fabsf(Rv[av][1]) > fabsf(Rv[au][0]);
    perm = fast_iy ? av : au;
  } else {
    perm = 0;
    fast_iy = fabsf(Rv[0][1]) > fabsf(Rv[0][0]);
  }
  const float lead = Rv[perm][fast_iy], best = fabsf(lead);
  const float off = sqrtf(fmaxf(0.f, 1.f - best * best));  // drift across the other axes per unit step
  cx.perm = perm;
  cx.fast_iy = fast_iy;
  cx.row_mode = (g.tune & kTuneRow) && off * 31.f * g.res <= 4.f && best * g.res <= 2.f;
  cx.dir = lead >= 0.f ? 1 : -1;
  const int dims[3] = {g.W, g.H, g.D};
  // strides of volume x, y, z in the x- / y- / z-fastest layouts
  const int sv[3][3] = {{1, g.W, g.H * g.W}, {g.H, 1, g.W * g.H}, {g.D, g.W * g.D, 1}};
  const int pw = fast_iy ? g.h : g.w, ps = fast_iy ? g.w : g.h;
  cx.tiles_f = (pw + kTileF - 1) / kTileF;
  cx.tiles_s = (ps + kTileS - 1) / kTileS;
  // bounding box of R.t over the PSF's tap box
  const float tlo[3] = {(float)(-(g.w_p / 2)), (float)(-(g.h_p / 2)), (float)(-(g.d_p / 2))};
  const float thi[3] = {(float)(g.w_p - 1 - g.w_p / 2), (float)(g.h_p - 1 - g.h_p / 2), (float)(g.d_p - 1 - g.d_p / 2)};
  for (int k = 0; k < 3; ++k) {
    const int a = (perm + k) % 3;
    float dmin = 0.f, dmax = 0.f;
    for (int b = 0; b < 3; ++b) {
      cx.R[k][b] = Rv[a][b];
      const float u = Rv[a][b] * tlo[b], v = Rv[a][b] * thi[b];
      dmin += fminf(u, v);
      dmax += fmaxf(u, v);
    }
    const float top = (float)(dims[a] - 1);
    cx.s[k] = sv[perm][a];
    cx.half[k] = (float)((dims[a] - 1) / 2.);
    cx.top[k] = top;
    cx.lo_in[k] = -dmin + kEps;
    cx.hi_in[k] = top - dmax - kEps;
    cx.lo_out[k] = -dmax - kEps;
    cx.hi_out[k] = top - dmin + kEps;
  }
}

// Shared-memory carve-up: Smem header | float4 tap[nnz_max] | float val[ntaps] | int xyz[ntaps]
struct Stage {
  Smem* sm;
  float4* tap;  // {ix_p, iy_p, iz_p, psf value} of the non-zero taps, reference order
  float* val;   // compacted non-zero taps, reference order
  int* xyz;     // packed signed bytes (tx, ty, tz)
};

__device__ __forceinline__ Stage carve(unsigned char* smem, int ntaps) {
  Stage st;
  st.sm = reinterpret_cast<Smem*>(smem);
  st.tap = reinterpret_cast<float4*>(smem + ((sizeof(Smem) + 15) / 16) * 16);
  st.val = reinterpret_cast<float*>(st.tap + ntaps);
  st.xyz = reinterpret_cast<int*>(st.val + ntaps);
  return st;
}

size_t smem_bytes(int ntaps) { return ((sizeof(Smem) + 15) / 16) * 16 + (size_t)ntaps * (16 + 4 + 4); }

__device__ __forceinline__ void unpack_tap(int packed, int t[3]) {
  t[0] = (int)(signed char)(packed & 0xff);
  t[1] = (int)(signed char)((packed >> 8) & 0xff);
  t[2] = (int)(signed char)((packed >> 16) & 0xff);
}

// once per CTA: ordered compaction of the non-zero taps by warp 0 (keeps the reference's summation order)
__device__ void stage_psf(const float* __restrict__ psf, const Geo& g, Stage& st) {
  const int ntaps = g.d_p * g.h_p * g.w_p;
  if (threadIdx.x < 32) {
    int base = 0;
    for (int i0 = 0; i0 < ntaps; i0 += 32) {
      const int i = i0 + threadIdx.x;
      const float v = i < ntaps ? psf[i] : 0.f;
      const bool keep = v != 0.f;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int pos = base + __popc(m & ((1u << threadIdx.x) - 1));
        const int tx = i % g.w_p - g.w_p / 2, ty = (i / g.w_p) % g.h_p - g.h_p / 2, tz = i / (g.w_p * g.h_p) - g.d_p / 2;
        st.val[pos] = v;
        st.xyz[pos] = (tx & 0xff) | ((ty & 0xff) << 8) | ((tz & 0xff) << 16);
      }
      base += __popc(m);
    }
    __syncwarp();
    for (int i = threadIdx.x; i < base; i += 32) {
      int t[3];
      unpack_tap(st.xyz[i], t);
      st.tap[i] = make_float4((float)t[0], (float)t[1], (float)t[2], st.val[i]);
    }
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < base; ++i) s += st.val[i];  // same order as the tap loops: identical to their sum
      st.sm->wsum = s;
      st.sm->nnz = base;
      st.sm->cur_slice = -1;
    }
  }
  __syncthreads();
}


// CTA-uniform: make `is` the current slice (pose, layout choice, classification bounds)
__device__ void enter_slice(int is, const float* __restrict__ transforms, const Geo& g, Stage& st) {
  if (st.sm->cur_slice == is) return;
  __syncthreads();  // everybody is done with the previous slice's context
  if (threadIdx.x == 0) {
    build_ctx(transforms + (size_t)is * 12, g, st.sm->ctx);
    st.sm->cur_slice = is;
  }
  __syncthreads();
}

// position of tap `tp` of a pixel, copy axis order; the reference's expression and association (slice_acq_cuda_kernel.cu:66-68)
#define NSV_TAP_POS(p, px, R, tp)                                                            \
  const float p[3] = {px.c[0] + R[0][0] * tp.x + R[0][1] * tp.y + R[0][2] * tp.z,            \
                      px.c[1] + R[1][0] * tp.x + R[1][1] * tp.y + R[1][2] * tp.z,            \
                      px.c[2] + R[2][0] * tp.x + R[2][1] * tp.y + R[2][2] * tp.z}

struct Pixel {
  int ix, iy;
  long idx;
  float s[3], c[3];  // slice-frame position (voxel units); volume-space position of the PSF centre in COPY axis order
  bool inside;       // inside the slice
  int cls;           // 0: every tap outside the volume, 1: border, 2: every tap inside
};

// lane -> pixel of tile `tile` of the current slice
__device__ __forceinline__ void locate(int is, int tile, const Geo& g, const Ctx& cx, Pixel& px) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tf = tile % cx.tiles_f, ts = tile / cx.tiles_f;
  int lf, ls;
  if (cx.row_mode) {
    lf = lane;
    ls = warp;
  } else {
    lf = (warp & 3) * 8 + (lane & 7);
    ls = (warp >> 2) * 4 + (lane >> 3);
  }
  const int pf = tf * kTileF + lf, ps = ts * kTileS + ls;
  px.ix = cx.fast_iy ? ps : pf;
  px.iy = cx.fast_iy ? pf : ps;
  px.inside = px.ix < g.w && px.iy < g.h;
  px.idx = ((long)is * g.h + px.iy) * g.w + px.ix;
  px.s[0] = (float)((px.ix - (g.w - 1) / 2.) * g.res + cx.T[0]);  // Q8: double, then narrowed
  px.s[1] = (float)((px.iy - (g.h - 1) / 2.) * g.res + cx.T[1]);
  px.s[2] = cx.T[2];
  bool all_in = true, out = false;
#pragma unroll
  for (int a = 0; a < 3; ++a) {  // copy axes
    const float v = cx.R[a][0] * px.s[0] + cx.R[a][1] * px.s[1] + cx.R[a][2] * px.s[2];
    px.c[a] = (float)((double)v + (double)cx.half[a]);  // x_center += (W - 1) / 2. (double, then narrowed)
    all_in = all_in && px.c[a] >= cx.lo_in[a] && px.c[a] <= cx.hi_in[a];
    out = out || px.c[a] < cx.lo_out[a] || px.c[a] > cx.hi_out[a];
  }
  px.cls = (g.tune & kTuneClassify) ? (out ? 0 : (all_in ? 2 : 1)) : 1;
}

// floor + fraction with full-rate instructions (FRND / F2I run at quarter rate); exact for |p| < 2^22
__device__ __forceinline__ void floor_frac(float p, int& g, float& w) {
  const float t = __fadd_rd(p, 12582912.f);
  g = __float_as_int(t) - 0x4B400000;
  w = p - (t - 12582912.f);
}

struct Cell {
  int base;          // element index of corner (0,0,0) in the chosen copy
  float w[3];        // fractional position
  __device__ __forceinline__ float wt(int c) const {
    return ((c & 1) ? w[0] : 1.f - w[0]) * (((c >> 1) & 1) ? w[1] : 1.f - w[1]) * ((c >> 2) ? w[2] : 1.f - w[2]);
  }
};

__device__ __forceinline__ Cell make_cell(const float p[3], const int s[3]) {
  Cell c;
  int g0, g1, g2;
  floor_frac(p[0], g0, c.w[0]);
  floor_frac(p[1], g1, c.w[1]);
  floor_frac(p[2], g2, c.w[2]);
  c.base = g0 * s[0] + g1 * s[1] + g2 * s[2];
  return c;
}

__device__ __forceinline__ bool in_bounds(const float p[3], const float top[3]) {
  return !(p[0] < 0 || p[1] < 0 || p[2] < 0 || p[0] >= top[0] || p[1] >= top[1] || p[2] >= top[2]);  // Q5
}

__device__ __forceinline__ int corner_off(int c, const int s[3]) { return (c & 1) * s[0] + ((c >> 1) & 1) * s[1] + (c >> 2) * s[2]; }

__device__ __forceinline__ void load8(const float* __restrict__ v, int base, const int s[3], float f[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) f[c] = __ldg(v + base + corner_off(c, s));
}

__device__ __forceinline__ float trilerp(const float f[8], const float w[3]) {
  const float a00 = fmaf(w[0], f[1] - f[0], f[0]), a10 = fmaf(w[0], f[3] - f[2], f[2]);
  const float a01 = fmaf(w[0], f[5] - f[4], f[4]), a11 = fmaf(w[0], f[7] - f[6], f[6]);
  const float b0 = fmaf(w[1], a10 - a00, a00), b1 = fmaf(w[1], a11 - a01, a01);
  return fmaf(w[2], b1 - b0, b0);
}

// d/dp of (trilinear interpolation of f) : g[a] += scale * d/dp_a
__device__ __forceinline__ void trigrad(const float f[8], const float w[3], float scale, float g[3]) {
  const float x00 = f[1] - f[0], x10 = f[3] - f[2], x01 = f[5] - f[4], x11 = f[7] - f[6];
  const float gx0 = fmaf(w[1], x10 - x00, x00), gx1 = fmaf(w[1], x11 - x01, x01);
  g[0] = fmaf(scale, fmaf(w[2], gx1 - gx0, gx0), g[0]);
  const float a00 = fmaf(w[0], x00, f[0]), a10 = fmaf(w[0], x10, f[2]), a01 = fmaf(w[0], x01, f[4]), a11 = fmaf(w[0], x11, f[6]);
  const float y0 = a10 - a00, y1 = a11 - a01;
  g[1] = fmaf(scale, fmaf(w[2], y1 - y0, y0), g[1]);
  const float b0 = fmaf(w[1], y0, a00), b1 = fmaf(w[1], y1, a01);
  g[2] = fmaf(scale, b1 - b0, g[2]);
}

// normalisation weight of the scatter / backward passes (Q3: in-bounds taps, vol_mask ignored)
__device__ __forceinline__ float border_weight(const Pixel& px, const Stage& st, const float top[3]) {
  const float (*R)[3] = st.sm->ctx.R;
  float weight = 0.f;
  const int nnz = st.sm->nnz;
  for (int i = 0; i < nnz; ++i) {
    const float4 tp = st.tap[i];
    NSV_TAP_POS(p, px, R, tp);
    if (in_bounds(p, top)) weight += tp.w;
  }
  return weight;
}

// pose-gradient accumulator (12 terms): dL/dR[a][b] += d[a] q[b], dL/dT[b] += (R^T d)[b], q = slice-frame tap position.
// d and the rows of R are in copy order; block_reduce_tf puts the rows back in volume order.
struct TfGrad {
  float g[12];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < 12; ++k) g[k] = 0.f;
  }
  __device__ __forceinline__ void add(const Ctx& cx, const Pixel& px, const float d[3], const int t[3]) {
    const float q[3] = {px.s[0] + t[0], px.s[1] + t[1], px.s[2] + t[2]};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) g[r * 4 + c] = fmaf(d[r], q[c], g[r * 4 + c]);
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c * 4 + 3] += d[0] * cx.R[0][c] + d[1] * cx.R[1][c] + d[2] * cx.R[2][c];
  }
};

__device__ void block_reduce_tf(const TfGrad& acc, float* __restrict__ grad_tf_slice, int perm) {
  // a thread's 12 sums cover the taps of ONE pixel (fp32, like the reference's per-thread accumulators); everything across
  // pixels -- where the cancellation is -- is summed in double and rounded once per CTA
  __shared__ double s_part[kThreads / 32][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    double v = (double)acc.g[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double v = 0.;
#pragma unroll
    for (int wv = 0; wv < kThreads / 32; ++wv) v += s_part[wv][threadIdx.x];
    // entry (k, c < 3) is dL/dR of copy row k = volume row (perm + k) % 3; entries (c, 3) are dL/dT[c] (slice frame)
    const int k = threadIdx.x >> 2, c = threadIdx.x & 3;
    const int out = c == 3 ? threadIdx.x : ((perm + k) % 3) * 4 + c;
    if (v != 0.) atomicAdd(grad_tf_slice + out, (float)v);
  }
  __syncthreads();
}

// ---- scatter of one tap's 8 corner values (a[c] into dst0, optionally b[c] into dst1) ----
// Row mode: the neighbour lane `lane - dir` sits one voxel lower along the fastest axis when the slice is grid-aligned;
// its +1 corners are this lane's +0 corners, so it hands them over and only the last lane of a run issues them.
// Convergent: every lane of the warp calls this for every tap (`valid` false = nothing to scatter).
template <bool TWO>
__device__ __forceinline__ void scatter8(float* __restrict__ dst0, float* __restrict__ dst1, int base, const int s[3], bool valid,
                                         float a[8], float b[8], bool merge, int dir) {
  if (merge) {
    const int lane = threadIdx.x & 31;
    const int src = lane - dir, dst = lane + dir;
    constexpr int fbit = 1;  // corner bit of the fastest (stride-1) axis in copy order
    const int nb_base = __shfl_sync(0xffffffffu, valid ? base : INT_MIN, src & 31);
    const bool take = valid && src >= 0 && src < 32 && nb_base != INT_MIN && nb_base + 1 == base;
    const bool gave = __shfl_sync(0xffffffffu, (int)take, dst & 31) != 0 && dst >= 0 && dst < 32;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (!(c & fbit)) continue;
      const float ra = __shfl_sync(0xffffffffu, a[c], src & 31);
      if (take) a[c ^ fbit] += ra;
      if (TWO) {
        const float rb = __shfl_sync(0xffffffffu, b[c], src & 31);
        if (take) b[c ^ fbit] += rb;
      }
    }
    if (!valid) return;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if ((c & fbit) && gave) continue;
      const int iv = base + corner_off(c, s);
      red_add(dst0 + iv, a[c]);
      if (TWO) red_add(dst1 + iv, b[c]);
    }
  } else {
    if (!valid) return;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int iv = base + corner_off(c, s);
      red_add(dst0 + iv, a[c]);
      if (TWO) red_add(dst1 + iv, b[c]);
    }
  }
}

// The tile loop shared by the four operators: `body(is, px)` runs convergently for all 256 threads of every tile that
// is not entirely outside the volume.
template <typename Body>
__device__ __forceinline__ void for_each_tile(const float* __restrict__ transforms, const float* __restrict__ psf, const Geo& g,
                                              Stage& st, Body body) {
  stage_psf(psf, g, st);
  const long n_patches = (long)g.n * g.tiles_max;
  for (long patch = blockIdx.x; patch < n_patches; patch += gridDim.x) {
    const int is = (int)(patch / g.tiles_max), tile = (int)(patch % g.tiles_max);
    enter_slice(is, transforms, g, st);
    const Ctx& cx = st.sm->ctx;
    if (tile >= cx.tiles_f * cx.tiles_s) continue;  // CTA-uniform
    Pixel px;
    locate(is, tile, g, cx, px);
    body(is, px);
  }
}

// ------------------------------------------------------------------------------------- forward (A)
__global__ void __launch_bounds__(kThreads)
    forward_kernel(const float* __restrict__ transforms, Vols vol, const uint8_t* __restrict__ vol_mask,
                   const uint8_t* __restrict__ slices_mask, const float* __restrict__ psf, float* __restrict__ slices,
                   float* __restrict__ slices_weight, Geo g) {
  extern __shared__ __align__(16) unsigned char smem[];
  Stage st = carve(smem, g.d_p * g.h_p * g.w_p);
  for_each_tile(transforms, psf, g, st, [&](int is, const Pixel& px) {
    const Ctx& cx = st.sm->ctx;
    const bool active = px.inside && px.cls != 0 && !(slices_mask && !slices_mask[px.idx]);
    if (!__any_sync(0xffffffffu, active)) return;
    const int s[3] = {1, cx.s[1], cx.s[2]};
      const float top[3] = {cx.top[0], cx.top[1], cx.top[2]};
      const float R[3][3] = {{cx.R[0][0], cx.R[0][1], cx.R[0][2]}, {cx.R[1][0], cx.R[1][1], cx.R[1][2]}, {cx.R[2][0], cx.R[2][1], cx.R[2][2]}};
    const float* __restrict__ v = vol.at(cx.perm);
    const int nnz = st.sm->nnz;
    float val = 0.f, weight = 0.f;
    if (!vol_mask) {
      if (__all_sync(0xffffffffu, !active || px.cls == 2)) {
        if (active) {
          for (int i = 0; i < nnz; ++i) {
            const float4 tp = st.tap[i];
            NSV_TAP_POS(p, px, R, tp);
            const Cell cell = make_cell(p, s);
            float f[8];
            load8(v, cell.base, s, f);
            val = fmaf(tp.w, trilerp(f, cell.w), val);
          }
          weight = st.sm->wsum;
        }
      } else if (active) {
        for (int i = 0; i < nnz; ++i) {
          const float4 tp = st.tap[i];
          NSV_TAP_POS(p, px, R, tp);
          if (!in_bounds(p, top)) continue;
          const Cell cell = make_cell(p, s);
          float f[8];
          load8(v, cell.base, s, f);
          val = fmaf(tp.w, trilerp(f, cell.w), val);
          weight += tp.w;
        }
      }
    } else if (active) {  // masked voxels drop out of value AND weight, corner by corner (layout: the caller's, perm == 0)
      for (int i = 0; i < nnz; ++i) {
        const float4 tp = st.tap[i];
        NSV_TAP_POS(p, px, R, tp);
        if (!in_bounds(p, top)) continue;
        const Cell cell = make_cell(p, s);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int iv = cell.base + corner_off(c, s);
          if (!vol_mask[iv]) continue;
          const float pw = cell.wt(c) * tp.w;
          val = fmaf(pw, __ldg(v + iv), val);
          weight += pw;
        }
      }
    }
    if (active && weight > 0.f) {  // Q1
      slices[px.idx] = val / weight;
      if (slices_weight) slices_weight[px.idx] = weight;
    }
  });
}

// ------------------------------------------------------------------------------ backward of A
__global__ void __launch_bounds__(kThreads)
    backward_kernel(const float* __restrict__ transforms, Vols vol, const uint8_t* __restrict__ vol_mask,
                    const float* __restrict__ psf, const float* __restrict__ grad_slices, const uint8_t* __restrict__ slices_mask,
                    Vols grad_vol, float* __restrict__ grad_tf, Geo g) {
  extern __shared__ __align__(16) unsigned char smem[];
  Stage st = carve(smem, g.d_p * g.h_p * g.w_p);
  const bool want_vol = grad_vol.p[0] != nullptr;
  for_each_tile(transforms, psf, g, st, [&](int is, const Pixel& px) {
    const Ctx& cx = st.sm->ctx;
    TfGrad acc;
    acc.clear();
    bool active = px.inside && px.cls != 0 && !(slices_mask && !slices_mask[px.idx]);
    float gs = 0.f;
    if (active) {
      gs = grad_slices[px.idx];
      active = gs != 0.f;  // Q2
    }
    if (__any_sync(0xffffffffu, active)) {
      const int s[3] = {1, cx.s[1], cx.s[2]};
      const float top[3] = {cx.top[0], cx.top[1], cx.top[2]};
      const float R[3][3] = {{cx.R[0][0], cx.R[0][1], cx.R[0][2]}, {cx.R[1][0], cx.R[1][1], cx.R[1][2]}, {cx.R[2][0], cx.R[2][1], cx.R[2][2]}};
      const float* __restrict__ v = vol.at(cx.perm);
      float* __restrict__ gv = grad_vol.at(cx.perm);
      const int nnz = st.sm->nnz;
      const bool merge = cx.row_mode && (g.tune & kTuneMerge) && want_vol && !vol_mask;
      if (active) {
        const float weight = px.cls == 2 ? st.sm->wsum : border_weight(px, st, top);
        active = weight != 0.f;
        if (active) gs /= weight;
      }
      for (int i = 0; i < nnz; ++i) {
        const float4 tp = st.tap[i];
        NSV_TAP_POS(p, px, R, tp);
        const bool valid = active && (px.cls == 2 || in_bounds(p, top));
        if (!merge && !valid) continue;
        Cell cell;
        cell.base = 0;
        cell.w[0] = cell.w[1] = cell.w[2] = 0.f;
        if (valid) cell = make_cell(p, s);
        const float tg = tp.w * gs;
        if (want_vol) {
          if (!vol_mask) {
            float a[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] = cell.wt(c) * tg;
            scatter8<false>(gv, nullptr, cell.base, s, valid, a, a, merge, cx.dir);
          } else if (valid) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int iv = cell.base + corner_off(c, s);
              if (vol_mask[iv]) red_add(gv + iv, cell.wt(c) * tg);
            }
          }
        }
        if (grad_tf && valid) {
          float f[8];
          if (!vol_mask) {
            load8(v, cell.base, s, f);
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int iv = cell.base + corner_off(c, s);
              f[c] = vol_mask[iv] ? __ldg(v + iv) : 0.f;  // a masked corner contributes nothing (slice_acq_cuda_kernel.cu:394-450)
            }
          }
          float d[3] = {0.f, 0.f, 0.f};
          trigrad(f, cell.w, tg, d);
          int t[3];
          unpack_tap(st.xyz[i], t);
          acc.add(cx, px, d, t);
        }
      }
    }
    if (grad_tf) block_reduce_tf(acc, grad_tf + (size_t)is * 12, cx.perm);
  });
}

// ------------------------------------------------------------------------- adjoint forward (A^T)
__global__ void __launch_bounds__(kThreads)
    adjoint_forward_kernel(const float* __restrict__ transforms, Vols vol, Vols vol_weight, const uint8_t* __restrict__ vol_mask,
                           const float* __restrict__ psf, const float* __restrict__ slices, const uint8_t* __restrict__ slices_mask,
                           Geo g) {
  extern __shared__ __align__(16) unsigned char smem[];
  Stage st = carve(smem, g.d_p * g.h_p * g.w_p);
  const bool two = vol_weight.p[0] != nullptr;
  for_each_tile(transforms, psf, g, st, [&](int is, const Pixel& px) {
    const Ctx& cx = st.sm->ctx;
    bool active = px.inside && px.cls != 0 && !(slices_mask && !slices_mask[px.idx]);
    float sv = 0.f;
    if (active) {
      sv = slices[px.idx];
      if (!two && (g.tune & kTuneZeroSkip) && sv == 0.f) active = false;  // adds 0 to every voxel it touches
    }
    if (!__any_sync(0xffffffffu, active)) return;
    const int s[3] = {1, cx.s[1], cx.s[2]};
      const float top[3] = {cx.top[0], cx.top[1], cx.top[2]};
      const float R[3][3] = {{cx.R[0][0], cx.R[0][1], cx.R[0][2]}, {cx.R[1][0], cx.R[1][1], cx.R[1][2]}, {cx.R[2][0], cx.R[2][1], cx.R[2][2]}};
    float* __restrict__ dv = vol.at(cx.perm);
    float* __restrict__ dw = two ? vol_weight.at(cx.perm) : nullptr;
    const int nnz = st.sm->nnz;
    const bool merge = cx.row_mode && (g.tune & kTuneMerge) && !vol_mask;
    float inv_w = 0.f;
    if (active) {
      const float weight = px.cls == 2 ? st.sm->wsum : border_weight(px, st, top);
      active = !(weight < 0.5f);  // Q4
      inv_w = 1.f / weight;
    }
    for (int i = 0; i < nnz; ++i) {
      const float4 tp = st.tap[i];
      NSV_TAP_POS(p, px, R, tp);
      const bool valid = active && (px.cls == 2 || in_bounds(p, top));
      if (!merge && !valid) continue;
      Cell cell;
      cell.base = 0;
      cell.w[0] = cell.w[1] = cell.w[2] = 0.f;
      if (valid) cell = make_cell(p, s);
      const float tn = tp.w * inv_w;
      if (!vol_mask) {
        float a[8], b[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          b[c] = cell.wt(c) * tn;
          a[c] = b[c] * sv;
        }
        if (two) scatter8<true>(dv, dw, cell.base, s, valid, a, b, merge, cx.dir);
        else scatter8<false>(dv, nullptr, cell.base, s, valid, a, b, merge, cx.dir);
      } else if (valid) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int iv = cell.base + corner_off(c, s);
          if (!vol_mask[iv]) continue;
          const float pw = cell.wt(c) * tn;
          red_add(dv + iv, pw * sv);
          if (two) red_add(dw + iv, pw);
        }
      }
    }
  });
}

// ------------------------------------------------------------------------------ backward of A^T
__global__ void __launch_bounds__(kThreads)
    adjoint_backward_kernel(const float* __restrict__ transforms, Vols grad_vol, const float* __restrict__ psf,
                            const float* __restrict__ slices, const uint8_t* __restrict__ slices_mask, Vols resid,
                            const uint8_t* __restrict__ vol_mask, float* __restrict__ grad_slices, float* __restrict__ grad_tf, Geo g) {
  extern __shared__ __align__(16) unsigned char smem[];
  Stage st = carve(smem, g.d_p * g.h_p * g.w_p);
  const bool has_resid = resid.p[0] != nullptr;
  for_each_tile(transforms, psf, g, st, [&](int is, const Pixel& px) {
    const Ctx& cx = st.sm->ctx;
    TfGrad acc;
    acc.clear();
    const bool active = px.inside && px.cls != 0 && !(slices_mask && !slices_mask[px.idx]);
    if (active) {
      const int s[3] = {1, cx.s[1], cx.s[2]};
      const float top[3] = {cx.top[0], cx.top[1], cx.top[2]};
      const float R[3][3] = {{cx.R[0][0], cx.R[0][1], cx.R[0][2]}, {cx.R[1][0], cx.R[1][1], cx.R[1][2]}, {cx.R[2][0], cx.R[2][1], cx.R[2][2]}};
      const float* __restrict__ gv = grad_vol.at(cx.perm);
      const float* __restrict__ rv = has_resid ? resid.at(cx.perm) : nullptr;
      const int nnz = st.sm->nnz;
      const float sval = slices[px.idx];
      float val = 0.f, weight = 0.f;
      for (int i = 0; i < nnz; ++i) {
        const float4 tp = st.tap[i];
        NSV_TAP_POS(p, px, R, tp);
        if (px.cls != 2 && !in_bounds(p, top)) continue;
        const Cell cell = make_cell(p, s);
        float f[8];
        if (!vol_mask) {
          load8(gv, cell.base, s, f);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int iv = cell.base + corner_off(c, s);
            f[c] = vol_mask[iv] ? __ldg(gv + iv) : 0.f;
          }
        }
        if (grad_slices) val = fmaf(tp.w, trilerp(f, cell.w), val);
        if (grad_tf) {
          if (has_resid) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int iv = cell.base + corner_off(c, s);
              f[c] *= (!vol_mask || vol_mask[iv]) ? sval - __ldg(rv + iv) : 0.f;
            }
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) f[c] *= sval;
          }
          float d[3] = {0.f, 0.f, 0.f};
          trigrad(f, cell.w, tp.w, d);
          int t[3];
          unpack_tap(st.xyz[i], t);
          acc.add(cx, px, d, t);
        }
        weight += tp.w;
      }
      if (weight > 0.f) {
        if (grad_slices) grad_slices[px.idx] = val / weight;
        const float iw = 1.f / weight;
#pragma unroll
        for (int k = 0; k < 12; ++k) acc.g[k] *= iw;
      } else {
        acc.clear();
      }
    }
    if (grad_tf) block_reduce_tf(acc, grad_tf + (size_t)is * 12, cx.perm);
  });
}

// ---- layout passes: batched 2-D transposes between the caller's x-fastest layout and the y- / z-fastest copies ----
// out[b * ob + c * oc + r] (op)= in[b * ib + r * ir + c],  r < R, c < C, b < B  (c contiguous in `in`, r contiguous in `out`)
template <bool ACC>
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C, int B,
                                                        long ib, long ir, long ob, long oc) {
  __shared__ float tile[32][33];
  const int tc = (C + 31) / 32, tr = (R + 31) / 32;
  const long n_tiles = (long)B * tc * tr;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int b = (int)(t / (tc * tr)), rem = (int)(t % (tc * tr));
    const int r0 = (rem / tc) * 32, c0 = (rem % tc) * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + ly + 8 * k, c = c0 + lx;
      if (r < R && c < C) tile[ly + 8 * k][lx] = in[b * ib + r * ir + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ly + 8 * k, r = r0 + lx;
      if (r < R && c < C) {
        float* o = out + b * ob + c * oc + r;
        if (ACC) *o += tile[lx][ly + 8 * k];
        else *o = tile[lx][ly + 8 * k];
      }
    }
    __syncthreads();
  }
}

int tr_grid(const Geo& g) {
  const long tiles = (long)g.D * ((g.H + 31) / 32) * ((g.W + 31) / 32);
  const long cap = (long)num_sms() * 8;
  return (int)(tiles < cap ? tiles : cap);
}

// x-fastest -> y-fastest (idx = z W H + x H + y) and z-fastest (idx = y W D + x D + z)
int make_copies(const float* src, float* cy, float* cz, const Geo& g, cudaStream_t st) {
  const long HW = (long)g.H * g.W;
  transpose_kernel<false><<<tr_grid(g), 256, 0, st>>>(src, cy, g.H, g.W, g.D, HW, g.W, HW, g.H);              // b = z, r = y, c = x
  transpose_kernel<false><<<tr_grid(g), 256, 0, st>>>(src, cz, g.D, g.W, g.H, g.W, HW, (long)g.W * g.D, g.D);  // b = y, r = z, c = x
  return check_launch("nsv_slice_acq(layout copies)");
}

// dst (x-fastest) += the y-fastest and z-fastest accumulators
int merge_copies(float* dst, const float* ay, const float* az, const Geo& g, cudaStream_t st) {
  const long HW = (long)g.H * g.W;
  transpose_kernel<true><<<tr_grid(g), 256, 0, st>>>(ay, dst, g.W, g.H, g.D, HW, g.H, HW, g.W);               // b = z, r = x, c = y
  transpose_kernel<true><<<tr_grid(g), 256, 0, st>>>(az, dst, g.W, g.D, g.H, (long)g.W * g.D, g.D, g.W, HW);   // b = y, r = x, c = z
  return check_launch("nsv_slice_acq(layout merge)");
}

// stream-ordered scratch: `count` volume-sized float buffers (no host synchronisation)
struct Scratch {
  float* base = nullptr;
  cudaStream_t st;
  size_t each;
  int alloc(int count, const Geo& g, cudaStream_t stream, bool zero) {
    st = stream;
    each = ((size_t)g.D * g.H * g.W + 63) / 64 * 64;
    static bool pool_ready = false;
    if (!pool_ready) {  // keep freed scratch in the device's default pool across synchronisations (the default gives it back)
      int dev = 0;
      cudaMemPool_t pool;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_ready = true;
    }
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&base), each * count * sizeof(float), st);
    if (e == cudaSuccess && zero) e = cudaMemsetAsync(base, 0, each * count * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("nsv_slice_acq: scratch allocation of %zu bytes failed: %s", each * count * sizeof(float), cudaGetErrorString(e));
      cudaGetLastError();
      return (int)e;
    }
    return NSV_OK;
  }
  float* at(int i) const { return base + each * i; }
  ~Scratch() {
    if (base) cudaFreeAsync(base, st);
  }
};

// default: everything but the neighbour merge -- measured on the BASELINE config-2 stacks (profiles/r02_kernelB_vs_reference.json)
// the merge's 5 extra shuffles per tap cost more than the reductions it saves once row warps already put a warp's
// reductions into 4-5 sectors (A^T 0.87 vs 1.18 ms, backward 1.53 vs 1.90 ms); it stays available as tuning bit 4
unsigned g_tune = kTunePerm | kTuneRow | kTuneZeroSkip | kTuneClassify;

int check_geo(const char* name, const Geo& d) {
  NSV_REQUIRE(d.D > 0 && d.H > 0 && d.W > 0 && d.d_p > 0 && d.h_p > 0 && d.w_p > 0 && d.n >= 0 && d.h > 0 && d.w > 0,
              "%s: non-positive dimension", name);
  NSV_REQUIRE((long)d.d_p * d.h_p * d.w_p <= kMaxTaps, "%s: PSF larger than %d taps", name, kMaxTaps);
  NSV_REQUIRE(d.d_p <= 255 && d.h_p <= 255 && d.w_p <= 255, "%s: PSF extent above 255", name);
  NSV_REQUIRE((long)d.D * d.H * d.W < (1L << 31) && (long)d.n * d.h * d.w < (1L << 31),
              "%s: int32 flat index overflow (same limit as the reference, slice_acq_cuda_kernel.cu:33-34)", name);
  return NSV_OK;
}

Geo make_geo(int D, int H, int W, int d_p, int h_p, int w_p, int n, int h, int w, float res, bool masked) {
  Geo g{D, H, W, d_p, h_p, w_p, n, h, w, res, 0, g_tune};
  const int t0 = ((w + kTileF - 1) / kTileF) * ((h + kTileS - 1) / kTileS), t1 = ((h + kTileF - 1) / kTileF) * ((w + kTileS - 1) / kTileS);
  g.tiles_max = t0 > t1 ? t0 : t1;
  if (masked) g.tune &= ~kTunePerm;  // vol_mask is indexed in the caller's layout
  return g;
}

int grid_for(const Geo& g) {
  const long patches = (long)g.n * g.tiles_max;
  const long cap = (long)num_sms() * 8;
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

Vols vols_of(float* x, float* y, float* z) {
  Vols v;
  v.p[0] = x;
  v.p[1] = y;
  v.p[2] = z;
  return v;
}

int run_forward(const float* transforms, const float* vol, const uint8_t* vol_mask, const uint8_t* slices_mask, const float* psf,
                float* slices, float* slices_weight, Geo g, cudaStream_t st) {
  if (int e = check_geo("nsv_slice_acq_forward", g)) return e;
  if (g.n == 0) return NSV_OK;
  NSV_REQUIRE(transforms && vol && psf && slices, "nsv_slice_acq_forward: NULL pointer");
  const size_t smem = smem_bytes(g.d_p * g.h_p * g.w_p);
  if (int e = prep(forward_kernel, smem)) return e;
  Scratch sc;
  Vols v = vols_of(const_cast<float*>(vol), nullptr, nullptr);
  if (g.tune & kTunePerm) {
    if (int e = sc.alloc(2, g, st, false)) return e;
    v = vols_of(const_cast<float*>(vol), sc.at(0), sc.at(1));
    if (int e = make_copies(vol, sc.at(0), sc.at(1), g, st)) return e;
  }
  forward_kernel<<<grid_for(g), kThreads, smem, st>>>(transforms, v, vol_mask, slices_mask, psf, slices, slices_weight, g);
  return check_launch("nsv_slice_acq_forward");
}

int run_backward(const float* transforms, const float* vol, const uint8_t* vol_mask, const float* psf, const float* grad_slices,
                 const uint8_t* slices_mask, float* grad_vol, float* grad_tf, Geo g, cudaStream_t st) {
  if (int e = check_geo("nsv_slice_acq_backward", g)) return e;
  if (g.n == 0 || (!grad_vol && !grad_tf)) return NSV_OK;
  NSV_REQUIRE(transforms && vol && psf && grad_slices, "nsv_slice_acq_backward: NULL pointer");
  const size_t smem = smem_bytes(g.d_p * g.h_p * g.w_p);
  if (int e = prep(backward_kernel, smem)) return e;
  Scratch sc;
  Vols v = vols_of(const_cast<float*>(vol), nullptr, nullptr), gv = vols_of(grad_vol, nullptr, nullptr);
  if (g.tune & kTunePerm) {
    if (int e = sc.alloc(4, g, st, false)) return e;
    if (grad_tf) {
      v = vols_of(const_cast<float*>(vol), sc.at(0), sc.at(1));
      if (int e = make_copies(vol, sc.at(0), sc.at(1), g, st)) return e;
    } else {
      v = vols_of(const_cast<float*>(vol), const_cast<float*>(vol), const_cast<float*>(vol));  // never read
    }
    if (grad_vol) {
      cudaMemsetAsync(sc.at(2), 0, 2 * sc.each * sizeof(float), st);
      gv = vols_of(grad_vol, sc.at(2), sc.at(3));
    }
  }
  backward_kernel<<<grid_for(g), kThreads, smem, st>>>(transforms, v, vol_mask, psf, grad_slices, slices_mask, gv, grad_tf, g);
  if (int e = check_launch("nsv_slice_acq_backward")) return e;
  if ((g.tune & kTunePerm) && grad_vol) return merge_copies(grad_vol, sc.at(2), sc.at(3), g, st);
  return NSV_OK;
}

int run_adjoint_forward(const float* transforms, const float* psf, const float* slices, const uint8_t* slices_mask,
                        const uint8_t* vol_mask, float* vol, float* vol_weight, Geo g, int equalize, cudaStream_t st) {
  if (int e = check_geo("nsv_slice_acq_adjoint_forward", g)) return e;
  if (g.n == 0) return NSV_OK;
  NSV_REQUIRE(transforms && psf && slices && vol, "nsv_slice_acq_adjoint_forward: NULL pointer");
  NSV_REQUIRE(!equalize || vol_weight, "nsv_slice_acq_adjoint_forward: equalize needs vol_weight");
  const size_t smem = smem_bytes(g.d_p * g.h_p * g.w_p);
  if (int e = prep(adjoint_forward_kernel, smem)) return e;
  Scratch sc;
  Vols dv = vols_of(vol, nullptr, nullptr), dw = vols_of(equalize ? vol_weight : nullptr, nullptr, nullptr);
  if (g.tune & kTunePerm) {
    if (int e = sc.alloc(equalize ? 4 : 2, g, st, true)) return e;
    dv = vols_of(vol, sc.at(0), sc.at(1));
    if (equalize) dw = vols_of(vol_weight, sc.at(2), sc.at(3));
  }
  adjoint_forward_kernel<<<grid_for(g), kThreads, smem, st>>>(transforms, dv, dw, vol_mask, psf, slices, slices_mask, g);
  if (int e = check_launch("nsv_slice_acq_adjoint_forward")) return e;
  if (g.tune & kTunePerm) {
    if (int e = merge_copies(vol, sc.at(0), sc.at(1), g, st)) return e;
    if (equalize)
      if (int e = merge_copies(vol_weight, sc.at(2), sc.at(3), g, st)) return e;
  }
  if (equalize) return sa_fma::equalize_f32(vol, vol_weight, 0, (int64_t)g.D * g.H * g.W, st);
  return NSV_OK;
}

int run_adjoint_backward(const float* transforms, float* grad_vol, const float* vol_weight, const uint8_t* vol_mask, const float* psf,
                         const float* slices, const uint8_t* slices_mask, const float* vol, float* grad_slices, float* grad_tf, Geo g,
                         int equalize, cudaStream_t st) {
  if (int e = check_geo("nsv_slice_acq_adjoint_backward", g)) return e;
  NSV_REQUIRE(grad_vol && (g.n == 0 || (transforms && psf && slices)), "nsv_slice_acq_adjoint_backward: NULL pointer");
  NSV_REQUIRE(!equalize || (vol_weight && vol), "nsv_slice_acq_adjoint_backward: equalize needs vol and vol_weight");
  if (equalize)
    if (int e = sa_fma::equalize_f32(grad_vol, vol_weight, 1, (int64_t)g.D * g.H * g.W, st)) return e;
  if (g.n == 0 || (!grad_slices && !grad_tf)) return NSV_OK;
  const size_t smem = smem_bytes(g.d_p * g.h_p * g.w_p);
  if (int e = prep(adjoint_backward_kernel, smem)) return e;
  Scratch sc;
  Vols gv = vols_of(grad_vol, nullptr, nullptr), rv = vols_of(equalize ? const_cast<float*>(vol) : nullptr, nullptr, nullptr);
  if (g.tune & kTunePerm) {
    const bool need_r = equalize && grad_tf;
    if (int e = sc.alloc(need_r ? 4 : 2, g, st, false)) return e;
    gv = vols_of(grad_vol, sc.at(0), sc.at(1));
    if (int e = make_copies(grad_vol, sc.at(0), sc.at(1), g, st)) return e;
    if (need_r) {
      rv = vols_of(const_cast<float*>(vol), sc.at(2), sc.at(3));
      if (int e = make_copies(vol, sc.at(2), sc.at(3), g, st)) return e;
    } else if (equalize) {
      rv = vols_of(const_cast<float*>(vol), const_cast<float*>(vol), const_cast<float*>(vol));  // never read without grad_tf
    }
  }
  adjoint_backward_kernel<<<grid_for(g), kThreads, smem, st>>>(transforms, gv, psf, slices, slices_mask, rv, vol_mask, grad_slices,
                                                              grad_tf, g);
  return check_launch("nsv_slice_acq_adjoint_backward");
}

int g_exact = -1;  // -1: not yet read from the environment
bool exact_mode() {
  if (g_exact < 0) {
    const char* e = getenv("NSV_SLICE_ACQ_EXACT");
    g_exact = (e && e[0] == '1') ? 1 : 0;
    const char* t = getenv("NSV_SLICE_ACQ_TUNE");
    if (t) g_tune = (unsigned)strtoul(t, nullptr, 0);
  }
  return g_exact == 1;
}

}  // namespace
}  // namespace sa_fast
}  // namespace nsv

extern "C" void nsv_set_slice_acq_exact(int exact) {
  nsv::sa_fast::exact_mode();
  nsv::sa_fast::g_exact = exact ? 1 : 0;
}
extern "C" int nsv_get_slice_acq_exact(void) { return nsv::sa_fast::exact_mode() ? 1 : 0; }
extern "C" void nsv_set_slice_acq_tuning(unsigned bits) {
  nsv::sa_fast::exact_mode();
  nsv::sa_fast::g_tune = bits;
}
extern "C" unsigned nsv_get_slice_acq_tuning(void) {
  nsv::sa_fast::exact_mode();
  return nsv::sa_fast::g_tune;
}

#define NSV_ARGS_DIMS D, H, W, d_p, h_p, w_p, n, h, w

extern "C" int nsv_slice_acq_forward_f32(const float* transforms, const float* vol, const uint8_t* vol_mask, const uint8_t* slices_mask,
                                         const float* psf, float* slices, float* slices_weight, int D, int H, int W, int d_p, int h_p,
                                         int w_p, int n, int h, int w, float res_slice, int interp_psf, void* stream) {
  using namespace nsv;
  if (sa_fast::exact_mode())
    return sa_exact::forward_f32(transforms, vol, vol_mask, slices_mask, psf, slices, slices_weight, NSV_ARGS_DIMS, res_slice, interp_psf, stream);
  if (interp_psf)
    return sa_fma::forward_f32(transforms, vol, vol_mask, slices_mask, psf, slices, slices_weight, NSV_ARGS_DIMS, res_slice, interp_psf, stream);
  return sa_fast::run_forward(transforms, vol, vol_mask, slices_mask, psf, slices, slices_weight,
                              sa_fast::make_geo(NSV_ARGS_DIMS, res_slice, vol_mask != nullptr), (cudaStream_t)stream);
}

extern "C" int nsv_slice_acq_backward_f32(const float* transforms, const float* vol, const uint8_t* vol_mask, const float* psf,
                                          const float* grad_slices, const uint8_t* slices_mask, float* grad_vol, float* grad_transforms,
                                          int D, int H, int W, int d_p, int h_p, int w_p, int n, int h, int w, float res_slice,
                                          int interp_psf, void* stream) {
  using namespace nsv;
  if (sa_fast::exact_mode())
    return sa_exact::backward_f32(transforms, vol, vol_mask, psf, grad_slices, slices_mask, grad_vol, grad_transforms, NSV_ARGS_DIMS, res_slice,
                                  interp_psf, stream);
  if (interp_psf)
    return sa_fma::backward_f32(transforms, vol, vol_mask, psf, grad_slices, slices_mask, grad_vol, grad_transforms, NSV_ARGS_DIMS, res_slice,
                                interp_psf, stream);
  return sa_fast::run_backward(transforms, vol, vol_mask, psf, grad_slices, slices_mask, grad_vol, grad_transforms,
                               sa_fast::make_geo(NSV_ARGS_DIMS, res_slice, vol_mask != nullptr), (cudaStream_t)stream);
}

extern "C" int nsv_slice_acq_adjoint_forward_f32(const float* transforms, const float* psf, const float* slices, const uint8_t* slices_mask,
                                                 const uint8_t* vol_mask, float* vol, float* vol_weight, int D, int H, int W, int d_p,
                                                 int h_p, int w_p, int n, int h, int w, float res_slice, int interp_psf, int equalize,
                                                 void* stream) {
  using namespace nsv;
  if (sa_fast::exact_mode())
    return sa_exact::adjoint_forward_f32(transforms, psf, slices, slices_mask, vol_mask, vol, vol_weight, NSV_ARGS_DIMS, res_slice, interp_psf,
                                         equalize, stream);
  if (interp_psf)
    return sa_fma::adjoint_forward_f32(transforms, psf, slices, slices_mask, vol_mask, vol, vol_weight, NSV_ARGS_DIMS, res_slice, interp_psf,
                                       equalize, stream);
  return sa_fast::run_adjoint_forward(transforms, psf, slices, slices_mask, vol_mask, vol, vol_weight,
                                      sa_fast::make_geo(NSV_ARGS_DIMS, res_slice, vol_mask != nullptr), equalize, (cudaStream_t)stream);
}

extern "C" int nsv_slice_acq_adjoint_backward_f32(const float* transforms, float* grad_vol, const float* vol_weight, const uint8_t* vol_mask,
                                                  const float* psf, const float* slices, const uint8_t* slices_mask, const float* vol,
                                                  float* grad_slices, float* grad_transforms, int D, int H, int W, int d_p, int h_p, int w_p,
                                                  int n, int h, int w, float res_slice, int interp_psf, int equalize, void* stream) {
  using namespace nsv;
  if (sa_fast::exact_mode())
    return sa_exact::adjoint_backward_f32(transforms, grad_vol, vol_weight, vol_mask, psf, slices, slices_mask, vol, grad_slices, grad_transforms,
                                          NSV_ARGS_DIMS, res_slice, interp_psf, equalize, stream);
  if (interp_psf)
    return sa_fma::adjoint_backward_f32(transforms, grad_vol, vol_weight, vol_mask, psf, slices, slices_mask, vol, grad_slices, grad_transforms,
                                        NSV_ARGS_DIMS, res_slice, interp_psf, equalize, stream);
  return sa_fast::run_adjoint_backward(transforms, grad_vol, vol_weight, vol_mask, psf, slices, slices_mask, vol, grad_slices, grad_transforms,
                                       sa_fast::make_geo(NSV_ARGS_DIMS, res_slice, vol_mask != nullptr), equalize, (cudaStream_t)stream);
}

extern "C" int nsv_equalize_f32(float* vol, const float* vol_weight, int is_grad, int64_t DHW, void* stream) {
  return nsv::sa_fast::exact_mode() ? nsv::sa_exact::equalize_f32(vol, vol_weight, is_grad, DHW, stream)
                                    : nsv::sa_fma::equalize_f32(vol, vol_weight, is_grad, DHW, stream);
}

// double precision: the generic kernels (exact flavour on request)
#define NSV_SA_F64(NAME, PARAMS, ARGS) \
  extern "C" int nsv_slice_acq_##NAME##_f64 PARAMS { return nsv::sa_fast::exact_mode() ? nsv::sa_exact::NAME##_f64 ARGS : nsv::sa_fma::NAME##_f64 ARGS; }

NSV_SA_F64(forward,
           (const double* transforms, const double* vol, const uint8_t* vol_mask, const uint8_t* slices_mask, const double* psf, double* slices,
            double* slices_weight, int D, int H, int W, int d_p, int h_p, int w_p, int n, int h, int w, double res_slice, int interp_psf,
            void* stream),
           (transforms, vol, vol_mask, slices_mask, psf, slices, slices_weight, NSV_ARGS_DIMS, res_slice, interp_psf, stream))
NSV_SA_F64(backward,
           (const double* transforms, const double* vol, const uint8_t* vol_mask, const double* psf, const double* grad_slices,
            const uint8_t* slices_mask, double* grad_vol, double* grad_transforms, int D, int H, int W, int d_p, int h_p, int w_p, int n, int h,
            int w, double res_slice, int interp_psf, void* stream),
           (transforms, vol, vol_mask, psf, grad_slices, slices_mask, grad_vol, grad_transforms, NSV_ARGS_DIMS, res_slice, interp_psf, stream))
NSV_SA_F64(adjoint_forward,
           (const double* transforms, const double* psf, const double* slices, const uint8_t* slices_mask, const uint8_t* vol_mask, double* vol,
            double* vol_weight, int D, int H, int W, int d_p, int h_p, int w_p, int n, int h, int w, double res_slice, int interp_psf,
            int equalize, void* stream),
           (transforms, psf, slices, slices_mask, vol_mask, vol, vol_weight, NSV_ARGS_DIMS, res_slice, interp_psf, equalize, stream))
NSV_SA_F64(adjoint_backward,
           (const double* transforms, double* grad_vol, const double* vol_weight, const uint8_t* vol_mask, const double* psf,
            const double* slices, const uint8_t* slices_mask, const double* vol, double* grad_slices, double* grad_transforms, int D, int H,
            int W, int d_p, int h_p, int w_p, int n, int h, int w, double res_slice, int interp_psf, int equalize, void* stream),
           (transforms, grad_vol, vol_weight, vol_mask, psf, slices, slices_mask, vol, grad_slices, grad_transforms, NSV_ARGS_DIMS, res_slice,
            interp_psf, equalize, stream))
extern "C" int nsv_equalize_f64(double* vol, const double* vol_weight, int is_grad, int64_t DHW, void* stream) {
  return nsv::sa_fast::exact_mode() ? nsv::sa_exact::equalize_f64(vol, vol_weight, is_grad, DHW, stream)
                                    : nsv::sa_fma::equalize_f64(vol, vol_weight, is_grad, DHW, stream);
}
