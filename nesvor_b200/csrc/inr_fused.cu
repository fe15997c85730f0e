// inr_fused.cu -- kernel A: one fused NeSVoR training iteration (and the forward-only renderer).
//
// Replaces, in ONE launch, the per-iteration op sequence of the reference
//   NeSVoR.forward -> ax_transform_points -> INR.forward (tcnn HashGrid + tcnn MLP) -> sigma_net ->
//   render / losses / edge regulariser -> autograd backward (MLP bwd, hash-grid scatter, pose VJP)
// (nesvor/nesvor/models.py:260-384, driven by nesvor/nesvor/train.py:183-190), which the reference
// executes as ~100 launches with every [N, *] intermediate round-tripping HBM.
//
// Organisation (sm_100a, persistent: one 256-thread CTA per SM, 8 warps):
//   * a CTA tile = 256 PSF samples = 256/S whole pixels, so the per-pixel mean over S samples and
//     the sample <-> S-1-sample pairing of the image regulariser stay inside the CTA;
//   * phase 0 (thread = sample): pose (Rodrigues in registers), x = R(p + sigma*eps + T), bbox
//     normalisation, L-level hash-grid gather from the fp16 table (L2-resident), trilinear blend in
//     fp32, features parked as fp16 in the CTA's shared tile (the only copy ever made);
//   * phase 1 (warp = 32 samples): density MLP (and sigma MLP) on tensor-core fragments with all
//     weights in shared memory, activations in registers (mlp_mma.cuh);
//   * phase 2 (thread = sample): softplus, per-pixel means, slice scale / variance, losses and their
//     analytic gradients;
//   * phase 3: MLP dgrad/wgrad on fragments (weight gradients live in registers across tiles and
//     are flushed once per CTA), then per-sample scatter of dL/d(features) into the fp32 table
//     gradient with vector reductions (red.global.add.v2.f32), and -- when poses are optimised --
//     dL/dx through the grid, reduced per pixel and pushed through the Rodrigues VJP.
// Nothing but the table gradient, O(weights) and O(slices) values is written to global memory.
#include <stdlib.h>

#include "inr_common.cuh"
#include "mlp_mma.cuh"

namespace nsv {
namespace fused {
namespace {

// Shared-memory plan.  A CTA hosts NG = 256/GR independent groups of GR rows (GR/32 warps); each
// group owns its activation tiles, the CTA shares weights, level table and the fp32 weight-gradient
// accumulators.  With GR = 128 the two groups run out of phase (named barriers), so one group's
// tensor-core phases overlap the other's gather / scatter phases.
template <int W, int DEPTH, bool SIGMA, int GR>
struct Layout {
  static constexpr int NG = kTile / GR, GW = GR / 16;
  static constexpr int ldx = kIn + kPad, ldh = W + kPad, ldg = kOutP + kPad;
  static constexpr bool alias_dx = DEPTH >= 2 && (W + kPad) * 2 >= kLddx * 4;  // dL/d(features) tile reuses the dead H_last slot
  static_assert((size_t)GR * kLddx * 4 <= (size_t)GR * ldh * 2 || !alias_dx, "dX does not fit the aliased slot");
  // ---- CTA-shared fp16 weights (offsets in halves) ----
  static constexpr size_t wd0 = 0;
  static constexpr size_t wdh = wd0 + (size_t)W * ldx;
  static constexpr size_t wdo = wdh + (size_t)(DEPTH - 1) * W * ldh;
  static constexpr size_t ws0 = wdo + (size_t)kOutP * ldh;
  static constexpr size_t wso = ws0 + (SIGMA ? (size_t)W * ldx : 0);
  static constexpr size_t w_halves = wso + (SIGMA ? (size_t)kOutP * ldh : 0);
  // ---- per-group fp16 tiles (offsets in halves from the group's base) ----
  static constexpr size_t sx = 0;
  static constexpr size_t sh = sx + (size_t)GR * ldx;
  static constexpr size_t sg = sh + (size_t)DEPTH * GR * ldh;
  static constexpr size_t ssx = sg + (size_t)GR * ldg;
  static constexpr size_t ssh = ssx + (SIGMA ? (size_t)GR * ldx : 0);
  static constexpr size_t g_halves = (ssh + (SIGMA ? (size_t)GR * ldh : 0) + 7) / 8 * 8;
  // ---- per-group fp32 scratch (offsets in floats from the group's float base) ----
  static constexpr size_t fz0 = 0;                       // z0, later dz0
  static constexpr size_t flv = fz0 + GR;                // log_var
  static constexpr size_t frho = flv + GR;
  static constexpr size_t fxw = frho + GR;               // [GR][3] world coords, later dL/dx rows
  static constexpr size_t fred = fxw + 3 * GR;           // [GW][16] warp partials (GW = GR/16 warps)
  static constexpr size_t fdx = fred + GW * 16;          // [GR][33] dL/d(features) unless aliased
  static constexpr size_t g_floats = fdx + (alias_dx ? 0 : (size_t)GR * kLddx);
  // ---- CTA-shared fp32: level table + weight-gradient accumulators (packed MLP layout) ----
  static constexpr size_t n_density = (size_t)W * kIn + (size_t)(DEPTH - 1) * W * W + (size_t)kOutP * W;
  static constexpr size_t n_sigma = SIGMA ? ((size_t)W * kIn + (size_t)kOutP * W) : 0;
  static constexpr size_t lt_floats = (sizeof(LevelTable) + 3) / 4;
  // ---- byte offsets ----
  static constexpr size_t b_groups = (w_halves * 2 + 15) / 16 * 16;
  static constexpr size_t b_gfloats = b_groups + (size_t)NG * g_halves * 2;
  static constexpr size_t b_lt = (b_gfloats + (size_t)NG * g_floats * 4 + 15) / 16 * 16;
  static constexpr size_t b_acc = b_lt + lt_floats * 4;
  static constexpr size_t bytes = b_acc + (n_density + n_sigma) * 4;
};

template <int W, int DEPTH>
struct RenderLayout {
  static constexpr int ldx = kIn + kPad, ldh = W + kPad;
  static constexpr size_t wd0 = 0;
  static constexpr size_t wdh = wd0 + (size_t)W * ldx;
  static constexpr size_t wdo = wdh + (size_t)(DEPTH - 1) * W * ldh;
  static constexpr size_t sx = wdo + (size_t)kOutP * ldh;
  static constexpr size_t halves = sx + (size_t)kTile * ldx;
  static constexpr size_t f_base = (halves * 2 + 15) / 16 * 16;
  static constexpr size_t fz0 = 0;
  static constexpr size_t flt = fz0 + kTile;
  static_assert(flt % 4 == 0, "LevelTable must stay 16-byte aligned");
  static constexpr size_t bytes = f_base + (flt + (sizeof(LevelTable) + 3) / 4) * 4;
};

// column 0 of a 2-n-tile accumulator -> per-row scalar array (rows of this warp)
template <int MTL>
__device__ __forceinline__ void store_col0(const float (&c)[MTL][2][4], float* dst, int row0) {
  const int lane = threadIdx.x & 31;
  if ((lane & 3) == 0) {
    const int g = lane >> 2;
#pragma unroll
    for (int m = 0; m < MTL; ++m) {
      dst[row0 + m * 16 + g] = c[m][0][0];
      dst[row0 + m * 16 + g + 8] = c[m][0][2];
    }
  }
}

// add a warp's freshly computed dW block into the CTA's shared fp32 accumulator (row-major [OUT][IN])
template <int OUT, int IN, int NW>
__device__ __forceinline__ void wgrad_to_smem(const float (&acc)[WgradSplit<OUT, IN, NW>::NTW][4], float* sacc, int warp_in_group) {
  using S = WgradSplit<OUT, IN, NW>;
  const int lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int mt = warp_in_group % S::MT, part = warp_in_group / S::MT;
  if (part >= S::PARTS) return;
#pragma unroll
  for (int j = 0; j < S::NTW; ++j) {
    const int nt = part * S::NTW + j;
    if (nt >= S::NTL) continue;
    float* p = sacc + (size_t)(mt * 16 + gq) * IN + nt * 8 + 2 * t;
    red_shared(p, acc[j][0]);
    red_shared(p + 1, acc[j][1]);
    red_shared(p + 8 * IN, acc[j][2]);
    red_shared(p + 8 * IN + 1, acc[j][3]);
  }
}

// one layer's weight gradient for this group's tile: registers -> shared accumulator
template <int OUT, int IN, int NW>
__device__ __forceinline__ void group_wgrad(float* sacc, const __half* dc_tile, int ld_dc, const __half* a_tile, int ld_a, int rows,
                                            int warp_in_group) {
  float acc[WgradSplit<OUT, IN, NW>::NTW][4] = {};
  warp_wgrad<OUT, IN, NW>(acc, dc_tile, ld_dc, a_tile, ld_a, rows, warp_in_group);
  wgrad_to_smem<OUT, IN, NW>(acc, sacc, warp_in_group);
}

template <int W, int DEPTH, bool SIGMA, int GR>
__global__ void __launch_bounds__(kThreads, 1) inr_train_kernel(const __grid_constant__ FusedArgs a) {
  using L = Layout<W, DEPTH, SIGMA, GR>;
  constexpr int GW = L::GW, NG = L::NG, GT = GR * 2;  // warps / threads per group
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp / GW, gw = warp % GW, row0 = gw * 16;
  const int xb = lane & 1, srow = row0 + (lane >> 1);  // this lane's sample row inside the group tile
  __half* swt = reinterpret_cast<__half*>(smem_raw);                                           // weights
  __half* sh = reinterpret_cast<__half*>(smem_raw + L::b_groups) + (size_t)grp * L::g_halves;  // this group's tiles
  float* sf = reinterpret_cast<float*>(smem_raw + L::b_gfloats) + (size_t)grp * L::g_floats;
  LevelTable& lt = *reinterpret_cast<LevelTable*>(smem_raw + L::b_lt);
  float* sacc = reinterpret_cast<float*>(smem_raw + L::b_acc);
  float* sacc_sigma = sacc + L::n_density;
  const nsv_inr_config& cfg = a.cfg;

  // ---- stage weights (fp16, padded rows), level table; clear the gradient accumulators ----
  {
    const __half* wd = a.mlp + a.off_density;
    stage_weights(swt + L::wd0, L::ldx, wd, W, kIn);
    for (int l = 0; l + 1 < DEPTH; ++l) stage_weights(swt + L::wdh + (size_t)l * W * L::ldh, L::ldh, wd + (size_t)W * kIn + (size_t)l * W * W, W, W);
    stage_weights(swt + L::wdo, L::ldh, wd + (size_t)W * kIn + (size_t)(DEPTH - 1) * W * W, kOutP, W);
    if (SIGMA) {
      const __half* ws = a.mlp + a.off_sigma;
      stage_weights(swt + L::ws0, L::ldx, ws, W, kIn);
      stage_weights(swt + L::wso, L::ldh, ws + (size_t)W * kIn, kOutP, W);
    }
    for (int i = tid; i < (int)(L::n_density + L::n_sigma); i += kThreads) sacc[i] = 0.f;
    stage_level_table(lt, cfg.grid, tid, a.agg_max, a.fast, a.table, a.g_table);
  }
  // ---- log-sum-exp of logit_coef (slice scale c_k = n_s softmax_k), per warp (n_slices is small) ----
  float lse = 0.f;
  if (cfg.slice_scale) {
    float mx = -INFINITY;
    for (int k = lane; k < a.n_slices; k += 32) mx = fmaxf(mx, a.logit_coef[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int k = lane; k < a.n_slices; k += 32) se += expf(a.logit_coef[k] - mx);
    se = warp_sum(se);
    lse = mx + logf(se);
  }
  __syncthreads();

  float loss_d = 0.f, loss_s = 0.f, loss_i = 0.f;
  const int S = a.S, wpp = S >> 4;  // warps per pixel
  const float invS = 1.f / (float)S, invB = 1.f / (float)a.B, gscale = cfg.grad_scale, inv_gscale = 1.f / cfg.grad_scale;
  const int64_t n_tiles = (a.B * (int64_t)S) / GR;

  for (int64_t tile = (int64_t)blockIdx.x * NG + grp; tile < n_tiles; tile += (int64_t)gridDim.x * NG) {
    group_barrier(grp, GT);  // previous tile's cooperative wgrad reads of this group's tiles are complete
    // ================= phase 0: sample geometry + encoding =================
    const int64_t sidx = tile * GR + srow;
    const int64_t p = sidx >> a.log2S;
    const int j = (int)(sidx & (S - 1));
    const int k = (int)a.slice_idx[p];
    float ax[6], R[9], y[3], xw[3], xn[3];
#pragma unroll
    for (int d = 0; d < 6; ++d) ax[d] = a.axisangle[(size_t)k * 6 + d];
    rodrigues<float>(ax, R);
    {
      float eps[3];
      if (a.noise) {
#pragma unroll
        for (int d = 0; d < 3; ++d) eps[d] = a.noise[sidx * 3 + d];
      } else {
        normal3(a.seed, a.offset + (uint64_t)sidx, eps);
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) y[d] = (a.xyz[p * 3 + d] + eps[d] * a.psf_sigma[(size_t)k * 3 + d]) + ax[3 + d];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        xw[i] = R[i * 3] * y[0] + R[i * 3 + 1] * y[1] + R[i * 3 + 2] * y[2];
        xn[i] = (xw[i] - cfg.bbox_lo[i]) / (cfg.bbox_hi[i] - cfg.bbox_lo[i]);
      }
    }
    bool slow;
    {
      __half* xrow = sh + L::sx + (size_t)srow * L::ldx;
      slow = encode_warp(xn, lt, cfg.grid.n_levels, a.table, [&](int l, __half2 v) { *reinterpret_cast<__half2*>(xrow + 2 * l) = v; },
                         [&](int c, uint4 v) { *reinterpret_cast<uint4*>(xrow + 8 * c) = v; });
    }
    __syncwarp();

    // ================= phase 1: density MLP forward (warp-local, one m16 tile per warp) =================
    uint32_t az[1][1][4];  // z0..z15 as an A fragment (input of sigma_net)
    uint64_t relu_mask[DEPTH], relu_mask_s = 0;  // pre-activation > 0 bits of this thread's fragment elements
    {
      uint32_t ain[1][kIn / 16][4];
      load_a_frags<kIn / 16>(ain, sh + L::sx, L::ldx, row0);
      float c[1][W / 8][4];
      warp_gemm_fwd<kIn / 16, W / 8>(c, ain, swt + L::wd0, L::ldx);
      uint32_t ah[1][W / 16][4];
      relu_mask[0] = relu_bits<W / 8>(c);
      acc_to_a<W / 8, true>(ah, c);
      store_a_frags<W / 16>(ah, sh + L::sh, L::ldh, row0);
#pragma unroll
      for (int l = 1; l < DEPTH; ++l) {
        warp_gemm_fwd<W / 16, W / 8>(c, ah, swt + L::wdh + (size_t)(l - 1) * W * L::ldh, L::ldh);
        relu_mask[l] = relu_bits<W / 8>(c);
        acc_to_a<W / 8, true>(ah, c);
        store_a_frags<W / 16>(ah, sh + L::sh + (size_t)l * GR * L::ldh, L::ldh, row0);
      }
      float co[1][2][4];
      warp_gemm_fwd<W / 16, 2>(co, ah, swt + L::wdo, L::ldh);
      store_col0(co, sf + L::fz0, row0);
      acc_to_a<2, false>(az, co);
    }
    // ================= phase 1b: sigma MLP forward =================
    if (SIGMA) {
      __half* sr = sh + L::ssx + (size_t)srow * L::ldx + 8 * xb;  // [slice embedding (16) | z (16)], 8 halves per lane
      const float* se = a.slice_embedding + (size_t)k * 16 + 8 * xb;
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<__half2*>(sr + 2 * q) = __floats2half2_rn(se[2 * q], se[2 * q + 1]);
      store_a_frags<1>(az, sh + L::ssx + 16, L::ldx, row0);
      __syncwarp();
      uint32_t asx[1][2][4];
      load_a_frags<2>(asx, sh + L::ssx, L::ldx, row0);
      float c[1][W / 8][4];
      warp_gemm_fwd<2, W / 8>(c, asx, swt + L::ws0, L::ldx);
      uint32_t ash[1][W / 16][4];
      relu_mask_s = relu_bits<W / 8>(c);
      acc_to_a<W / 8, true>(ash, c);
      store_a_frags<W / 16>(ash, sh + L::ssh, L::ldh, row0);
      float co[1][2][4];
      warp_gemm_fwd<W / 16, 2>(co, ash, swt + L::wso, L::ldh);
      store_col0(co, sf + L::flv, row0);
    }
    __syncwarp();

    // ================= phase 2: render, losses, gradients w.r.t. z0 / log_var (both lanes of a pair) =================
    const float z0 = sf[L::fz0 + srow];
    const float rho = softplus_f(z0);
    const float lv = SIGMA ? sf[L::flv + srow] : 0.f;
    const float u = SIGMA ? expf(lv) : 1.f;
    if (xb == 0) {
      sf[L::frho + srow] = rho;
      sf[L::fxw + 3 * srow] = xw[0];
      sf[L::fxw + 3 * srow + 1] = xw[1];
      sf[L::fxw + 3 * srow + 2] = xw[2];
    }
    {
      const float s_rho = warp_sum(xb ? 0.f : rho), s_u = warp_sum(xb ? 0.f : u);
      if (lane == 0) {
        sf[L::fred + gw * 2] = s_rho;
        sf[L::fred + gw * 2 + 1] = s_u;
      }
    }
    group_barrier(grp, GT);
    float m_pix = 0.f, q_pix = 0.f;
    {
      const int w0 = (gw / wpp) * wpp;
      for (int q = 0; q < wpp; ++q) {
        m_pix += sf[L::fred + (w0 + q) * 2];
        q_pix += sf[L::fred + (w0 + q) * 2 + 1];
      }
      m_pix *= invS;
      q_pix *= invS;
    }
    const float ck = cfg.slice_scale ? (float)a.n_slices * expf(a.logit_coef[k] - lse) : 1.f;
    const float vhat = ck * m_pix;
    const float r = ck * q_pix;
    float var = SIGMA ? r * r : 1.f;
    const float evs = cfg.slice_variance ? expf(a.log_var_slice[k]) : 0.f;
    var += evs;
    const float e = vhat - a.v[p];
    const float d_vhat = e / var * invB;
    const float d_var = (SIGMA || cfg.slice_variance) ? (0.5f / var - 0.5f * e * e / (var * var)) * invB : 0.f;
    float d_rho = ck * d_vhat * invS;
    const float d_lv = SIGMA ? (u * invS) * ck * 2.f * r * d_var : 0.f;
    if (j == 0 && xb == 0) {
      loss_d += 0.5f * e * e / var * invB;
      if (SIGMA || cfg.slice_variance) loss_s += 0.5f * logf(var) * invB;
      if (a.v_out) a.v_out[p] = vhat;
      if (cfg.slice_scale) red_add(a.g_c + k, m_pix * d_vhat);
      if (cfg.slice_variance) red_add(a.g_lvs + k, evs * d_var);
    }
    if (cfg.image_reg) {
      const int tp = (srow & ~(S - 1)) + (S - 1 - j);
      const float dr = rho - sf[L::frho + tp];
      const float dx0 = xw[0] - sf[L::fxw + 3 * tp], dx1 = xw[1] - sf[L::fxw + 3 * tp + 1], dx2 = xw[2] - sf[L::fxw + 3 * tp + 2];
      const float d2 = dx0 * dx0 + dx1 * dx1 + dx2 * dx2 + 1e-6f;
      const float nbs = invB * invS;
      float li;
      if (cfg.image_reg == 2) {  // edge
        const float sq = sqrtf(1.f + dr * dr / (d2 * cfg.delta * cfg.delta));
        li = sq * nbs;
        d_rho += cfg.w_image * 2.f * dr / (cfg.delta * d2 * sq) * nbs;
      } else if (cfg.image_reg == 1) {  // TV
        const float dd = sqrtf(d2);
        li = fabsf(dr) / dd * nbs;
        d_rho += cfg.w_image * 2.f * (dr > 0.f ? 1.f : (dr < 0.f ? -1.f : 0.f)) / dd * nbs;
      } else {  // L2
        li = dr * dr / d2 * nbs;
        d_rho += cfg.w_image * 4.f * dr / d2 * nbs;
      }
      if (xb == 0) loss_i += li;
    }
    const float dz0 = (z0 > 20.f ? 1.f : sigmoid_f(z0)) * d_rho * gscale;
    __syncwarp();  // both lanes of every pair have read z0
    if (xb == 0) {
      sf[L::fz0 + srow] = dz0;  // z0 is dead
      if (SIGMA) {
        __half* grow = sh + L::sg + (size_t)srow * L::ldg;
        *reinterpret_cast<uint4*>(grow) = make_uint4((uint32_t)__half_as_ushort(__float2half_rn(d_lv * gscale)), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(grow + 8) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    group_barrier(grp, GT);  // [B0] sG / activations of every warp of the group are in place

    // ================= phase 3: backward =================
    float c2[1][2][4];  // dL/dz (16 columns) of the density net, fp32 fragment
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int q = 0; q < 4; ++q) c2[0][n][q] = 0.f;
    if (SIGMA) {
      group_wgrad<kOutP, W, GW>(sacc_sigma + (size_t)W * kIn, sh + L::sg, L::ldg, sh + L::ssh, L::ldh, GR, gw);
      uint32_t ag[1][1][4];
      load_a_frags<1>(ag, sh + L::sg, L::ldg, row0);
      float c[1][W / 8][4];
      warp_gemm_dgrad<1, W / 8>(c, ag, swt + L::wso, L::ldh);
      apply_relu_bits<W / 8>(c, relu_mask_s);
      uint32_t adz[1][W / 16][4];
      acc_to_a<W / 8, false>(adz, c);
      group_barrier(grp, GT);  // [B1]
      store_a_frags<W / 16>(adz, sh + L::ssh, L::ldh, row0);
      group_barrier(grp, GT);  // [B2]
      group_wgrad<W, kIn, GW>(sacc_sigma, sh + L::ssh, L::ldh, sh + L::ssx, L::ldx, GR, gw);
      float cin[1][4][4];
      warp_gemm_dgrad<W / 16, 4>(cin, adz, swt + L::ws0, L::ldx);
      // d(slice embedding): column sums over the warp's 16 rows (one pixel, one slice per warp)
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        float s0 = cin[0][n][0] + cin[0][n][2];
        float s1 = cin[0][n][1] + cin[0][n][3];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (lane < 4) red_add_v2(a.g_se + (size_t)k * 16 + n * 8 + 2 * lane, s0 * inv_gscale, s1 * inv_gscale);
      }
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int q = 0; q < 4; ++q) c2[0][n][q] = cin[0][2 + n][q];
    }
    if ((lane & 3) == 0) {  // add dL/dz0 of the render path to column 0
      const int g = lane >> 2;
      c2[0][0][0] += sf[L::fz0 + row0 + g];
      c2[0][0][2] += sf[L::fz0 + row0 + g + 8];
    }
    uint32_t adz[1][W / 16][4];
    {
      uint32_t ag[1][1][4];
      acc_to_a<2, false>(ag, c2);
      store_a_frags<1>(ag, sh + L::sg, L::ldg, row0);  // sG is free: sigma's wgrad finished before [B1]
      group_barrier(grp, GT);  // [B3]
      __half* sHl = sh + L::sh + (size_t)(DEPTH - 1) * GR * L::ldh;
      group_wgrad<kOutP, W, GW>(sacc + (size_t)W * kIn + (size_t)(DEPTH - 1) * W * W, sh + L::sg, L::ldg, sHl, L::ldh, GR, gw);
      float c[1][W / 8][4];
      warp_gemm_dgrad<1, W / 8>(c, ag, swt + L::wdo, L::ldh);
      apply_relu_bits<W / 8>(c, relu_mask[DEPTH - 1]);
      acc_to_a<W / 8, false>(adz, c);
      group_barrier(grp, GT);  // [B4]
      store_a_frags<W / 16>(adz, sHl, L::ldh, row0);
      group_barrier(grp, GT);  // [B5]
    }
#pragma unroll
    for (int l = DEPTH - 1; l >= 1; --l) {
      __half* sDz = sh + L::sh + (size_t)l * GR * L::ldh;
      __half* sHp = sh + L::sh + (size_t)(l - 1) * GR * L::ldh;
      group_wgrad<W, W, GW>(sacc + (size_t)W * kIn + (size_t)(l - 1) * W * W, sDz, L::ldh, sHp, L::ldh, GR, gw);
      float c[1][W / 8][4];
      warp_gemm_dgrad<W / 16, W / 8>(c, adz, swt + L::wdh + (size_t)(l - 1) * W * L::ldh, L::ldh);
      apply_relu_bits<W / 8>(c, relu_mask[l - 1]);
      acc_to_a<W / 8, false>(adz, c);
      group_barrier(grp, GT);
      store_a_frags<W / 16>(adz, sHp, L::ldh, row0);
      group_barrier(grp, GT);
    }
    group_wgrad<W, kIn, GW>(sacc, sh + L::sh, L::ldh, sh + L::sx, L::ldx, GR, gw);
    // dL/d(features): fp32 rows, either in their own scratch or in the dead H_last slot (DEPTH >= 2:
    // its last reader was the wgrad of layer DEPTH-1, two group barriers ago)
    float* sdx = L::alias_dx ? reinterpret_cast<float*>(sh + L::sh + (size_t)(DEPTH - 1) * GR * L::ldh) : sf + L::fdx;
    {
      float cx[1][kIn / 8][4];
      warp_gemm_dgrad<W / 16, kIn / 8>(cx, adz, swt + L::wd0, L::ldx);
      const int g = lane >> 2, t = lane & 3;
#pragma unroll
      for (int n = 0; n < kIn / 8; ++n) {
        float* d0 = sdx + (size_t)(row0 + g) * kLddx + n * 8 + 2 * t;
        d0[0] = cx[0][n][0];
        d0[1] = cx[0][n][1];
        d0[8 * kLddx] = cx[0][n][2];
        d0[8 * kLddx + 1] = cx[0][n][3];
      }
    }
    __syncwarp();
    // ---- scatter into the table gradient (+ pose gradient) ----
    const float* grow_s = sdx + (size_t)srow * kLddx;
    auto fetch = [&](int l) { return make_float2(grow_s[2 * l], grow_s[2 * l + 1]); };
    if (cfg.pose_grad) {
      float gx[3];
      scatter_warp<true>(xn, lt, cfg.grid.n_levels, a.table, fetch, inv_gscale, a.g_table, gx, slow);
      float gwd[3], part[12];
#pragma unroll
      for (int i = 0; i < 3; ++i) gwd[i] = xb ? 0.f : gx[i] / (cfg.bbox_hi[i] - cfg.bbox_lo[i]);  // one lane per sample contributes
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int q = 0; q < 3; ++q) part[i * 3 + q] = gwd[i] * y[q];  // dL/dR
#pragma unroll
      for (int q = 0; q < 3; ++q) part[9 + q] = R[q] * gwd[0] + R[3 + q] * gwd[1] + R[6 + q] * gwd[2];  // dL/dT = R^T g
#pragma unroll
      for (int q = 0; q < 12; ++q) part[q] = warp_sum(part[q]);
      group_barrier(grp, GT);  // fred is free again (phase-2 reads are long done)
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 12; ++q) sf[L::fred + gw * 16 + q] = part[q];
      }
      group_barrier(grp, GT);
      if (j == 0 && xb == 0) {
        float G[9], gT[3], gwv[3];
#pragma unroll
        for (int q = 0; q < 12; ++q) {
          float sm = 0.f;
          for (int ww = 0; ww < wpp; ++ww) sm += sf[L::fred + (gw + ww) * 16 + q];
          if (q < 9) G[q] = sm; else gT[q - 9] = sm;
        }
        rodrigues_vjp<float>(ax, G, gwv);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          red_add(a.g_axisangle + (size_t)k * 6 + q, gwv[q]);
          red_add(a.g_axisangle + (size_t)k * 6 + 3 + q, gT[q]);
        }
      }
    } else {
      float gx[3];
      scatter_warp<false>(xn, lt, cfg.grid.n_levels, a.table, fetch, inv_gscale, a.g_table, gx, slow);
    }
  }

  // ================= epilogue: flush weight gradients and losses =================
  __syncthreads();
  for (int i = tid; i < (int)L::n_density; i += kThreads) {
    const float v = sacc[i];
    if (v != 0.f) red_add(a.g_mlp + a.off_density + i, v * inv_gscale);
  }
  if (SIGMA)
    for (int i = tid; i < (int)L::n_sigma; i += kThreads) {
      const float v = sacc_sigma[i];
      if (v != 0.f) red_add(a.g_mlp + a.off_sigma + i, v * inv_gscale);
    }
  loss_d = warp_sum(loss_d);
  loss_s = warp_sum(loss_s);
  loss_i = warp_sum(loss_i);
  if (lane == 0) {
    red_add(a.losses + 0, loss_d);
    red_add(a.losses + 1, loss_s);
    red_add(a.losses + 3, loss_i);
  }
}

// ------------------------------------------------------------------------------------- renderer
struct RenderArgs {
  nsv_inr_config cfg;
  const __half* table;
  const __half* mlp;
  int64_t off_density;
  const float* xyz;
  const float* mat;
  int mat_per_point;
  const float* psf_sigma;
  int sigma_per_point;
  const float* noise;
  uint64_t seed, offset;
  float* out;
  int64_t M;
  int S;
};

template <int W, int DEPTH>
__global__ void __launch_bounds__(kThreads, 1) inr_render_kernel(const __grid_constant__ RenderArgs a) {
  using L = RenderLayout<W, DEPTH>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* sh = reinterpret_cast<__half*>(smem_raw);
  float* sf = reinterpret_cast<float*>(smem_raw + L::f_base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row0 = warp * 16;
  const int xb = lane & 1, srow = row0 + (lane >> 1);
  const nsv_inr_config& cfg = a.cfg;
  const __half* wd = a.mlp + a.off_density;
  stage_weights(sh + L::wd0, L::ldx, wd, W, kIn);
  for (int l = 0; l + 1 < DEPTH; ++l) stage_weights(sh + L::wdh + (size_t)l * W * L::ldh, L::ldh, wd + (size_t)W * kIn + (size_t)l * W * W, W, W);
  stage_weights(sh + L::wdo, L::ldh, wd + (size_t)W * kIn + (size_t)(DEPTH - 1) * W * W, kOutP, W);
  LevelTable& lt = *reinterpret_cast<LevelTable*>(sf + L::flt);
  stage_level_table(lt, cfg.grid, tid, 0u, 1, a.table, nullptr);
  __syncthreads();
  const int64_t total = a.M * (int64_t)a.S;
  const int64_t n_tiles = (total + kTile - 1) / kTile;
  const float invS = 1.f / (float)a.S;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t sidx = tile * kTile + srow;
    const bool valid = sidx < total;
    const int64_t p = valid ? sidx / a.S : 0;
    float xn[3] = {0.f, 0.f, 0.f};
    if (valid) {
      float y[3];
      float eps[3] = {0.f, 0.f, 0.f};
      if (a.S > 1 || a.noise) {
        if (a.noise) {
#pragma unroll
          for (int d = 0; d < 3; ++d) eps[d] = a.noise[sidx * 3 + d];
        } else {
          normal3(a.seed, a.offset + (uint64_t)sidx, eps);
        }
      }
      const float* sg = a.psf_sigma + (a.sigma_per_point ? p * 3 : 0);
#pragma unroll
      for (int d = 0; d < 3; ++d) y[d] = a.xyz[p * 3 + d] + eps[d] * sg[d];
      float xw[3] = {y[0], y[1], y[2]};
      if (a.mat) {
        const float* mt = a.mat + (a.mat_per_point ? p * 12 : 0);
#pragma unroll
        for (int i = 0; i < 3; ++i) xw[i] = mt[i * 4] * (y[0] + mt[3]) + mt[i * 4 + 1] * (y[1] + mt[7]) + mt[i * 4 + 2] * (y[2] + mt[11]);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) xn[i] = (xw[i] - cfg.bbox_lo[i]) / (cfg.bbox_hi[i] - cfg.bbox_lo[i]);
    }
    {
      __half* xrow = sh + L::sx + (size_t)srow * L::ldx;
      encode_warp(xn, lt, cfg.grid.n_levels, a.table, [&](int l, __half2 v) { *reinterpret_cast<__half2*>(xrow + 2 * l) = v; },
                         [&](int c, uint4 v) { *reinterpret_cast<uint4*>(xrow + 8 * c) = v; });
    }
    __syncwarp();
    uint32_t ain[1][kIn / 16][4];
    load_a_frags<kIn / 16>(ain, sh + L::sx, L::ldx, row0);
    float c[1][W / 8][4];
    warp_gemm_fwd<kIn / 16, W / 8>(c, ain, sh + L::wd0, L::ldx);
    uint32_t ah[1][W / 16][4];
    acc_to_a<W / 8, true>(ah, c);
#pragma unroll
    for (int l = 1; l < DEPTH; ++l) {
      warp_gemm_fwd<W / 16, W / 8>(c, ah, sh + L::wdh + (size_t)(l - 1) * W * L::ldh, L::ldh);
      acc_to_a<W / 8, true>(ah, c);
    }
    float co[1][2][4];
    warp_gemm_fwd<W / 16, 2>(co, ah, sh + L::wdo, L::ldh);
    store_col0(co, sf + L::fz0, row0);
    __syncwarp();
    float rho = (valid && xb == 0) ? softplus_f(sf[L::fz0 + srow]) * invS : 0.f;
    // warp-aggregate when the whole warp renders one point
    const int64_t p0 = __shfl_sync(0xffffffffu, p, 0), p31 = __shfl_sync(0xffffffffu, p, 31);
    const bool all_valid = __all_sync(0xffffffffu, valid);
    if (all_valid && p0 == p31) {
      rho = warp_sum(rho);
      if (lane == 0) red_add(a.out + p, rho);
    } else if (valid && xb == 0) {
      red_add(a.out + p, rho);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------- host side
struct MlpLayout {
  int64_t off_density, off_sigma, off_bias, total;
};

MlpLayout mlp_layout(const nsv_inr_config& c) {
  MlpLayout l;
  const int64_t W = c.width;
  const int64_t per_density = W * kIn + (int64_t)(c.depth - 1) * W * W + kOutP * W;
  const int64_t per_head = W * kIn + (int64_t)(c.depth - 1) * W * W + kOutP * W;
  l.off_density = 0;
  l.off_sigma = per_density;
  l.off_bias = l.off_sigma + (c.pixel_variance ? per_head : 0);
  l.total = l.off_bias + (c.n_levels_bias ? per_head : 0);
  return l;
}

int validate_common(const char* name, const nsv_inr_config* c) {
  NSV_REQUIRE(c != nullptr, "%s: NULL config", name);
  if (c->grid.n_features != 2 || c->grid.n_levels < 1 || c->grid.n_levels * 2 > kIn) {
    set_error("%s: fused path needs F=2 and n_levels <= 16 (got F=%d, L=%d)", name, c->grid.n_features, c->grid.n_levels);
    return NSV_EUNSUPPORTED;
  }
  if (!(c->width == 64 || c->width == 32) || c->depth < 1 || c->depth > 3) {
    set_error("%s: fused path instantiated for width 32|64, depth 1..3 (got %d, %d)", name, c->width, c->depth);
    return NSV_EUNSUPPORTED;
  }
  return NSV_OK;
}

template <typename K>
int set_dyn_smem(K kernel, size_t bytes, const char* name) {
  if (bytes > 227 * 1024) {
    set_error("%s: configuration needs %zu bytes of shared memory (> 227 KB)", name, bytes);
    return NSV_EUNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    return (int)e;
  }
  return NSV_OK;
}

template <int W, int DEPTH, bool SIGMA, int GR>
int launch_train_gr(const FusedArgs& a, cudaStream_t st) {
  using L = Layout<W, DEPTH, SIGMA, GR>;
  if (int e = set_dyn_smem(inr_train_kernel<W, DEPTH, SIGMA, GR>, L::bytes, "nsv_inr_train_step")) return e;
  const int64_t ctas = (a.B * (int64_t)a.S / GR + L::NG - 1) / L::NG;
  const int grid = (int)(ctas < num_sms() ? ctas : num_sms());
  inr_train_kernel<W, DEPTH, SIGMA, GR><<<grid, kThreads, L::bytes, st>>>(a);
  if (int e = check_launch("nsv_inr_train_step")) return e;
  inr_finalize_kernel<<<1, 256, 0, st>>>(a.logit_coef, a.g_c, a.losses, a.n_slices, a.cfg.slice_scale, a.cfg.image_reg, a.cfg.delta);
  return check_launch("nsv_inr_train_step(finalize)");
}

// pixels of up to 128 samples: two independent 128-row groups per CTA; 256 samples: one 256-row group
template <int W, int DEPTH, bool SIGMA>
int launch_train(const FusedArgs& a, cudaStream_t st) {
  if (a.S <= 128) return launch_train_gr<W, DEPTH, SIGMA, 128>(a, st);
  return launch_train_gr<W, DEPTH, SIGMA, 256>(a, st);
}

template <int W, int DEPTH>
int launch_render(const RenderArgs& a, cudaStream_t st) {
  using L = RenderLayout<W, DEPTH>;
  if (int e = set_dyn_smem(inr_render_kernel<W, DEPTH>, L::bytes, "nsv_inr_render")) return e;
  const int64_t tiles = (a.M * (int64_t)a.S + kTile - 1) / kTile;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  inr_render_kernel<W, DEPTH><<<grid, kThreads, L::bytes, st>>>(a);
  return check_launch("nsv_inr_render");
}

}  // namespace
}  // namespace fused
}  // namespace nsv

namespace nsv { namespace fused {
static int g_fused_impl = 0;
static long g_agg_max = -1;  // -1: NSV_AGG_MAX or the built-in default
static int g_fast_path = -1;
static int g_smem_levels = -2;  // -2: NSV_SMEM_LEVELS or "as many as fit" (-1)
static int g_tile_order = -1;   // -1: NSV_TILE_ORDER or 0 (strided)
static int g_tc_groups = -1;    // -1: NSV_TC_GROUPS or the built-in default
static long long* g_timers = nullptr;
} }

extern "C" int nsv_set_fused_timers(void* device_counters) {
  nsv::fused::g_timers = (long long*)device_counters;
  return NSV_OK;
}

// ---- verification hook: the PSF-noise generator of the fused kernels, exposed sample by sample ----
namespace nsv {
namespace fused {
static __global__ void debug_normal3_kernel(uint64_t seed, uint64_t offset, int64_t n, float* __restrict__ normals, uint32_t* __restrict__ raw) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t idx = offset + (uint64_t)i;
    if (normals) {
      float e[3];
      normal3(seed, idx, e);
      normals[3 * i] = e[0];
      normals[3 * i + 1] = e[1];
      normals[3 * i + 2] = e[2];
    }
    if (raw) {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      raw[4 * i] = r.x;
      raw[4 * i + 1] = r.y;
      raw[4 * i + 2] = r.z;
      raw[4 * i + 3] = r.w;
    }
  }
}
}  // namespace fused
}  // namespace nsv

extern "C" int nsv_debug_normal3(uint64_t seed, uint64_t offset, int64_t n, float* normals, uint32_t* raw, void* stream) {
  NSV_REQUIRE(n >= 0 && (n == 0 || normals || raw), "nsv_debug_normal3: bad arguments");
  if (n == 0) return NSV_OK;
  const int64_t blocks = (n + 255) / 256;
  nsv::fused::debug_normal3_kernel<<<(int)(blocks < 1184 ? blocks : 1184), 256, 0, (cudaStream_t)stream>>>(seed, offset, n, normals, raw);
  return nsv::check_launch("nsv_debug_normal3");
}

extern "C" int nsv_set_fused_tc_groups(int groups) {
  NSV_REQUIRE(groups == -1 || groups == 2 || groups == 3, "nsv_set_fused_tc_groups: 2, 3 or -1 (default)");
  nsv::fused::g_tc_groups = groups;
  return NSV_OK;
}

extern "C" int nsv_set_fused_tile_order(int contiguous) {
  nsv::fused::g_tile_order = contiguous < 0 ? -1 : (contiguous != 0);
  return NSV_OK;
}

extern "C" int nsv_set_fused_smem_levels(int levels) {
  nsv::fused::g_smem_levels = levels < -1 ? -2 : levels;
  return NSV_OK;
}

extern "C" int nsv_set_fused_tuning(int64_t agg_max_entries, int fast_path) {
  nsv::fused::g_agg_max = agg_max_entries < 0 ? -1 : (long)agg_max_entries;
  nsv::fused::g_fast_path = fast_path < 0 ? -1 : (fast_path != 0);
  return NSV_OK;
}

extern "C" int nsv_set_fused_impl(int impl) {
  if (impl < 0 || impl > 3) {
    nsv::set_error("nsv_set_fused_impl: 0 = auto, 1 = mma.sync, 2 = tcgen05, 3 = tcgen05 warp-specialised");
    return NSV_EINVAL;
  }
  nsv::fused::g_fused_impl = impl;
  return NSV_OK;
}

extern "C" int64_t nsv_inr_mlp_layout(const nsv_inr_config* cfg, int64_t* offsets) {
  if (!cfg) {
    nsv::set_error("nsv_inr_mlp_layout: NULL config");
    return NSV_EINVAL;
  }
  const nsv::fused::MlpLayout l = nsv::fused::mlp_layout(*cfg);
  if (offsets) {
    offsets[0] = l.off_density;
    offsets[1] = l.off_sigma;
    offsets[2] = l.off_bias;
  }
  return l.total;
}

extern "C" int nsv_inr_train_step(const nsv_inr_config* cfg, const nsv_inr_params* prm, const nsv_inr_grads* g, const float* xyz,
                                  const float* v, const int64_t* slice_idx, const float* noise, uint64_t seed, uint64_t offset,
                                  float* v_out, int64_t B, int S, void* stream) {
  using namespace nsv;
  using namespace nsv::fused;
  if (int e = validate_common("nsv_inr_train_step", cfg)) return e;
  NSV_REQUIRE(prm && g && xyz && v && slice_idx, "nsv_inr_train_step: NULL pointer");
  NSV_REQUIRE(prm->table_f16 && prm->mlp_f16 && prm->axisangle && prm->psf_sigma && g->table && g->mlp && g->losses,
              "nsv_inr_train_step: NULL parameter / gradient buffer");
  NSV_REQUIRE(B > 0, "nsv_inr_train_step: B must be positive");
  int log2S = 0;
  while ((1 << log2S) < S) ++log2S;
  if (!(S >= 32 && S <= kTile && (1 << log2S) == S) || (B * (int64_t)S) % (S <= 128 ? 128 : kTile) != 0) {
    set_error("nsv_inr_train_step: fused path needs n_samples in {32,64,128,256} and B*S a multiple of the 128/256-sample tile (got B=%lld S=%d)", (long long)B, S);
    return NSV_EUNSUPPORTED;
  }
  const bool bias_head = cfg->n_levels_bias != 0;
  if (bias_head && (cfg->n_levels_bias < 0 || cfg->n_levels_bias > 4 || cfg->n_levels_bias > cfg->grid.n_levels || !cfg->pixel_variance ||
                    cfg->width != 64 || g_fused_impl == 1 || g_fused_impl == 3)) {
    set_error("nsv_inr_train_step: the fused bias-field head needs 1 <= n_levels_bias <= 4, the sigma_net heads on (pixel variance, "
              "depth 1, n_features_slice 16), width 64 and the tcgen05 all-phases kernel");
    return NSV_EUNSUPPORTED;
  }
  if (cfg->pixel_variance && (cfg->depth != 1 || cfg->n_features_slice != 16 || cfg->n_features_z != 15)) {
    set_error("nsv_inr_train_step: fused sigma_net needs depth 1, n_features_slice 16, n_features_z 15");
    return NSV_EUNSUPPORTED;
  }
  NSV_REQUIRE(!cfg->pixel_variance || prm->slice_embedding, "nsv_inr_train_step: pixel variance needs slice_embedding");
  NSV_REQUIRE(!cfg->pixel_variance || g->slice_embedding, "nsv_inr_train_step: pixel variance needs grads.slice_embedding");
  NSV_REQUIRE(!cfg->slice_scale || (prm->logit_coef && g->slice_scale_c), "nsv_inr_train_step: slice scale needs logit_coef and its gradient");
  NSV_REQUIRE(!cfg->slice_variance || (prm->log_var_slice && g->log_var_slice), "nsv_inr_train_step: slice variance needs log_var_slice and its gradient");
  NSV_REQUIRE(!cfg->pose_grad || g->axisangle, "nsv_inr_train_step: pose_grad needs grads.axisangle");
  NSV_REQUIRE(cfg->grad_scale > 0.f, "nsv_inr_train_step: grad_scale must be positive");
  const MlpLayout ml = mlp_layout(*cfg);
  FusedArgs a;
  a.cfg = *cfg;
  a.table = (const __half*)prm->table_f16;
  a.mlp = (const __half*)prm->mlp_f16;
  a.axisangle = prm->axisangle;
  a.psf_sigma = prm->psf_sigma;
  a.slice_embedding = prm->slice_embedding;
  a.logit_coef = prm->logit_coef;
  a.log_var_slice = prm->log_var_slice;
  a.n_slices = prm->n_slices;
  a.g_table = g->table;
  a.g_mlp = g->mlp;
  a.g_axisangle = g->axisangle;
  a.g_se = g->slice_embedding;
  a.g_c = g->slice_scale_c;
  a.g_lvs = g->log_var_slice;
  a.losses = g->losses;
  a.xyz = xyz;
  a.v = v;
  a.slice_idx = slice_idx;
  a.noise = noise;
  a.seed = seed;
  a.offset = offset;
  a.v_out = v_out;
  a.B = B;
  a.S = S;
  a.log2S = log2S;
  a.off_density = ml.off_density;
  a.off_sigma = ml.off_sigma;
  a.off_bias = ml.off_bias;
  {  // tuning knobs (read once): NSV_AGG_MAX = largest dense level (entries) whose gradient is pre-reduced per warp
    static const long agg_env = getenv("NSV_AGG_MAX") ? atol(getenv("NSV_AGG_MAX")) : (long)kAggDefault;
    static const int fast_env = getenv("NSV_FAST_PATH") ? atoi(getenv("NSV_FAST_PATH")) : 1;
    a.agg_max = (uint32_t)(g_agg_max >= 0 ? g_agg_max : (agg_env < 0 ? 0 : agg_env));
    a.fast = g_fast_path >= 0 ? g_fast_path : fast_env;
    static const int smem_env = getenv("NSV_SMEM_LEVELS") ? atoi(getenv("NSV_SMEM_LEVELS")) : -1;
    a.smem_levels = g_smem_levels > -2 ? g_smem_levels : smem_env;
    a.smem_table_bytes = 0;
    static const int order_env = getenv("NSV_TILE_ORDER") ? atoi(getenv("NSV_TILE_ORDER")) : 0;
    a.tile_order = g_tile_order >= 0 ? g_tile_order : order_env;
    static const int groups_env = getenv("NSV_TC_GROUPS") ? atoi(getenv("NSV_TC_GROUPS")) : 2;
    a.tc_groups = g_tc_groups >= 2 ? g_tc_groups : groups_env;
    a.timers = g_timers;
    a.ablate = getenv("NSV_ABLATE") ? (uint32_t)atoi(getenv("NSV_ABLATE")) : 0u;  // profiling only, tcgen05 kernel
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool sig = cfg->pixel_variance != 0;
  // warp-specialised variant: measured faster where the MLP chain is short and wide in heads (reference defaults:
  // depth 1 + sigma_net, 0.667 vs 0.771 ms at 2^20 queries), slower on the 3-hidden-layer config-2 model (0.74 vs 0.72 ms;
  // profiles/r01_phase_breakdown.md) -> auto picks it for the sigma_net instantiation only
  if (bias_head) return launch_train_tc(a, st);  // the one instantiation of the bias-field head (losses[4] = mean log_bias, see header)
  if (g_fused_impl == 3 || (g_fused_impl == 0 && sig)) {
    const int rc = launch_train_ws(a, st);
    if (rc != NSV_EUNSUPPORTED || g_fused_impl == 3) return rc;
  }
  if (g_fused_impl != 1) {  // tcgen05 / TMEM implementation when it is instantiated for this configuration
    const int rc = launch_train_tc(a, st);
    if (rc != NSV_EUNSUPPORTED || g_fused_impl == 2) return rc;
  }
  if (cfg->width == 64) {
    if (sig) return launch_train<64, 1, true>(a, st);
    if (cfg->depth == 1) return launch_train<64, 1, false>(a, st);
    if (cfg->depth == 2) return launch_train<64, 2, false>(a, st);
    return launch_train<64, 3, false>(a, st);
  }
  if (sig) return launch_train<32, 1, true>(a, st);
  if (cfg->depth == 1) return launch_train<32, 1, false>(a, st);
  if (cfg->depth == 2) return launch_train<32, 2, false>(a, st);
  return launch_train<32, 3, false>(a, st);
}

extern "C" int nsv_inr_render(const nsv_inr_config* cfg, const nsv_inr_params* prm, const float* xyz, const float* mat,
                              int mat_per_point, const float* psf_sigma, int sigma_per_point, const float* noise, uint64_t seed,
                              uint64_t offset, float* out, int64_t M, int S, void* stream) {
  using namespace nsv;
  using namespace nsv::fused;
  if (int e = validate_common("nsv_inr_render", cfg)) return e;
  NSV_REQUIRE(prm && prm->table_f16 && prm->mlp_f16, "nsv_inr_render: NULL parameters");
  NSV_REQUIRE(M >= 0 && S >= 1, "nsv_inr_render: bad sizes");
  if (M == 0) return NSV_OK;
  NSV_REQUIRE(xyz && out && (psf_sigma || (S == 1 && !noise)), "nsv_inr_render: NULL pointer");
  static const float zero_sigma[3] = {0.f, 0.f, 0.f};
  (void)zero_sigma;
  RenderArgs a;
  a.cfg = *cfg;
  a.table = (const __half*)prm->table_f16;
  a.mlp = (const __half*)prm->mlp_f16;
  a.off_density = mlp_layout(*cfg).off_density;
  a.xyz = xyz;
  a.mat = mat;
  a.mat_per_point = mat_per_point;
  a.psf_sigma = psf_sigma ? psf_sigma : xyz;  // never dereferenced meaningfully when S == 1 without noise (eps = 0)
  a.sigma_per_point = psf_sigma ? sigma_per_point : 0;
  a.noise = noise;
  a.seed = seed;
  a.offset = offset;
  a.out = out;
  a.M = M;
  a.S = S;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)M * sizeof(float), st);
  if (e != cudaSuccess) {
    set_error("nsv_inr_render: cudaMemsetAsync: %s", cudaGetErrorString(e));
    return (int)e;
  }
  if (cfg->width == 64) {
    if (cfg->depth == 1) return launch_render<64, 1>(a, st);
    if (cfg->depth == 2) return launch_render<64, 2>(a, st);
    return launch_render<64, 3>(a, st);
  }
  if (cfg->depth == 1) return launch_render<32, 1>(a, st);
  if (cfg->depth == 2) return launch_render<32, 2>(a, st);
  return launch_render<32, 3>(a, st);
}
