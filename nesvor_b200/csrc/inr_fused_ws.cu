// inr_fused_ws.cu -- kernel A, warp-specialised: the tcgen05 / TMEM implementation of inr_fused_tc.cu with
// the two kinds of work given to different warps so that they overlap instead of alternating.
//
// Measured on the non-specialised kernel (profiles/r01_phase_breakdown.md): a warp spends 62 % of its
// cycles in the gather / scatter phases (instruction-issue and LSU bound, needs all 16 warps to fill
// the SM) and 38 % in the MMA -> epilogue -> barrier chain of the MLPs (latency bound: ~1000 cycles per
// MMA step with nothing to do), and because both 128-sample groups of a CTA run the same program they
// stay in lockstep -- the chain's latency is never hidden.  Here a CTA (one per SM, persistent) has
//   * 16 "memory" warps (2 streams x 8 warps, lane pair = sample as before, 88 registers): for their
//     stream's tiles they run  gather(t+1), scatter(t), gather(t+2), scatter(t+1), ...  and never touch
//     the tensor core;
//   * 8 "chain" warps (2 streams x 4 warps, thread = TMEM lane = sample row, 64 registers): MMA issue,
//     epilogues (ReLU / ReLU mask / fp16 rounding), render + losses, for tile t while the memory warps
//     already gather tile t+1 and scatter tile t-1.
// Hand-over is by mbarriers in shared memory: x_full / x_empty for the double-buffered encoded-feature
// tile (+ the samples' world coordinates for the regulariser), dx_full / dx_empty for dL/d(features).
// x_empty is signalled by the tensor core itself (tcgen05.commit of the tile's last MMAs).
// Register budget (setmaxnreg): launched with 80 registers x 768 threads = 61440 = 512 x 88 + 256 x 64.
// Everything else -- operand layouts, TMEM-resident weight gradients, loss math, gather / scatter
// loops, reference op sequence (nesvor/nesvor/models.py:260-384) -- is shared with inr_fused_tc.cu.
// Instantiated for width 64, depth 1..3 (+ sigma_net at depth 1), n_samples in {32, 64, 128, 256}.
#include "inr_common.cuh"
#include "umma.cuh"

namespace nsv {
namespace fused {
namespace {

constexpr int kW = 64, kGR = 128, kNStreams = 2;
constexpr int kMemWarps = 16, kChainWarps = 8;
constexpr int kMemThreads = kMemWarps * 32, kWsThreads = (kMemWarps + kChainWarps) * 32;
constexpr int kChainGT = 128;  // threads of one stream's chain warpgroup

template <int DEPTH, bool SIGMA>
struct WsLayout {
  // ---- CTA-shared canonical weight tiles (byte offsets) ----
  static constexpr size_t w0 = 0;                                         // [64][32]
  static constexpr size_t wh = w0 + 64 * 32 * 2;                          // (DEPTH-1) x [64][64]
  static constexpr size_t wo = wh + (size_t)(DEPTH - 1) * 64 * 64 * 2;    // [16][64]
  static constexpr size_t ws0 = wo + 16 * 64 * 2;                         // [64][32]
  static constexpr size_t wso = ws0 + (SIGMA ? 64 * 32 * 2 : 0);          // [16][64]
  static constexpr size_t w_end = wso + (SIGMA ? 16 * 64 * 2 : 0);
  // ---- per-stream tiles (byte offsets from the stream base) ----
  static constexpr size_t tx = 0;                                         // 2 x [128][32] encoded features (double buffer)
  static constexpr size_t th = tx + 2 * 128 * 32 * 2;                     // DEPTH x [128][64] hidden, later dZ
  static constexpr size_t tg = th + (size_t)DEPTH * 128 * 64 * 2;         // [128][16] dL/dz
  static constexpr size_t tsx = tg + 128 * 16 * 2;                        // [128][32] sigma_net input
  static constexpr size_t tsh = tsx + (SIGMA ? 128 * 32 * 2 : 0);         // [128][64] sigma_net hidden
  static constexpr size_t dx = tsh + (SIGMA ? 128 * 64 * 2 : 0);          // [128][32] fp32 dL/d(features), XOR-swizzled
  static constexpr size_t s_bytes = dx + 128 * 32 * 4;
  // ---- CTA-level fp32 scratch, indexed by CTA row (stream * 128 + row) ----
  static constexpr size_t b_streams = (w_end + 127) / 128 * 128;
  static constexpr size_t b_scr = b_streams + kNStreams * s_bytes;
  static constexpr size_t frho = 0, fxw = 256, fred = fxw + 2 * 768, fend = fred + 16;  // floats; fxw: [2 buffers][256 rows][3]
  static constexpr size_t b_lt = b_scr + fend * 4;
  static constexpr size_t b_sync = (b_lt + sizeof(LevelTable) + 15) / 16 * 16;
  // mbarriers (8 bytes each), per stream s: mma[s], x_full[s][2], x_empty[s][2], dx_full[s], dx_empty[s]
  static constexpr int m_mma = 0, m_xfull = 2, m_xempty = 6, m_dxfull = 10, m_dxempty = 12, n_mbar = 14;
  static constexpr size_t b_misc = b_sync + n_mbar * 8;                    // tmem slot, stream-ran flags
  static constexpr size_t bytes = b_misc + 32;
  // ---- TMEM columns ----
  static constexpr uint32_t c_d = 0;                                      // stream s: [64 s, 64 s + 64)
  static constexpr uint32_t c_w0 = 128;                                   // dW0   [64 x 32]
  static constexpr uint32_t c_wh = c_w0 + 32;                             // dWh_l [64 x 64]
  static constexpr uint32_t c_wo = c_wh + 64 * (DEPTH - 1);               // dWo^T [64 x 16]
  static constexpr uint32_t c_ws0 = c_wo + 16;                            // dWs0  [64 x 32]
  static constexpr uint32_t c_wso = c_ws0 + 32;                           // dWso^T [64 x 16]
  static constexpr uint32_t c_end = c_wso + 16;
  static_assert(c_end <= 512, "TMEM columns");
};

template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::saddr(mbar)) : "memory");
}
__device__ __forceinline__ void chain_barrier(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }
__device__ __forceinline__ void chain_barrier_all() { asm volatile("bar.sync 3, 256;" ::: "memory"); }

// ---- epilogues: this thread owns TMEM lane (= sample row) `row`; 32 accumulator columns starting at c0 ----
__device__ __forceinline__ void ws_relu_store(uint32_t taddr, unsigned char* tile, int row, int c0) {
  uint32_t r[32];
  umma::tmem_ld32(taddr, r);
  umma::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 v;
    uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __half2 h = __floats2half2_rn(fmaxf(__uint_as_float(r[8 * i + 2 * q]), 0.f), fmaxf(__uint_as_float(r[8 * i + 2 * q + 1]), 0.f));
      pv[q] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(tile + umma::tile_off(row, c0 + 8 * i, 64)) = v;
  }
}
// dA = D masked by (parked activation > 0), rounded to fp16, written over the activation (in place)
__device__ __forceinline__ void ws_mask_store(uint32_t taddr, unsigned char* tile, int row, int c0) {
  uint32_t r[32];
  umma::tmem_ld32(taddr, r);
  umma::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4* p = reinterpret_cast<uint4*>(tile + umma::tile_off(row, c0 + 8 * i, 64));
    const uint4 hv = *p;
    const uint32_t* ph = reinterpret_cast<const uint32_t*>(&hv);
    uint4 v;
    uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&ph[q]));
      const float d0 = h.x > 0.f ? __uint_as_float(r[8 * i + 2 * q]) : 0.f;
      const float d1 = h.y > 0.f ? __uint_as_float(r[8 * i + 2 * q + 1]) : 0.f;
      const __half2 o = __floats2half2_rn(d0, d1);
      pv[q] = *reinterpret_cast<const uint32_t*>(&o);
    }
    *p = v;
  }
}
__device__ __forceinline__ uint4 pack8(const float* g) {
  uint4 v;
  uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __half2 h = __floats2half2_rn(g[2 * q], g[2 * q + 1]);
    pv[q] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return v;
}

// TIMED (profiling builds only, a.timers != NULL): per-phase warp cycles -> a.timers[0..15]:
//   memory warps: 0 gather, 1 wait x_empty, 2 scatter, 3 wait dx_full
//   chain warps : 8 wait x_full, 9 MMA issue -> mbarrier, 10 epilogues, 11 chain barriers, 12 render + losses, 13 wait dx_empty
template <int DEPTH, bool SIGMA, bool POSE, bool TIMED = false>
__global__ void __launch_bounds__(kWsThreads, 1) inr_train_ws_kernel(const __grid_constant__ FusedArgs a) {
  using L = WsLayout<DEPTH, SIGMA>;
  long long t_acc[TIMED ? 6 : 1] = {};
  long long t_last = 0;
  auto tick = [&](int seg) {
    if (TIMED) {
      const long long t = clock64();
      t_acc[seg] += t - t_last;
      t_last = t;
    }
  };
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* wt = smem;
  float* sf = reinterpret_cast<float*>(smem + L::b_scr);
  LevelTable& lt = *reinterpret_cast<LevelTable*>(smem + L::b_lt);
  uint64_t* mbars = reinterpret_cast<uint64_t*>(smem + L::b_sync);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::b_misc);
  uint32_t* grp_ran = reinterpret_cast<uint32_t*>(smem + L::b_misc + 8);
  const nsv_inr_config& cfg = a.cfg;

  // ---- one-time setup: weights -> canonical tiles, level table, TMEM, mbarriers ----
  {
    const __half* wd = a.mlp + a.off_density;
    umma::stage_tile(wt + L::w0, wd, 64, 32, tid, kWsThreads);
    for (int l = 0; l + 1 < DEPTH; ++l) umma::stage_tile(wt + L::wh + (size_t)l * 64 * 64 * 2, wd + 64 * 32 + (size_t)l * 64 * 64, 64, 64, tid, kWsThreads);
    umma::stage_tile(wt + L::wo, wd + 64 * 32 + (size_t)(DEPTH - 1) * 64 * 64, 16, 64, tid, kWsThreads);
    if (SIGMA) {
      const __half* ws = a.mlp + a.off_sigma;
      umma::stage_tile(wt + L::ws0, ws, 64, 32, tid, kWsThreads);
      umma::stage_tile(wt + L::wso, ws + 64 * 32, 16, 64, tid, kWsThreads);
    }
    stage_level_table(lt, cfg.grid, tid, a.agg_max, a.fast, a.table, a.g_table, a.ablate);
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
      for (int s = 0; s < kNStreams; ++s) {
        umma::mbar_init(mbars + L::m_mma + s, 1);
        for (int b = 0; b < 2; ++b) {
          umma::mbar_init(mbars + L::m_xfull + 2 * s + b, 8);   // one arrive per memory warp of the stream
          umma::mbar_init(mbars + L::m_xempty + 2 * s + b, 1);  // tcgen05.commit
        }
        umma::mbar_init(mbars + L::m_dxfull + s, 4);            // one arrive per chain warp of the stream
        umma::mbar_init(mbars + L::m_dxempty + s, 8);
      }
      umma::mbar_fence_init();
      grp_ran[0] = grp_ran[1] = 0;
    }
  }
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const int S = a.S;
  const int64_t n_tiles = (a.B * (int64_t)S) / kGR;
  const int64_t stride = (int64_t)gridDim.x * kNStreams;
  const float inv_gscale = 1.f / cfg.grad_scale;

  if (warp < kMemWarps) {
    // =====================================================================================================
    // memory warps: geometry + hash-grid gather for tile t+1, gradient scatter for tile t
    // =====================================================================================================
    reg_inc<88>();
    const int grp = warp >> 3, gw = warp & 7;
    const int xb = lane & 1, srow = gw * 16 + (lane >> 1);  // lane pair = one sample row
    const int crow = grp * kGR + srow;
    unsigned char* gt = smem + L::b_streams + (size_t)grp * L::s_bytes;
    uint64_t* x_full = mbars + L::m_xfull + 2 * grp;
    uint64_t* x_empty = mbars + L::m_xempty + 2 * grp;
    uint64_t* dx_full = mbars + L::m_dxfull + grp;
    uint64_t* dx_empty = mbars + L::m_dxempty + grp;
    const float* sdx = reinterpret_cast<const float*>(gt + L::dx);
    const int64_t first = (int64_t)blockIdx.x * kNStreams + grp;
    const int64_t n_my = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;

    struct Geo {
      float xn[3], y[POSE ? 3 : 1];  // y = p + sigma * eps + T and the slice are kept for the pose VJP only
      int k;
      bool slow;
    };
    auto gather = [&](int64_t tl, Geo& g) {
      const int b = (int)(tl & 1);
      tick(2);
      if (tl >= 2) umma::mbar_wait(x_empty + b, (uint32_t)(((tl >> 1) - 1) & 1));  // chain(tl - 2) has released the buffer
      tick(1);
      const int64_t tile = first + tl * stride;
      const int64_t sidx = tile * kGR + srow;
      const int64_t p = sidx >> a.log2S;
      const int k = (int)a.slice_idx[p];
      float ax[6], R[9], xw[3];
#pragma unroll
      for (int d = 0; d < 6; ++d) ax[d] = a.axisangle[(size_t)k * 6 + d];
      rodrigues<float>(ax, R);
      float eps[3];
      if (a.noise) {
#pragma unroll
        for (int d = 0; d < 3; ++d) eps[d] = a.noise[sidx * 3 + d];
      } else {
        normal3(a.seed, a.offset + (uint64_t)sidx, eps);
      }
      float y[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) y[d] = (a.xyz[p * 3 + d] + eps[d] * a.psf_sigma[(size_t)k * 3 + d]) + ax[3 + d];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        xw[i] = R[i * 3] * y[0] + R[i * 3 + 1] * y[1] + R[i * 3 + 2] * y[2];
        g.xn[i] = (xw[i] - cfg.bbox_lo[i]) / (cfg.bbox_hi[i] - cfg.bbox_lo[i]);
      }
      if (POSE) {
#pragma unroll
        for (int d = 0; d < 3; ++d) g.y[d] = y[d];
        g.k = k;
      }
      if (xb == 0) {
        float* pxw = sf + L::fxw + (size_t)b * 768 + 3 * crow;
        pxw[0] = xw[0];
        pxw[1] = xw[1];
        pxw[2] = xw[2];
      }
      unsigned char* txb = gt + L::tx + (size_t)b * 128 * 32 * 2;
      g.slow = encode_warp(g.xn, lt, cfg.grid.n_levels, a.table,
                           [&](int l, __half2 v) { *reinterpret_cast<__half2*>(txb + umma::tile_off(srow, 2 * l, 32)) = v; },
                           [&](int c, uint4 v) { *reinterpret_cast<uint4*>(txb + umma::tile_off(srow, 8 * c, 32)) = v; });
      umma::fence_smem_to_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_full + b);
      tick(0);
    };
    auto scatter = [&](int64_t tl, const Geo& g) {
      umma::mbar_wait(dx_full, (uint32_t)(tl & 1));
      tick(3);
      auto fetch = [&](int l) {
        return *reinterpret_cast<const float2*>(&sdx[srow * 32 + ((2 * l) ^ ((srow & 15) << 1))]);
      };
      float gx[3];
      if (POSE) {
        scatter_warp<true>(g.xn, lt, cfg.grid.n_levels, a.table, fetch, inv_gscale, a.g_table, gx, g.slow);
        __syncwarp();
        if (lane == 0) mbar_arrive(dx_empty);
        float ax[6], R[9], gwd[3], part[12];
#pragma unroll
        for (int d = 0; d < 6; ++d) ax[d] = a.axisangle[(size_t)g.k * 6 + d];
        rodrigues<float>(ax, R);
#pragma unroll
        for (int i = 0; i < 3; ++i) gwd[i] = xb ? 0.f : gx[i] / (cfg.bbox_hi[i] - cfg.bbox_lo[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int q = 0; q < 3; ++q) part[i * 3 + q] = gwd[i] * g.y[q];  // dL/dR
#pragma unroll
        for (int q = 0; q < 3; ++q) part[9 + q] = R[q] * gwd[0] + R[3 + q] * gwd[1] + R[6 + q] * gwd[2];  // dL/dT = R^T g
#pragma unroll
        for (int q = 0; q < 12; ++q) part[q] = warp_sum(part[q]);
        // the Rodrigues VJP is linear in dL/dR, so every warp (16 samples of one slice) pushes its own partial
        if (lane == 0) {
          float gwv[3];
          rodrigues_vjp<float>(ax, part, gwv);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            red_add(a.g_axisangle + (size_t)g.k * 6 + q, gwv[q]);
            red_add(a.g_axisangle + (size_t)g.k * 6 + 3 + q, part[9 + q]);
          }
        }
      } else {
        scatter_warp<false>(g.xn, lt, cfg.grid.n_levels, a.table, fetch, inv_gscale, a.g_table, gx, g.slow);
        __syncwarp();
        if (lane == 0) mbar_arrive(dx_empty);
      }
    };
    Geo ga, gb;
    if (TIMED) t_last = clock64();
    for (int64_t t = -1; t < n_my; ++t) {  // one call site each: the code of both loops stays resident in the instruction cache
      if (t + 1 < n_my) gather(t + 1, gb);
      if (t >= 0) scatter(t, ga);
      ga = gb;
    }
    if (TIMED) {
      tick(2);
      if (lane == 0 && a.timers)
        for (int i = 0; i < 4; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a.timers) + i, (unsigned long long)t_acc[i]);
    }
  } else {
    // =====================================================================================================
    // chain warps: MLP forward / backward on tcgen05, render + losses; thread = TMEM lane = sample row
    // =====================================================================================================
    reg_dec<64>();
    const int cwarp = warp - kMemWarps, grp = cwarp >> 2, cw = cwarp & 3;
    const int row = 32 * cw + lane, crow = grp * kGR + row;
    unsigned char* gt = smem + L::b_streams + (size_t)grp * L::s_bytes;
    uint64_t* mbar = mbars + L::m_mma + grp;
    uint64_t* x_full = mbars + L::m_xfull + 2 * grp;
    uint64_t* x_empty = mbars + L::m_xempty + 2 * grp;
    uint64_t* dx_full = mbars + L::m_dxfull + grp;
    uint64_t* dx_empty = mbars + L::m_dxempty + grp;
    const uint32_t td = tm + L::c_d + 64u * grp;              // this stream's forward / dgrad region
    const uint32_t tacc = tm + ((16u * grp) << 16);           // this stream's wgrad accumulators (lane offset)
    const uint32_t tlane = (uint32_t)(32 * cw) << 16;         // this warp's lane quarter
    const bool issuer = (cw == 0 && lane == 0);
    uint32_t ph = 0, acc_on = 0;

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

    const uint32_t s_w0 = umma::saddr(wt + L::w0), s_wh = umma::saddr(wt + L::wh), s_wo = umma::saddr(wt + L::wo);
    const uint32_t s_ws0 = umma::saddr(wt + L::ws0), s_wso = umma::saddr(wt + L::wso);
    const uint32_t s_tx0 = umma::saddr(gt + L::tx), s_th = umma::saddr(gt + L::th), s_tg = umma::saddr(gt + L::tg);
    const uint32_t s_tsx = umma::saddr(gt + L::tsx), s_tsh = umma::saddr(gt + L::tsh);
    constexpr uint32_t RG64 = 8 * 128, RG32 = 4 * 128, RG16 = 2 * 128;  // byte stride between 8-row groups of a tile
    auto mma_fwd = [&](uint32_t s_a, uint32_t rg_a, uint32_t s_w, uint32_t rg_w, int K, int N) {
      for (int k = 0; k < K / 16; ++k)
        umma::mma_f16(td, umma::smem_desc(s_a + k * 256, 128, rg_a), umma::smem_desc(s_w + k * 256, 128, rg_w),
                      umma::instr_desc(128, N, false, false), k > 0);
    };
    auto mma_dgrad = [&](uint32_t s_dc, uint32_t rg_dc, uint32_t s_w, uint32_t rg_w, int K, int N) {
      for (int k = 0; k < K / 16; ++k)
        umma::mma_f16(td, umma::smem_desc(s_dc + k * 256, 128, rg_dc), umma::smem_desc(s_w + k * 2 * rg_w, rg_w, 128),
                      umma::instr_desc(128, N, false, true), k > 0);
    };
    auto mma_wgrad = [&](uint32_t col, uint32_t s_p, uint32_t s_q, uint32_t rg_q, int N) {
      for (int k = 0; k < kGR / 16; ++k)
        umma::mma_f16(tacc + col, umma::smem_desc(s_p + k * 2 * RG64, RG64, 128), umma::smem_desc(s_q + k * 2 * rg_q, rg_q, 128),
                      umma::instr_desc(64, N, true, true), acc_on | (uint32_t)(k > 0));
    };
    auto publish = [&]() {
      tick(2);
      umma::fence_smem_to_async();
      umma::fence_before_sync();
      chain_barrier(grp);
      tick(3);
    };
    auto wait_mma = [&]() {
      umma::mbar_wait(mbar, ph);
      ph ^= 1u;
      umma::fence_after_sync();
      tick(1);
    };
    if (TIMED) t_last = clock64();

    float loss_d = 0.f, loss_s = 0.f, loss_i = 0.f;
    const int wpp = S >> 5;           // chain warps per pixel (>= 1)
    const bool wide = S > kGR;        // a pixel spans both streams' tiles: pixel-level syncs are CTA-wide
    const float invS = 1.f / (float)S, invB = 1.f / (float)a.B, gscale = cfg.grad_scale;
    const int64_t first = (int64_t)blockIdx.x * kNStreams + grp;
    const int64_t n_my = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;
    float* sdx = reinterpret_cast<float*>(gt + L::dx);

    for (int64_t t = 0; t < n_my; ++t) {
      const int b = (int)(t & 1);
      const int64_t tile = first + t * stride;
      const int64_t sidx = tile * kGR + row;
      const int64_t p = sidx >> a.log2S;
      const int j = (int)(sidx & (S - 1));
      const int k = (int)a.slice_idx[p];
      const float v_p = a.v[p];
      const float ck = cfg.slice_scale ? (float)a.n_slices * expf(a.logit_coef[k] - lse) : 1.f;
      const float evs = cfg.slice_variance ? expf(a.log_var_slice[k]) : 0.f;
      const uint32_t s_tx = s_tx0 + (uint32_t)b * 128 * 32 * 2;

      tick(3);
      umma::mbar_wait(x_full + b, (uint32_t)((t >> 1) & 1));
      umma::fence_after_sync();
      tick(0);
      // ================= density MLP forward =================
      if (issuer) {
        mma_fwd(s_tx, RG32, s_w0, RG32, 32, 64);
        umma::commit(mbar);
      }
      wait_mma();
      ws_relu_store(td + tlane, gt + L::th, row, 0);
      ws_relu_store(td + tlane + 32, gt + L::th, row, 32);
#pragma unroll
      for (int l = 1; l < DEPTH; ++l) {
        publish();
        if (issuer) {
          umma::fence_after_sync();
          mma_fwd(s_th + (l - 1) * 128 * 64 * 2, RG64, s_wh + (l - 1) * 64 * 64 * 2, RG64, 64, 64);
          umma::commit(mbar);
        }
        wait_mma();
        ws_relu_store(td + tlane, gt + L::th + (size_t)l * 128 * 64 * 2, row, 0);
        ws_relu_store(td + tlane + 32, gt + L::th + (size_t)l * 128 * 64 * 2, row, 32);
      }
      publish();
      if (issuer) {
        umma::fence_after_sync();
        mma_fwd(s_th + (DEPTH - 1) * 128 * 64 * 2, RG64, s_wo, RG64, 64, 16);
        umma::commit(mbar);
      }
      wait_mma();
      float z0, lv = 0.f;
      {
        uint32_t z[16];
        umma::tmem_ld16(td + tlane, z);
        umma::tmem_ld_wait();
        z0 = __uint_as_float(z[0]);
        if (SIGMA) {  // sigma_net input = [slice embedding (16) | z (16; z0 meets a structurally zero weight column)]
          float zf[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) zf[c] = __uint_as_float(z[c]);
          *reinterpret_cast<uint4*>(gt + L::tsx + umma::tile_off(row, 16, 32)) = pack8(zf);
          *reinterpret_cast<uint4*>(gt + L::tsx + umma::tile_off(row, 24, 32)) = pack8(zf + 8);
          const float4* se = reinterpret_cast<const float4*>(a.slice_embedding + (size_t)k * 16);
          float sv[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 q = __ldg(se + c);
            sv[4 * c] = q.x;
            sv[4 * c + 1] = q.y;
            sv[4 * c + 2] = q.z;
            sv[4 * c + 3] = q.w;
          }
          *reinterpret_cast<uint4*>(gt + L::tsx + umma::tile_off(row, 0, 32)) = pack8(sv);
          *reinterpret_cast<uint4*>(gt + L::tsx + umma::tile_off(row, 8, 32)) = pack8(sv + 8);
        }
      }
      // ================= sigma MLP forward =================
      if (SIGMA) {
        publish();
        if (issuer) {
          umma::fence_after_sync();
          mma_fwd(s_tsx, RG32, s_ws0, RG32, 32, 64);
          umma::commit(mbar);
        }
        wait_mma();
        ws_relu_store(td + tlane, gt + L::tsh, row, 0);
        ws_relu_store(td + tlane + 32, gt + L::tsh, row, 32);
        publish();
        if (issuer) {
          umma::fence_after_sync();
          mma_fwd(s_tsh, RG64, s_wso, RG64, 64, 16);
          umma::commit(mbar);
        }
        wait_mma();
        uint32_t z[16];
        umma::tmem_ld16(td + tlane, z);
        umma::tmem_ld_wait();
        lv = __uint_as_float(z[0]);
      }
      // ================= render, losses, gradients w.r.t. z0 / log_var (thread = sample) =================
      tick(2);
      const float rho = softplus_f(z0);
      const float u = SIGMA ? expf(lv) : 1.f;
      sf[L::frho + crow] = rho;
      {
        const float s_rho = warp_sum(rho), s_u = warp_sum(u);
        if (lane == 0) {
          sf[L::fred + cwarp * 2] = s_rho;
          sf[L::fred + cwarp * 2 + 1] = s_u;
        }
      }
      umma::fence_before_sync();
      if (wide) chain_barrier_all(); else chain_barrier(grp);
      float m_pix = 0.f, q_pix = 0.f;
      {
        const int w0 = (cwarp / wpp) * wpp;
        for (int q = 0; q < wpp; ++q) {
          m_pix += sf[L::fred + (w0 + q) * 2];
          q_pix += sf[L::fred + (w0 + q) * 2 + 1];
        }
        m_pix *= invS;
        q_pix *= invS;
      }
      const float vhat = ck * m_pix;
      const float r = ck * q_pix;
      float var = SIGMA ? r * r : 1.f;
      var += evs;
      const float e = vhat - v_p;
      const float d_vhat = e / var * invB;
      const float d_var = (SIGMA || cfg.slice_variance) ? (0.5f / var - 0.5f * e * e / (var * var)) * invB : 0.f;
      float d_rho = ck * d_vhat * invS;
      const float d_lv = SIGMA ? (u * invS) * ck * 2.f * r * d_var : 0.f;
      if (j == 0) {
        loss_d += 0.5f * e * e / var * invB;
        if (SIGMA || cfg.slice_variance) loss_s += 0.5f * logf(var) * invB;
        if (a.v_out) a.v_out[p] = vhat;
        if (cfg.slice_scale) red_add(a.g_c + k, m_pix * d_vhat);
        if (cfg.slice_variance) red_add(a.g_lvs + k, evs * d_var);
      }
      if (cfg.image_reg) {
        const int tp = (crow & ~(S - 1)) + (S - 1 - j);
        const float* xw = sf + L::fxw + (size_t)b * 768 + 3 * crow;
        const float* xp = sf + L::fxw + (size_t)b * 768 + 3 * tp;
        const float dr = rho - sf[L::frho + tp];
        const float dx0 = xw[0] - xp[0], dx1 = xw[1] - xp[1], dx2 = xw[2] - xp[2];
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
        loss_i += li;
      }
      const float dz0 = (z0 > 20.f ? 1.f : sigmoid_f(z0)) * d_rho * gscale;
      {
        // dL/dz tile: the sigma pass first carries dL/d(log_var) in column 0; the density pass carries dz0 (+ sigma's dz)
        const float g0 = SIGMA ? d_lv * gscale : dz0;
        *reinterpret_cast<uint4*>(gt + L::tg + umma::tile_off(row, 0, 16)) =
            make_uint4((uint32_t)__half_as_ushort(__float2half_rn(g0)), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(gt + L::tg + umma::tile_off(row, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
      }
      if (wide) chain_barrier_all();  // the partner stream has finished reading this stream's rho / xw rows
      tick(4);
      publish();

      // ================= backward =================
      if (SIGMA) {
        if (issuer) {
          umma::fence_after_sync();
          mma_wgrad(L::c_wso, s_tsh, s_tg, RG16, 16);       // dWso^T += Hs^T G
          mma_dgrad(s_tg, RG16, s_wso, RG64, 16, 64);        // dHs = G Wso
          umma::commit(mbar);
        }
        wait_mma();
        ws_mask_store(td + tlane, gt + L::tsh, row, 0);
        ws_mask_store(td + tlane + 32, gt + L::tsh, row, 32);
        publish();
        if (issuer) {
          umma::fence_after_sync();
          mma_wgrad(L::c_ws0, s_tsh, s_tsx, RG32, 32);       // dWs0 += dZs^T [se | z]
          mma_dgrad(s_tsh, RG64, s_ws0, RG32, 64, 32);       // d[se | z] = dZs Ws0
          umma::commit(mbar);
        }
        wait_mma();
        {
          uint32_t d[16];
          umma::tmem_ld16(td + tlane, d);  // d(slice embedding): column sums over the warp's 32 rows (one pixel, one slice)
          umma::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float s = warp_sum(__uint_as_float(d[c]));
            if (lane == c) red_add(a.g_se + (size_t)k * 16 + c, s * inv_gscale);
          }
          umma::tmem_ld16(td + tlane + 16, d);  // dL/dz from sigma_net (+ dz0 of the render path in column 0)
          umma::tmem_ld_wait();
          float g[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) g[c] = __uint_as_float(d[c]);
          g[0] += dz0;
          *reinterpret_cast<uint4*>(gt + L::tg + umma::tile_off(row, 0, 16)) = pack8(g);
          *reinterpret_cast<uint4*>(gt + L::tg + umma::tile_off(row, 8, 16)) = pack8(g + 8);
        }
        publish();
      }
      {
        const uint32_t s_hl = s_th + (DEPTH - 1) * 128 * 64 * 2;
        if (issuer) {
          umma::fence_after_sync();
          mma_wgrad(L::c_wo, s_hl, s_tg, RG16, 16);          // dWo^T += H_last^T G
          mma_dgrad(s_tg, RG16, s_wo, RG64, 16, 64);          // dH_last = G Wo
          umma::commit(mbar);
        }
        wait_mma();
        ws_mask_store(td + tlane, gt + L::th + (size_t)(DEPTH - 1) * 128 * 64 * 2, row, 0);
        ws_mask_store(td + tlane + 32, gt + L::th + (size_t)(DEPTH - 1) * 128 * 64 * 2, row, 32);
        publish();
      }
#pragma unroll
      for (int l = DEPTH - 1; l >= 1; --l) {
        const uint32_t s_dz = s_th + l * 128 * 64 * 2, s_hp = s_th + (l - 1) * 128 * 64 * 2;
        if (issuer) {
          umma::fence_after_sync();
          mma_wgrad(L::c_wh + 64 * (l - 1), s_dz, s_hp, RG64, 64);          // dWh_{l-1} += dZ_l^T H_{l-1}
          mma_dgrad(s_dz, RG64, s_wh + (l - 1) * 64 * 64 * 2, RG64, 64, 64);  // dH_{l-1} = dZ_l Wh_{l-1}
          umma::commit(mbar);
        }
        wait_mma();
        ws_mask_store(td + tlane, gt + L::th + (size_t)(l - 1) * 128 * 64 * 2, row, 0);
        ws_mask_store(td + tlane + 32, gt + L::th + (size_t)(l - 1) * 128 * 64 * 2, row, 32);
        publish();
      }
      if (issuer) {
        umma::fence_after_sync();
        mma_wgrad(L::c_w0, s_th, s_tx, RG32, 32);            // dW0 += dZ_0^T X
        mma_dgrad(s_th, RG64, s_w0, RG32, 64, 32);           // dX = dZ_0 W0
        umma::commit(mbar);
        umma::commit(x_empty + b);                           // the tile's last reader of X[b] is done -> memory warps may refill it
      }
      wait_mma();
      acc_on = 1u;
      // dL/d(features): fp32 [128][32], feature pair (2l, 2l+1) of row r at r*32 + ((2l) ^ ((r & 15) << 1)) -- 8-byte
      // accesses, conflict-free for both the row-per-lane writes here and the sample-pair reads of the scatter
      tick(2);
      if (t >= 1) umma::mbar_wait(dx_empty, (uint32_t)((t - 1) & 1));  // scatter(t - 1) has consumed the previous dX
      tick(5);
      {
        uint32_t d[32];
        umma::tmem_ld32(td + tlane, d);
        umma::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; c += 2)
          *reinterpret_cast<float2*>(&sdx[row * 32 + (c ^ ((row & 15) << 1))]) = make_float2(__uint_as_float(d[c]), __uint_as_float(d[c + 1]));
      }
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(dx_full);
      chain_barrier(grp);  // every warp has drained its TMEM reads before the next tile's first MMA overwrites the region
    }
    if (TIMED) {
      tick(3);
      if (lane == 0 && a.timers)
        for (int i = 0; i < 6; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a.timers) + 8 + i, (unsigned long long)t_acc[i]);
    }
    if (n_my > 0 && cw == 0 && lane == 0) grp_ran[grp] = 1u;
    loss_d = warp_sum(loss_d);
    loss_s = warp_sum(loss_s);
    loss_i = warp_sum(loss_i);
    if (lane == 0) {
      red_add(a.losses + 0, loss_d);
      red_add(a.losses + 1, loss_s);
      red_add(a.losses + 3, loss_i);
    }
  }

  // ================= epilogue: weight gradients TMEM -> global =================
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (warp >= kMemWarps) {
    // lanes [0,16) of a quarter hold stream 0's accumulator rows 16 q + lane, lanes [16,32) stream 1's
    const int cwarp = warp - kMemWarps;
    const int q = cwarp & 3, part = cwarp >> 2;  // 2 column partitions across the 8 chain warps
    const int arow = 16 * q + (lane & 15), agrp = lane >> 4;
    const bool live = grp_ran[agrp] != 0;
    const uint32_t tq = tm + ((uint32_t)(32 * q) << 16);
    float* gd = a.g_mlp + a.off_density;
    float* gs = a.g_mlp + a.off_sigma;
    int chunk = 0;
    auto flush = [&](uint32_t col, int ncols, float* dst, int ld, bool transposed) {
      for (int c0 = 0; c0 < ncols; c0 += 16, ++chunk) {
        if ((chunk & 1) != part) continue;
        uint32_t d[16];
        umma::tmem_ld16(tq + col + c0, d);
        umma::tmem_ld_wait();
        if (!live) continue;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float v = __uint_as_float(d[c]) * inv_gscale;
          if (v != 0.f) red_add(transposed ? dst + (size_t)(c0 + c) * ld + arow : dst + (size_t)arow * ld + c0 + c, v);
        }
      }
    };
    flush(L::c_w0, 32, gd, 32, false);
    for (int l = 0; l + 1 < DEPTH; ++l) flush(L::c_wh + 64 * l, 64, gd + 64 * 32 + (size_t)l * 64 * 64, 64, false);
    flush(L::c_wo, 16, gd + 64 * 32 + (size_t)(DEPTH - 1) * 64 * 64, 64, true);
    if (SIGMA) {
      flush(L::c_ws0, 32, gs, 32, false);
      flush(L::c_wso, 16, gs + 64 * 32, 64, true);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 512);
}

template <int DEPTH, bool SIGMA, bool POSE>
int launch_ws(const FusedArgs& a, cudaStream_t st) {
  using L = WsLayout<DEPTH, SIGMA>;
  static_assert(L::bytes <= 227 * 1024, "shared memory");
  cudaError_t e = cudaFuncSetAttribute(inr_train_ws_kernel<DEPTH, SIGMA, POSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (e != cudaSuccess) {
    set_error("nsv_inr_train_step(ws): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  const int64_t ctas = (a.B * (int64_t)a.S / kGR + kNStreams - 1) / kNStreams;
  const int grid = (int)(ctas < num_sms() ? ctas : num_sms());
  if (a.timers && DEPTH == 3 && !SIGMA && !POSE) {  // profiling build of the config-2 instantiation
    cudaFuncSetAttribute(inr_train_ws_kernel<3, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WsLayout<3, false>::bytes);
    inr_train_ws_kernel<3, false, false, true><<<grid, kWsThreads, L::bytes, st>>>(a);
  } else {
    inr_train_ws_kernel<DEPTH, SIGMA, POSE><<<grid, kWsThreads, L::bytes, st>>>(a);
  }
  if (int err = check_launch("nsv_inr_train_step(ws)")) return err;
  inr_finalize_kernel<<<1, 256, 0, st>>>(a.logit_coef, a.g_c, a.losses, a.n_slices, a.cfg.slice_scale, a.cfg.image_reg, a.cfg.delta);
  return check_launch("nsv_inr_train_step(finalize)");
}

}  // namespace

int launch_train_ws(const FusedArgs& a, cudaStream_t st) {
  const nsv_inr_config& c = a.cfg;
  if (c.width != kW || c.depth < 1 || c.depth > 3 || (c.pixel_variance && c.depth != 1) || (a.B * (int64_t)a.S) % (kGR * kNStreams) != 0) {
    set_error("nsv_inr_train_step: no warp-specialised instantiation for width=%d depth=%d (needs width 64, depth 1..3, B*S %% 256 == 0)", c.width, c.depth);
    return NSV_EUNSUPPORTED;
  }
  if (c.pose_grad) {
    if (c.pixel_variance) return launch_ws<1, true, true>(a, st);
    if (c.depth == 1) return launch_ws<1, false, true>(a, st);
    if (c.depth == 2) return launch_ws<2, false, true>(a, st);
    return launch_ws<3, false, true>(a, st);
  }
  if (c.pixel_variance) return launch_ws<1, true, false>(a, st);
  if (c.depth == 1) return launch_ws<1, false, false>(a, st);
  if (c.depth == 2) return launch_ws<2, false, false>(a, st);
  return launch_ws<3, false, false>(a, st);
}

}  // namespace fused
}  // namespace nsv
