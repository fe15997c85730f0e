// inr_fused_tc.cu -- kernel A on Blackwell's 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract and phases as inr_fused.cu (one fused NeSVoR training iteration, reference op
// sequence nesvor/nesvor/models.py:260-384 + train.py:183-190), but every matrix product of the
// density / sigma MLPs -- forward, dgrad and wgrad -- is a tcgen05.mma issued by ONE thread per
// 128-sample group, with operands read straight from shared memory and accumulators in TMEM:
//   * activations / gradients live in shared memory as un-swizzled canonical tiles (umma.cuh); the
//     same tile is the K-major A operand of forward / dgrad and the MN-major operand of wgrad;
//   * forward / dgrad results (128 x N fp32) land in the group's 64-column TMEM region and are
//     pulled to registers with tcgen05.ld by the epilogue threads (thread = TMEM lane = sample row),
//     which apply ReLU / the ReLU mask, round to fp16 and write the next operand tile;
//   * weight gradients (M = 64 accumulators) stay resident in TMEM for the whole kernel -- the two
//     groups of a CTA interleave theirs on TMEM lanes [0,16) / [16,32) of every lane quarter -- and are
//     read out and reduced to global memory once per CTA: wgrad costs zero SM instructions per tile;
//   * no ldmatrix / mma.sync / fragment shuffling remains, which removes ~2/3 of the MLP-phase
//     instructions and all shared-memory operand traffic through the LSU.
// Gather / scatter phases, loss math and lane mapping are shared with inr_fused.cu (inr_common.cuh).
// Instantiated for width 64 (UMMA M = 64 wgrad), depth 1..3, n_samples in {32,...,256}.
#include "inr_common.cuh"
#include "umma.cuh"

namespace nsv {
namespace fused {
namespace {

constexpr int kW = 64, kGR = 128, kGT = 256;
// NG = number of 128-sample groups per CTA (256 threads each).  NG = 2: the round-1 kernel (512 threads, 128 registers).
// NG = 3 (round 2; density-only configurations with n_samples <= 128): 768 threads at 80 registers.  With two groups the MMA
// chain of one group (17 K cycles per tile, almost all of it latency) is covered by only 8 gather / scatter warps, which
// reach 75 % of the LSU sector rate; measured, the kernel ran in exactly (LSU time at full rate) + (chain time).  A third
// group keeps 16 memory warps busy while one group waits on the tensor core.  TMEM has no room for a third set of
// weight-gradient accumulators next to three forward / dgrad regions (3 x 64 + 3 x 176 > 512 columns), and does not need one:
// accumulating tcgen05.mma instructions of different issuing threads compose exactly (probed by nsv_umma_shared_accumulator_test,
// tests/test_gpu_umma.py), so the three groups add into ONE zero-initialised set.

template <int DEPTH, bool SIGMA, bool BIAS = false, int NG = 2>
struct TcLayout {
  static_assert(NG == 2 || (NG == 3 && !SIGMA && !BIAS), "the 3-group kernel is instantiated for the density-only configurations");
  static_assert(!BIAS || (SIGMA && DEPTH == 1), "the fused bias-field head rides on the sigma_net instantiation (depth 1, slice embedding on)");
  // ---- CTA-shared canonical weight tiles (byte offsets) ----
  static constexpr size_t w0 = 0;                                         // [64][32]
  static constexpr size_t wh = w0 + 64 * 32 * 2;                          // (DEPTH-1) x [64][64]
  static constexpr size_t wo = wh + (size_t)(DEPTH - 1) * 64 * 64 * 2;    // [16][64]
  static constexpr size_t ws0 = wo + 16 * 64 * 2;                         // [64][32]
  static constexpr size_t wso = ws0 + (SIGMA ? 64 * 32 * 2 : 0);          // [16][64]
  static constexpr size_t wb0 = wso + (SIGMA ? 16 * 64 * 2 : 0);           // [64][32] b_net first layer: [slice embedding(16) | pe_bias(8) | 0]
  static constexpr size_t wbo = wb0 + (BIAS ? 64 * 32 * 2 : 0);           // [16][64]
  static constexpr size_t w_end = wbo + (BIAS ? 16 * 64 * 2 : 0);
  // ---- per-group canonical activation tiles (byte offsets from the group base) ----
  static constexpr size_t tx = 0;                                         // [128][32] encoded features
  static constexpr size_t th = tx + 128 * 32 * 2;                         // DEPTH x [128][64] hidden, later dZ
  static constexpr size_t tg = th + (size_t)DEPTH * 128 * 64 * 2;         // [128][16] dL/dz
  static constexpr size_t tsx = tg + 128 * 16 * 2;                        // [128][32] sigma_net input
  static constexpr size_t tsh = tsx + (SIGMA ? 128 * 32 * 2 : 0);         // [128][64] sigma_net hidden
  static constexpr size_t tbx = tsh + (SIGMA ? 128 * 64 * 2 : 0);         // [128][32] b_net input
  static constexpr size_t tbh = tbx + (BIAS ? 128 * 32 * 2 : 0);          // [128][64] b_net hidden, later dZb; dead after the bias pass
  static constexpr size_t tgb = tbh + (BIAS ? 128 * 64 * 2 : 0);          // [128][16] dL/d(log_bias)
  static constexpr size_t dxb = tgb + (BIAS ? 128 * 16 * 2 : 0);          // [128][8] fp32: dL/d(pe_bias) coming out of b_net
  static constexpr bool alias_dx = DEPTH >= 2;                            // dL/d(features) reuses the dead H_last slot
  static constexpr size_t dx = dxb + (BIAS ? 128 * 8 * 4 : 0);            // [128][32] fp32, XOR-swizzled (BIAS: the dead tbh slot)
  static constexpr size_t g_bytes = dx + ((alias_dx || BIAS) ? 0 : 128 * 32 * 4);
  // ---- CTA-level fp32 scratch, indexed by CTA row (group * 128 + row) ----
  static constexpr size_t b_groups = (w_end + 127) / 128 * 128;
  static constexpr size_t b_scr = b_groups + NG * g_bytes;
  static constexpr size_t nrow = (size_t)NG * 128;  // sample rows per CTA
  static constexpr size_t fz0 = 0, flv = nrow, frho = 2 * nrow, fxw = 3 * nrow, fred = 6 * nrow, flb = fred + 16 * 16;  // floats
  static constexpr size_t fend = flb + (BIAS ? nrow : 0);
  static constexpr size_t b_lt = b_scr + fend * 4;
  static constexpr size_t b_sync = (b_lt + sizeof(LevelTable) + 15) / 16 * 16;  // mbar[4] @0, tmem slot @32, grp_ran[4] @40, table mbarrier @56
  static constexpr size_t bytes = (b_sync + 64 + 127) / 128 * 128;
  // [bytes, bytes + staged table bytes): shared-memory copy of the coarsest levels of the fp16 hash table (FusedArgs::smem_levels)
  // ---- TMEM columns ----
  static constexpr uint32_t c_d = 0;                                      // group g: [64 g, 64 g + 64)
  static constexpr uint32_t c_w0 = 64 * NG;                               // dW0   [64 x 32]
  static constexpr uint32_t c_wh = c_w0 + 32;                             // dWh_l [64 x 64]
  static constexpr uint32_t c_wo = c_wh + 64 * (DEPTH - 1);               // dWo^T [64 x 16]
  static constexpr uint32_t c_ws0 = c_wo + 16;                            // dWs0  [64 x 32]
  static constexpr uint32_t c_wso = c_ws0 + 32;                           // dWso^T [64 x 16]
  static constexpr uint32_t c_wb0 = c_wso + 16;                           // dWb0  [64 x 32]
  static constexpr uint32_t c_wbo = c_wb0 + 32;                           // dWbo^T [64 x 16]
  static constexpr uint32_t c_end = c_wbo + 16;
  static constexpr uint32_t c_d2 = 384;                                   // BIAS: b_net's forward / dgrad region, group g: [384 + 64 g, +64)
  static_assert(c_end <= (BIAS ? c_d2 : 512), "TMEM columns");
};

__device__ __forceinline__ void cta_barrier_all() { asm volatile("bar.sync 4, 512;" ::: "memory"); }  // 2-group kernel only (S = 256)

// ---- epilogues: this thread owns TMEM lane (= sample row) `row` and 32 accumulator columns starting at c0 ----
__device__ __forceinline__ void epi_relu_store(uint32_t taddr, unsigned char* tile, int row, int c0) {
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
__device__ __forceinline__ void epi_mask_store(uint32_t taddr, unsigned char* tile, int row, int c0) {
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

// TIMED (profiling builds only, a.timers != NULL): every warp accumulates the clock cycles it spends per phase
// (0 geometry + gather, 1 publish / group barriers, 2 MMA issue -> mbarrier wait, 3 epilogues, 4 render + losses,
//  5 scatter, 6 pixel barrier) and adds them to a.timers[phase] at the end
template <int DEPTH, bool SIGMA, bool TIMED = false, bool BIAS = false, int NG = 2>
__global__ void __launch_bounds__(256 * NG, 1) inr_train_tc_kernel(const __grid_constant__ FusedArgs a) {
  using L = TcLayout<DEPTH, SIGMA, BIAS, NG>;
  constexpr int kNGroups = NG, kThreads = 256 * NG;
  long long t_acc[TIMED ? 8 : 1] = {};
  long long t_last = TIMED ? clock64() : 0;
  auto tick = [&](int seg) {
    if (TIMED) {
      const long long t = clock64();
      t_acc[seg] += t - t_last;
      t_last = t;
    }
  };
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 3, gw = warp & 7, row0 = gw * 16;
  const int xb = lane & 1, srow = row0 + (lane >> 1);  // gather / loss mapping: lane pair = one sample row
  const int crow = grp * kGR + srow;                   // CTA-level row of that sample
  const int q4 = gw & 3, half = gw >> 2;               // epilogue mapping: TMEM lane quarter, column half
  const int erow = 32 * q4 + lane;                     // sample row owned in epilogues
  unsigned char* wt = smem;
  unsigned char* gt = smem + L::b_groups + (size_t)grp * L::g_bytes;
  float* sf = reinterpret_cast<float*>(smem + L::b_scr);
  LevelTable& lt = *reinterpret_cast<LevelTable*>(smem + L::b_lt);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + L::b_sync) + grp;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::b_sync + 32);
  uint32_t* grp_ran = reinterpret_cast<uint32_t*>(smem + L::b_sync + 40);
  uint64_t* tbar = reinterpret_cast<uint64_t*>(smem + L::b_sync + 56);  // completion of the staged-table bulk copy
  const nsv_inr_config& cfg = a.cfg;
  // TMA-staged table prefix: levels [0, smem_levels) are contiguous at the start of the flat fp16 table (level-major tcnn
  // layout, every level a multiple of 8 entries = 32 bytes), so ONE bulk copy per CTA brings them into shared memory
  const uint32_t stab_bytes = a.smem_table_bytes;

  // ---- one-time setup: weights -> canonical tiles, level table, TMEM, mbarriers ----
  {
    const __half* wd = a.mlp + a.off_density;
    umma::stage_tile(wt + L::w0, wd, 64, 32, tid, kThreads);
    for (int l = 0; l + 1 < DEPTH; ++l) umma::stage_tile(wt + L::wh + (size_t)l * 64 * 64 * 2, wd + 64 * 32 + (size_t)l * 64 * 64, 64, 64, tid, kThreads);
    umma::stage_tile(wt + L::wo, wd + 64 * 32 + (size_t)(DEPTH - 1) * 64 * 64, 16, 64, tid, kThreads);
    if (SIGMA) {
      const __half* ws = a.mlp + a.off_sigma;
      umma::stage_tile(wt + L::ws0, ws, 64, 32, tid, kThreads);
      umma::stage_tile(wt + L::wso, ws + 64 * 32, 16, 64, tid, kThreads);
    }
    if (BIAS) {
      const __half* wb = a.mlp + a.off_bias;
      umma::stage_tile(wt + L::wb0, wb, 64, 32, tid, kThreads);
      umma::stage_tile(wt + L::wbo, wb + 64 * 32, 16, 64, tid, kThreads);
    }
    stage_level_table(lt, cfg.grid, tid, a.agg_max, a.fast, a.table, a.g_table, a.ablate);
    if (tid < a.smem_levels)  // same thread that wrote tbl[tid] above: the level is gathered from the shared-memory copy
      lt.tbl[tid] = reinterpret_cast<const __half2*>(smem + L::bytes) + lt.offset[tid];
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
      for (int g = 0; g < NG; ++g) umma::mbar_init(reinterpret_cast<uint64_t*>(smem + L::b_sync) + g, 1);
      umma::mbar_init(tbar, 1);
      umma::mbar_fence_init();
      for (int g = 0; g < 4; ++g) grp_ran[g] = 0;
      if (stab_bytes) {  // in flight while the CTA stages its weights; waited for before the first gather
        umma::mbar_expect_tx(tbar, stab_bytes);
        for (uint32_t o = 0; o < stab_bytes; o += 32768u)
          umma::bulk_g2s(smem + L::bytes + o, reinterpret_cast<const unsigned char*>(a.table) + o, min(32768u, stab_bytes - o), tbar);
      }
    }
  }
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
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (stab_bytes) umma::mbar_wait(tbar, 0);
  const uint32_t tm = *tmem_slot;
  const uint32_t td = tm + L::c_d + 64u * grp;                       // this group's forward / dgrad region
  const uint32_t td2 = tm + L::c_d2 + 64u * grp;                     // BIAS: second region, so that b_net's products ride in the same MMA rounds
  // wgrad accumulators: NG = 2 each group its own (TMEM lane offset 16 g); NG = 3 ONE set shared by all groups (lane offset 0)
  const uint32_t tacc = NG == 3 ? tm : tm + ((16u * grp) << 16);
  const uint32_t tlane = (uint32_t)(32 * q4) << 16;                  // epilogue lane quarter
  const bool issuer = (gw == 0 && lane == 0);
  uint32_t ph = 0;                                                   // mbarrier phase parity
  uint32_t acc_on = NG == 3 ? 1u : 0u;                               // NG = 2: 0 on the group's first tile (wgrad MMAs overwrite)
  if (NG == 3) {  // the shared accumulators start from zero: every warp clears the columns of its own lane quarter
    const uint32_t tq0 = tm + ((uint32_t)(32 * (warp & 3)) << 16);
    for (uint32_t c = L::c_w0 + 16 * (uint32_t)(warp >> 2); c < L::c_end; c += 16 * (kThreads / 128)) umma::tmem_st16_fill(tq0 + c, 0u);
    umma::tmem_st_wait();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }

  const uint32_t s_w0 = umma::saddr(wt + L::w0), s_wh = umma::saddr(wt + L::wh), s_wo = umma::saddr(wt + L::wo);
  const uint32_t s_ws0 = umma::saddr(wt + L::ws0), s_wso = umma::saddr(wt + L::wso);
  const uint32_t s_tx = umma::saddr(gt + L::tx), s_th = umma::saddr(gt + L::th), s_tg = umma::saddr(gt + L::tg);
  const uint32_t s_tsx = umma::saddr(gt + L::tsx), s_tsh = umma::saddr(gt + L::tsh);
  const uint32_t s_wb0 = umma::saddr(wt + L::wb0), s_wbo = umma::saddr(wt + L::wbo);
  const uint32_t s_tbx = umma::saddr(gt + L::tbx), s_tbh = umma::saddr(gt + L::tbh), s_tgb = umma::saddr(gt + L::tgb);
  // batch mean of log_bias (biasReg = mean^2, models.py:323), produced by nsv_inr_bias_mean before this launch
  const float bias_mean = BIAS ? __ldg(a.losses + 4) : 0.f;
  constexpr uint32_t RG64 = 8 * 128, RG32 = 4 * 128, RG16 = 2 * 128;  // byte stride between 8-row groups of a tile
  // forward: D[128 x N] = A[128 x K] W[N x K]^T  (A, W K-major)
  auto mma_fwd_d = [&](uint32_t d, uint32_t s_a, uint32_t rg_a, uint32_t s_w, uint32_t rg_w, int K, int N) {
    for (int k = 0; k < K / 16; ++k)
      umma::mma_f16(d, umma::smem_desc(s_a + k * 256, 128, rg_a), umma::smem_desc(s_w + k * 256, 128, rg_w),
                    umma::instr_desc(128, N, false, false), k > 0);
  };
  auto mma_fwd = [&](uint32_t s_a, uint32_t rg_a, uint32_t s_w, uint32_t rg_w, int K, int N) { mma_fwd_d(td, s_a, rg_a, s_w, rg_w, K, N); };
  // dgrad: D[128 x N] = dC[128 x K] W[K x N]  (W tile as MN-major B)
  auto mma_dgrad_d = [&](uint32_t d, uint32_t s_dc, uint32_t rg_dc, uint32_t s_w, uint32_t rg_w, int K, int N) {
    for (int k = 0; k < K / 16; ++k)
      umma::mma_f16(d, umma::smem_desc(s_dc + k * 256, 128, rg_dc), umma::smem_desc(s_w + k * 2 * rg_w, rg_w, 128),
                    umma::instr_desc(128, N, false, true), k > 0);
  };
  auto mma_dgrad = [&](uint32_t s_dc, uint32_t rg_dc, uint32_t s_w, uint32_t rg_w, int K, int N) { mma_dgrad_d(td, s_dc, rg_dc, s_w, rg_w, K, N); };
  // wgrad: acc[64 x N] += P[128 x 64]^T Q[128 x N]  (both tiles MN-major, K = the 128 sample rows)
  auto mma_wgrad = [&](uint32_t col, uint32_t s_p, uint32_t s_q, uint32_t rg_q, int N) {
    for (int k = 0; k < kGR / 16; ++k)
      umma::mma_f16(tacc + col, umma::smem_desc(s_p + k * 2 * RG64, RG64, 128), umma::smem_desc(s_q + k * 2 * rg_q, rg_q, 128),
                    umma::instr_desc(64, N, true, true), acc_on | (uint32_t)(k > 0));
  };
  // writers publish their shared-memory stores to the tensor core, then the group meets
  auto publish = [&]() {
    tick(3);
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    group_barrier(grp, kGT);
    tick(1);
  };
  // The group's 8 warps used to poll the mbarrier each (try_wait wakes every few dozen cycles): 36 M of the kernel's 298 M
  // warp instructions were polling, issued from the same schedulers the other group's gather / scatter code needs
  // (profiles/r02_kernelA_tcgen05_ncu_summary.txt).  Now ONE warp polls and releases the others through the group's named
  // barrier, where waiting costs no issue slots.  a.ablate bit 3 (profiling) restores the old scheme.
  const bool poll_all = (a.ablate & 8u) != 0;
  auto wait_mma = [&]() {
    if (poll_all || gw == 0) umma::mbar_wait(mbar, ph);
    ph ^= 1u;
    if (!poll_all) group_barrier(grp, kGT);
    umma::fence_after_sync();
    tick(2);
  };

  float loss_d = 0.f, loss_s = 0.f, loss_i = 0.f;
  const int S = a.S, wpp = S >> 4;  // warps per pixel
  const bool wide = S > kGR;        // a pixel spans both groups: pixel-level syncs are CTA-wide
  const float invS = 1.f / (float)S, invB = 1.f / (float)a.B, gscale = cfg.grad_scale, inv_gscale = 1.f / cfg.grad_scale;
  const int64_t n_tiles = (a.B * (int64_t)S) / kGR;
  auto pixel_barrier = [&]() {
    if (wide) cta_barrier_all(); else group_barrier(grp, kGT);
  };

  // tile order: strided over the grid, or one contiguous run of tiles per CTA (an even count, so that the two halves of a
  // 256-sample pixel stay in one CTA) -- with a spatially ordered batch the CTA then walks neighbouring pixels
  const int64_t run = ((n_tiles / kNGroups + gridDim.x - 1) / gridDim.x) * kNGroups;
  const int64_t tile_first = a.tile_order ? blockIdx.x * run + grp : (int64_t)blockIdx.x * kNGroups + grp;
  const int64_t tile_stop = a.tile_order ? (n_tiles < (blockIdx.x + 1) * run ? n_tiles : (blockIdx.x + 1) * run) : n_tiles;
  const int64_t tile_step = a.tile_order ? kNGroups : (int64_t)gridDim.x * kNGroups;
  for (int64_t tile = tile_first; tile < tile_stop; tile += tile_step) {
    // ================= phase 0: sample geometry + encoding (lane pair = sample) =================
    tick(1);
    const int64_t sidx = tile * kGR + srow;
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
    const bool slow = encode_warp(xn, lt, cfg.grid.n_levels, a.table,
                [&](int l, __half2 v) { *reinterpret_cast<__half2*>(gt + L::tx + umma::tile_off(srow, 2 * l, 32)) = v; },
                [&](int c, uint4 v) { *reinterpret_cast<uint4*>(gt + L::tx + umma::tile_off(srow, 8 * c, 32)) = v; });
    if (BIAS) {
      // b_net's input [slice embedding (16) | pe_bias (2 n_levels_bias) | 0] (models.py:344-347) depends on nothing the MLPs
      // produce: the gather lanes build it here (each lane of a pair converts 8 of the 16 embedding values, which also
      // serve sigma_net's input tile), so that b_net's products share the density / sigma MMA rounds
      const float4* se = reinterpret_cast<const float4*>(a.slice_embedding + (size_t)k * 16 + 8 * xb);
      const float4 s0 = __ldg(se), s1 = __ldg(se + 1);
      uint4 v;
      {
        const __half2 h0 = __floats2half2_rn(s0.x, s0.y), h1 = __floats2half2_rn(s0.z, s0.w);
        const __half2 h2 = __floats2half2_rn(s1.x, s1.y), h3 = __floats2half2_rn(s1.z, s1.w);
        v = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                       *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
      }
      *reinterpret_cast<uint4*>(gt + L::tsx + umma::tile_off(srow, 8 * xb, 32)) = v;
      *reinterpret_cast<uint4*>(gt + L::tbx + umma::tile_off(srow, 8 * xb, 32)) = v;
      __syncwarp();
      uint4 f = make_uint4(0u, 0u, 0u, 0u);
      if (xb == 0) {  // features of levels 0..3 of this row (stored by this lane above), masked to the first n_levels_bias
        f = *reinterpret_cast<const uint4*>(gt + L::tx + umma::tile_off(srow, 0, 32));
        const int nb = cfg.n_levels_bias;
        if (nb < 2) f.y = 0u;
        if (nb < 3) f.z = 0u;
        if (nb < 4) f.w = 0u;
      }
      *reinterpret_cast<uint4*>(gt + L::tbx + umma::tile_off(srow, 16 + 8 * xb, 32)) = f;
    }
    tick(0);
    publish();
    if (!(a.ablate & 4u)) {
    // ================= phase 1: density MLP forward on tcgen05 (+ b_net, same rounds, second TMEM region) =================
    if (issuer) {
      umma::fence_after_sync();
      mma_fwd(s_tx, RG32, s_w0, RG32, 32, 64);
      if (BIAS) mma_fwd_d(td2, s_tbx, RG32, s_wb0, RG32, 32, 64);
      umma::commit(mbar);
    }
    wait_mma();
    epi_relu_store(td + tlane + 32 * half, gt + L::th, erow, 32 * half);
    if (BIAS) epi_relu_store(td2 + tlane + 32 * half, gt + L::tbh, erow, 32 * half);
#pragma unroll
    for (int l = 1; l < DEPTH; ++l) {
      publish();
      if (issuer) {
        umma::fence_after_sync();
        mma_fwd(s_th + (l - 1) * 128 * 64 * 2, RG64, s_wh + (l - 1) * 64 * 64 * 2, RG64, 64, 64);
        umma::commit(mbar);
      }
      wait_mma();
      epi_relu_store(td + tlane + 32 * half, gt + L::th + (size_t)l * 128 * 64 * 2, erow, 32 * half);
    }
    publish();
    if (issuer) {
      umma::fence_after_sync();
      mma_fwd(s_th + (DEPTH - 1) * 128 * 64 * 2, RG64, s_wo, RG64, 64, 16);
      if (BIAS) mma_fwd_d(td2, s_tbh, RG64, s_wbo, RG64, 64, 16);
      umma::commit(mbar);
    }
    wait_mma();
    if (half == 0) {  // z[0..15] of row erow
      uint32_t z[16];
      umma::tmem_ld16(td + tlane, z);
      umma::tmem_ld_wait();
      sf[L::fz0 + grp * kGR + erow] = __uint_as_float(z[0]);
      if (SIGMA) {  // sigma_net input columns 16..31 = z (column 16, z0, meets a structurally zero weight column)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint4 v;
          uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const __half2 h = __floats2half2_rn(__uint_as_float(z[8 * i + 2 * q]), __uint_as_float(z[8 * i + 2 * q + 1]));
            pv[q] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(gt + L::tsx + umma::tile_off(erow, 16 + 8 * i, 32)) = v;
        }
      }
    } else if (BIAS) {  // log_bias of row erow (the slice-embedding columns were written in phase 0)
      uint32_t zb[16];
      umma::tmem_ld16(td2 + tlane, zb);
      umma::tmem_ld_wait();
      sf[L::flb + grp * kGR + erow] = __uint_as_float(zb[0]);
    } else if (SIGMA) {  // sigma_net input columns 0..15 = slice embedding of the row's slice
      const int64_t pe = (tile * kGR + erow) >> a.log2S;
      const float* se = a.slice_embedding + (size_t)a.slice_idx[pe] * 16;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint4 v;
        uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __half2 h = __floats2half2_rn(se[8 * i + 2 * q], se[8 * i + 2 * q + 1]);
          pv[q] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(gt + L::tsx + umma::tile_off(erow, 8 * i, 32)) = v;
      }
    }
    // ================= phase 1b: sigma MLP forward =================
    if (SIGMA) {
      publish();
      if (issuer) {
        umma::fence_after_sync();
        mma_fwd(s_tsx, RG32, s_ws0, RG32, 32, 64);
        umma::commit(mbar);
      }
      wait_mma();
      epi_relu_store(td + tlane + 32 * half, gt + L::tsh, erow, 32 * half);
      publish();
      if (issuer) {
        umma::fence_after_sync();
        mma_fwd(s_tsh, RG64, s_wso, RG64, 64, 16);
        umma::commit(mbar);
      }
      wait_mma();
      if (half == 0) {
        uint32_t z[16];
        umma::tmem_ld16(td + tlane, z);
        umma::tmem_ld_wait();
        sf[L::flv + grp * kGR + erow] = __uint_as_float(z[0]);
      }
    }
    tick(3);
    umma::fence_before_sync();
    group_barrier(grp, kGT);  // z0 / log_var of every row are in shared memory
    tick(1);

    // ================= phase 2: render, losses, gradients w.r.t. z0 / log_var (lane pair = sample) =================
    const float z0 = sf[L::fz0 + crow];
    const float rho = softplus_f(z0);
    const float lv = SIGMA ? sf[L::flv + crow] : 0.f;
    const float u = SIGMA ? expf(lv) : 1.f;
    const float bias = BIAS ? expf(sf[L::flb + crow]) : 1.f;  // exp(log_bias); detached inside var (models.py:289-291,309)
    if (xb == 0) {
      sf[L::frho + crow] = rho;
      sf[L::fxw + 3 * crow] = xw[0];
      sf[L::fxw + 3 * crow + 1] = xw[1];
      sf[L::fxw + 3 * crow + 2] = xw[2];
    }
    {
      const float s_rho = warp_sum(xb ? 0.f : bias * rho), s_u = warp_sum(xb ? 0.f : bias * u);
      if (lane == 0) {
        sf[L::fred + warp * 2] = s_rho;
        sf[L::fred + warp * 2 + 1] = s_u;
      }
    }
    tick(4);
    pixel_barrier();
    tick(6);
    float m_pix = 0.f, q_pix = 0.f;
    {
      const int w0 = (warp / wpp) * wpp;
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
    float d_rho = ck * d_vhat * invS * bias;
    const float d_lv = SIGMA ? (bias * u * invS) * ck * 2.f * r * d_var : 0.f;
    // v_out path + biasReg = mean(log_bias)^2 over the whole batch (weight w_bias)
    const float d_lb = BIAS ? ck * d_vhat * invS * bias * rho + cfg.w_bias * 2.f * bias_mean * invB * invS : 0.f;
    if (j == 0 && xb == 0) {
      loss_d += 0.5f * e * e / var * invB;
      if (SIGMA || cfg.slice_variance) loss_s += 0.5f * logf(var) * invB;
      if (a.v_out) a.v_out[p] = vhat;
      if (cfg.slice_scale) red_add(a.g_c + k, m_pix * d_vhat);
      if (cfg.slice_variance) red_add(a.g_lvs + k, evs * d_var);
    }
    if (cfg.image_reg) {
      const int tp = (crow & ~(S - 1)) + (S - 1 - j);
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
    // dL/dz tile: sigma pass first carries dL/d(log_var) in column 0; the density pass carries dz0 (+ sigma's dz)
    __syncwarp();  // both lanes of every pair have consumed z0 / log_var
    if (xb == 0) {
      sf[L::flv + crow] = dz0;  // parked for the density pass when SIGMA (flv is dead: lv was read above by this pair only)
      const float g0 = SIGMA ? d_lv * gscale : dz0;
      *reinterpret_cast<uint4*>(gt + L::tg + umma::tile_off(srow, 0, 16)) =
          make_uint4((uint32_t)__half_as_ushort(__float2half_rn(g0)), 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(gt + L::tg + umma::tile_off(srow, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
      if (BIAS) {
        *reinterpret_cast<uint4*>(gt + L::tgb + umma::tile_off(srow, 0, 16)) =
            make_uint4((uint32_t)__half_as_ushort(__float2half_rn(fminf(fmaxf(d_lb * gscale, -65504.f), 65504.f))), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(gt + L::tgb + umma::tile_off(srow, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    if (wide) cta_barrier_all();  // the partner group has finished reading this group's rho / xw rows
    tick(4);
    publish();

    // ================= phase 3: backward on tcgen05 (b_net rides in the sigma rounds, second TMEM region) =================
    if (SIGMA) {
      if (issuer) {
        umma::fence_after_sync();
        mma_wgrad(L::c_wso, s_tsh, s_tg, RG16, 16);       // dWso^T += Hs^T G
        mma_dgrad(s_tg, RG16, s_wso, RG64, 16, 64);        // dHs = G Wso
        if (BIAS) {
          mma_wgrad(L::c_wbo, s_tbh, s_tgb, RG16, 16);            // dWbo^T += Hb^T Gb
          mma_dgrad_d(td2, s_tgb, RG16, s_wbo, RG64, 16, 64);     // dHb = Gb Wbo
        }
        umma::commit(mbar);
      }
      wait_mma();
      epi_mask_store(td + tlane + 32 * half, gt + L::tsh, erow, 32 * half);
      if (BIAS) epi_mask_store(td2 + tlane + 32 * half, gt + L::tbh, erow, 32 * half);
      publish();
      if (issuer) {
        umma::fence_after_sync();
        mma_wgrad(L::c_ws0, s_tsh, s_tsx, RG32, 32);       // dWs0 += dZs^T [se | z]
        mma_dgrad(s_tsh, RG64, s_ws0, RG32, 64, 32);       // d[se | z] = dZs Ws0
        if (BIAS) {
          mma_wgrad(L::c_wb0, s_tbh, s_tbx, RG32, 32);            // dWb0 += dZb^T [se | pe_bias | 0]
          mma_dgrad_d(td2, s_tbh, RG64, s_wb0, RG32, 64, 32);     // d[se | pe_bias | 0] = dZb Wb0
        }
        umma::commit(mbar);
      }
      wait_mma();
      {
        uint32_t d[16];
        umma::tmem_ld16(td + tlane + 16 * half, d);
        uint32_t db[16] = {};
        if (BIAS) umma::tmem_ld16(td2 + tlane + 16 * half, db);
        umma::tmem_ld_wait();
        if (half == 0) {  // d(slice embedding): column sums over the warp's 32 rows (one pixel, one slice), both heads
          const int64_t pe = (tile * kGR + erow) >> a.log2S;
          const int ke = (int)a.slice_idx[pe];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float s = warp_sum(__uint_as_float(d[c]) + (BIAS ? __uint_as_float(db[c]) : 0.f));
            if (lane == c) red_add(a.g_se + (size_t)ke * 16 + c, s * inv_gscale);
          }
        } else {
          if (BIAS) {  // b_net columns 16..23: dL/d(pe_bias), parked until the density pass has produced dL/d(features)
            const int nb = cfg.n_levels_bias;
            float* pb = reinterpret_cast<float*>(gt + L::dxb) + erow * 8;
            float gb[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) gb[c] = (c >> 1) < nb ? __uint_as_float(db[c]) : 0.f;
            *reinterpret_cast<float4*>(pb) = make_float4(gb[0], gb[1], gb[2], gb[3]);
            *reinterpret_cast<float4*>(pb + 4) = make_float4(gb[4], gb[5], gb[6], gb[7]);
          }  // dL/dz from sigma_net (+ dz0 of the render path in column 0) -> the density pass's G tile
          float g[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) g[c] = __uint_as_float(d[c]);
          g[0] += sf[L::flv + grp * kGR + erow];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint4 v;
            uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __half2 h = __floats2half2_rn(g[8 * i + 2 * q], g[8 * i + 2 * q + 1]);
              pv[q] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(gt + L::tg + umma::tile_off(erow, 8 * i, 16)) = v;
          }
        }
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
      epi_mask_store(td + tlane + 32 * half, gt + L::th + (size_t)(DEPTH - 1) * 128 * 64 * 2, erow, 32 * half);
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
      epi_mask_store(td + tlane + 32 * half, gt + L::th + (size_t)(l - 1) * 128 * 64 * 2, erow, 32 * half);
      publish();
    }
    if (issuer) {
      umma::fence_after_sync();
      mma_wgrad(L::c_w0, s_th, s_tx, RG32, 32);            // dW0 += dZ_0^T X
      mma_dgrad(s_th, RG64, s_w0, RG32, 64, 32);           // dX = dZ_0 W0
      umma::commit(mbar);
    }
    wait_mma();
    }
    acc_on = 1u;
    // dL/d(features): fp32 [128][32], feature pair (2l, 2l+1) of row r at r*32 + ((2l) ^ ((r & 15) << 1)) -- 8-byte
    // accesses, conflict-free for both the row-per-lane epilogue writes and the sample-pair reads of the scatter
    float* sdx = reinterpret_cast<float*>(gt + (BIAS ? L::tbh : (L::alias_dx ? L::th + (size_t)(DEPTH - 1) * 128 * 64 * 2 : L::dx)));
    if (!(a.ablate & 4u)) {
      uint32_t d[16];
      umma::tmem_ld16(td + tlane + 16 * half, d);
      umma::tmem_ld_wait();
      if (BIAS && half == 0) {  // + the bias head's gradient w.r.t. the first 8 encoded features
        const float* pb = reinterpret_cast<const float*>(gt + L::dxb) + erow * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) d[c] = __float_as_uint(__uint_as_float(d[c]) + pb[c]);
      }
#pragma unroll
      for (int c = 0; c < 16; c += 2)
        *reinterpret_cast<float2*>(&sdx[erow * 32 + ((16 * half + c) ^ ((erow & 15) << 1))]) =
            make_float2(__uint_as_float(d[c]), __uint_as_float(d[c + 1]));
    }
    tick(3);
    umma::fence_before_sync();
    group_barrier(grp, kGT);
    tick(1);
    // ---- scatter into the table gradient (+ pose gradient), lane pair = sample ----
    auto fetch = [&](int l) {
      return *reinterpret_cast<const float2*>(&sdx[srow * 32 + ((2 * l) ^ ((srow & 15) << 1))]);
    };
    if (cfg.pose_grad) {
      float gx[3];
      scatter_warp<true>(xn, lt, cfg.grid.n_levels, a.table, fetch, inv_gscale, a.g_table, gx, slow);
      float gwd[3], part[12];
#pragma unroll
      for (int i = 0; i < 3; ++i) gwd[i] = xb ? 0.f : gx[i] / (cfg.bbox_hi[i] - cfg.bbox_lo[i]);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int q = 0; q < 3; ++q) part[i * 3 + q] = gwd[i] * y[q];  // dL/dR
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
          red_add(a.g_axisangle + (size_t)k * 6 + q, gwv[q]);
          red_add(a.g_axisangle + (size_t)k * 6 + 3 + q, part[9 + q]);
        }
      }
    } else {
      float gx[3];
      scatter_warp<false>(xn, lt, cfg.grid.n_levels, a.table, fetch, inv_gscale, a.g_table, gx, slow);
    }
    tick(5);
    group_barrier(grp, kGT);  // the group's tiles (incl. the aliased dX slot) are free for the next tile
  }
  if (TIMED) {
    tick(1);
    if (lane == 0 && a.timers)
      for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a.timers) + i, (unsigned long long)t_acc[i]);
  }

  // ================= epilogue: weight gradients TMEM -> global, losses =================
  if (acc_on && gw == 0 && lane == 0) grp_ran[grp] = 1u;
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  {
    // NG = 2: lanes [0,16) of a quarter hold group 0's accumulator rows 16 q4 + lane, lanes [16,32) group 1's;
    // NG = 3: lanes [0,16) hold the shared accumulators, lanes [16,32) nothing
    const int q = warp & 3, part = warp >> 2;  // kThreads / 128 column partitions across the warps
    const int arow = 16 * q + (lane & 15), agrp = lane >> 4;
    const bool live = NG == 3 ? (agrp == 0 && (grp_ran[0] | grp_ran[1] | grp_ran[2]) != 0) : grp_ran[agrp] != 0;
    const uint32_t tq = tm + ((uint32_t)(32 * q) << 16);
    float* gd = a.g_mlp + a.off_density;
    float* gs = a.g_mlp + a.off_sigma;
    int chunk = 0;
    auto flush = [&](uint32_t col, int ncols, float* dst, int ld, bool transposed) {
      for (int c0 = 0; c0 < ncols; c0 += 16, ++chunk) {
        if ((chunk % (kThreads / 128)) != part) continue;
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
    if (BIAS) {
      float* gb = a.g_mlp + a.off_bias;
      flush(L::c_wb0, 32, gb, 32, false);
      flush(L::c_wbo, 16, gb + 64 * 32, 64, true);
    }
  }
  loss_d = warp_sum(loss_d);
  loss_s = warp_sum(loss_s);
  loss_i = warp_sum(loss_i);
  if (lane == 0) {
    red_add(a.losses + 0, loss_d);
    red_add(a.losses + 1, loss_s);
    red_add(a.losses + 3, loss_i);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 512);
}

template <int DEPTH, bool SIGMA, bool BIAS = false, int NG = 2>
int launch_tc(const FusedArgs& a_in, cudaStream_t st) {
  using L = TcLayout<DEPTH, SIGMA, BIAS, NG>;
  constexpr int kNGroups = NG, kThreads = 256 * NG;
  static_assert(L::bytes <= 227 * 1024, "shared memory");
  // N1 (north star: "TMA-staged hash tables in shared memory"): as many leading DENSE levels as fit beside the tiles
  FusedArgs a = a_in;
  {
    const nsv_grid_meta& m = a.cfg.grid;
    const size_t room = 227 * 1024 - L::bytes;
    int n = 0;
    while (n < m.n_levels && !m.hashed[n] && (size_t)m.offset[n + 1] * 4 <= room) ++n;
    // default: at most 2 levels.  Measured on config 2 (profiles/r02_kernelA_smem_levels.txt): 0 / 1 / 2 / 3 staged levels =
    // 0.7116 / 0.7137 / 0.7117 / 0.7158 ms -- the coarse levels' loads were L1 hits already (the table-load ablation of
    // round 1 is worth 0.03 ms in total), so staging buys nothing and a third level costs shared-memory bandwidth.
    a.smem_levels = a_in.smem_levels < 0 ? (n < 2 ? n : 2) : (a_in.smem_levels < n ? a_in.smem_levels : n);
    if (!a.fast || (m.n_levels & 3)) a.smem_levels = 0;  // the generic loops address the table globally
  }
  a.smem_table_bytes = a.smem_levels > 0 ? a.cfg.grid.offset[a.smem_levels] * 4u : 0u;
  const size_t smem_total = L::bytes + a.smem_table_bytes;
  cudaError_t e = cudaFuncSetAttribute(inr_train_tc_kernel<DEPTH, SIGMA, false, BIAS, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total);
  if (e != cudaSuccess) {
    set_error("nsv_inr_train_step(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return (int)e;
  }
  const int64_t ctas = (a.B * (int64_t)a.S / kGR + kNGroups - 1) / kNGroups;
  const int grid = (int)(ctas < num_sms() ? ctas : num_sms());
  if (a.timers && DEPTH == 3 && !SIGMA && NG == 2) {  // profiling build of the config-2 instantiation
    cudaFuncSetAttribute(inr_train_tc_kernel<3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total);
    inr_train_tc_kernel<3, false, true><<<grid, kThreads, smem_total, st>>>(a);
  } else {
    inr_train_tc_kernel<DEPTH, SIGMA, false, BIAS, NG><<<grid, kThreads, smem_total, st>>>(a);
  }
  if (int err = check_launch("nsv_inr_train_step(tcgen05)")) return err;
  inr_finalize_kernel<<<1, 256, 0, st>>>(a.logit_coef, a.g_c, a.losses, a.n_slices, a.cfg.slice_scale, a.cfg.image_reg, a.cfg.delta,
                                         a.cfg.n_levels_bias);
  return check_launch("nsv_inr_train_step(finalize)");
}

}  // namespace

int launch_train_tc(const FusedArgs& a, cudaStream_t st) {
  const nsv_inr_config& c = a.cfg;
  // 256-sample pixels span both groups of a CTA: they must march through the same tiles
  if (c.width != kW || c.depth < 1 || c.depth > 3 || (c.pixel_variance && c.depth != 1) || (a.B * (int64_t)a.S) % (kGR * 2) != 0) {
    set_error("nsv_inr_train_step: no tcgen05 instantiation for width=%d depth=%d (needs width 64, depth 1..3, B*S %% 256 == 0)", c.width, c.depth);
    return NSV_EUNSUPPORTED;
  }
  if (c.n_levels_bias) {
    if (!c.pixel_variance || c.n_levels_bias > 4) {
      set_error("nsv_inr_train_step: the fused bias-field head needs the sigma_net heads on (pixel variance, depth 1) and n_levels_bias <= 4");
      return NSV_EUNSUPPORTED;
    }
    return launch_tc<1, true, true>(a, st);
  }
  if (c.pixel_variance) return launch_tc<1, true>(a, st);
  if (a.tc_groups == 3 && a.S <= kGR && !a.timers) {  // three 128-sample groups per CTA: density-only heads, pixels of <= 128 samples
    if (c.depth == 1) return launch_tc<1, false, false, 3>(a, st);
    if (c.depth == 2) return launch_tc<2, false, false, 3>(a, st);
    return launch_tc<3, false, false, 3>(a, st);
  }
  if (c.depth == 1) return launch_tc<1, false>(a, st);
  if (c.depth == 2) return launch_tc<2, false>(a, st);
  return launch_tc<3, false>(a, st);
}

}  // namespace fused
}  // namespace nsv
