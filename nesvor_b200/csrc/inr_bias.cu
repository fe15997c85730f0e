// inr_bias.cu -- batch mean of log_bias, the one quantity of a NeSVoR iteration that couples all samples.
//
// biasReg = mean(log_bias)^2 (nesvor/nesvor/models.py:323) puts the same cotangent 2 * w * mean / (B S) on every
// sample's log_bias, so the mean over the WHOLE batch must be known before kernel A back-propagates its first tile.
// This forward-only pre-pass recomputes exactly what kernel A will feed into b_net (models.py:247-258,344-347):
// the sample positions (same noise: caller tensor or Philox(seed, offset + sample index)), the first
// n_levels_bias <= 4 hash-grid levels (all dense and a few thousand entries: L1-resident), the slice embedding,
// and the 24 -> 64 -> 1 ReLU MLP with kernel A's rounding points (fp16 inputs / weights / hidden activations, fp32
// accumulation).  Thread = sample on the CUDA cores; the slice-embedding part of the first layer (16 of the 24 inputs)
// is shared by all samples of a pixel and computed once per pixel in shared memory, which leaves ~0.6 kFLOP per
// sample -- a few percent of kernel A's time, with none of its tile machinery.  The result is added to *out_mean (= losses[4] of nsv_inr_grads), which the caller may
// all-reduce over data-parallel ranks before nsv_inr_train_step consumes it.
#include "inr_common.cuh"

namespace nsv {
namespace fused {
namespace {

constexpr int kBiasThreads = 256, kBiasIn = 24, kW64 = 64;

__global__ void __launch_bounds__(kBiasThreads) inr_bias_mean_kernel(const __grid_constant__ FusedArgs a, float* __restrict__ out_mean) {
  __shared__ __align__(16) float w0[kW64 * kBiasIn];  // first layer, fp16 values widened once: [64][24]
  __shared__ float wo[kW64];                          // row 0 of the [16][64] output layer
  __shared__ float red[kBiasThreads / 32];
  __shared__ float pse[8 * kW64];                     // slice-embedding part of the pre-activations, per pixel of the block's chunk
  const nsv_inr_config& cfg = a.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __half* wb = a.mlp + a.off_bias;
  for (int i = tid; i < kW64 * kBiasIn; i += kBiasThreads) w0[i] = __half2float(wb[(i / kBiasIn) * kIn + i % kBiasIn]);
  if (tid < kW64) wo[tid] = __half2float(wb[kW64 * kIn + tid]);
  __syncthreads();
  const int nb = cfg.n_levels_bias;
  const int64_t N = a.B * (int64_t)a.S;
  float acc = 0.f;
  // a block walks chunks of 256 consecutive samples = 256 / S whole pixels (or part of one): the 16 slice-embedding
  // inputs are shared by all samples of a pixel, so their 64 partial pre-activations are computed once per pixel
  const int npix = a.S >= kBiasThreads ? 1 : kBiasThreads >> a.log2S;
  for (int64_t base = (int64_t)blockIdx.x * kBiasThreads; base < N; base += (int64_t)gridDim.x * kBiasThreads) {
    const int64_t p0 = base >> a.log2S;
    __syncthreads();  // the previous chunk's readers of pse are done
    for (int i = tid; i < npix * kW64; i += kBiasThreads) {
      const int pix = i / kW64, j = i % kW64;
      float s = 0.f;
      if (p0 + pix < a.B) {
        const float* se = a.slice_embedding + (size_t)a.slice_idx[p0 + pix] * 16;
#pragma unroll
        for (int c = 0; c < 16; ++c) s = fmaf(w0[j * kBiasIn + c], __half2float(__float2half_rn(__ldg(se + c))), s);
      }
      pse[i] = s;
    }
    __syncthreads();
    const int64_t sidx = base + tid;
    if (sidx >= N) continue;
    const int64_t p = sidx >> a.log2S;
    const float* ps = pse + (int)(p - p0) * kW64;
    const int k = (int)a.slice_idx[p];
    // ---- sample position: identical arithmetic to kernel A's phase 0 ----
    float ax[6], R[9], y[3], xn[3];
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
        const float xw = R[i * 3] * y[0] + R[i * 3 + 1] * y[1] + R[i * 3 + 2] * y[2];
        xn[i] = (xw - cfg.bbox_lo[i]) / (cfg.bbox_hi[i] - cfg.bbox_lo[i]);
      }
    }
    // ---- b_net input [slice embedding (16) | pe_bias (8)], rounded to fp16 like kernel A's operand tile ----
    float x[8];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      float f0 = 0.f, f1 = 0.f;
      if (l < nb) {
        const LevelGeom lv = level_geom(cfg.grid, l);
        uint32_t g[3];
        float w[3];
        level_pos(xn, lv.scale, g, w);
        // same summation order as the lane pair of kernel A: 4 (y,z) corners per x-corner, then the pair sum
        float part0[2] = {0.f, 0.f}, part1[2] = {0.f, 0.f};
#pragma unroll
        for (int xb = 0; xb < 2; ++xb) {
          uint32_t e[4];
          corner_entries(lv, g[0] + xb, g[1], g[2], e);
          const float wx = xb ? w[0] : 1.f - w[0];
          const float wy0 = wx * (1.f - w[1]), wy1 = wx * w[1];
          const float wq[4] = {wy0 * (1.f - w[2]), wy1 * (1.f - w[2]), wy0 * w[2], wy1 * w[2]};
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = load_pair(a.table, e[q]);
            a0 = fmaf(wq[q], f.x, a0);
            a1 = fmaf(wq[q], f.y, a1);
          }
          part0[xb] = a0;
          part1[xb] = a1;
        }
        const __half2 h = __floats2half2_rn(part0[0] + part0[1], part1[0] + part1[1]);
        const float2 hf = __half22float2(h);
        f0 = hf.x;
        f1 = hf.y;
      }
      x[2 * l] = f0;
      x[2 * l + 1] = f1;
    }
    // ---- 24 -> 64 (ReLU, fp16 activations) -> 1 ----
    float lb = 0.f;
#pragma unroll 4
    for (int j = 0; j < kW64; ++j) {
      const float4* wr = reinterpret_cast<const float4*>(w0 + j * kBiasIn + 16);
      float s = ps[j];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 wv = wr[c];
        s = fmaf(wv.x, x[4 * c], s);
        s = fmaf(wv.y, x[4 * c + 1], s);
        s = fmaf(wv.z, x[4 * c + 2], s);
        s = fmaf(wv.w, x[4 * c + 3], s);
      }
      const float h = __half2float(__float2half_rn(fmaxf(s, 0.f)));
      lb = fmaf(wo[j], h, lb);
    }
    acc += lb;
  }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kBiasThreads / 32; ++i) s += red[i];
    red_add(out_mean, s / (float)N);
  }
}

}  // namespace

int launch_bias_mean(const FusedArgs& a, float* out_mean, cudaStream_t st) {
  const int64_t N = a.B * (int64_t)a.S;
  const int64_t blocks = (N + kBiasThreads - 1) / kBiasThreads;
  const int64_t cap = (int64_t)num_sms() * 8;
  inr_bias_mean_kernel<<<(int)(blocks < cap ? blocks : cap), kBiasThreads, 0, st>>>(a, out_mean);
  return check_launch("nsv_inr_bias_mean");
}

}  // namespace fused
}  // namespace nsv

extern "C" int nsv_inr_bias_mean(const nsv_inr_config* cfg, const nsv_inr_params* prm, const float* xyz, const int64_t* slice_idx,
                                 const float* noise, uint64_t seed, uint64_t offset, float* out_mean, int64_t B, int S, void* stream) {
  using namespace nsv;
  using namespace nsv::fused;
  NSV_REQUIRE(cfg && prm && xyz && slice_idx && out_mean, "nsv_inr_bias_mean: NULL pointer");
  NSV_REQUIRE(prm->table_f16 && prm->mlp_f16 && prm->axisangle && prm->psf_sigma && prm->slice_embedding,
              "nsv_inr_bias_mean: NULL parameter buffer (the bias-field head reads the table, b_net, poses and the slice embedding)");
  NSV_REQUIRE(B > 0 && S > 0 && (S & (S - 1)) == 0, "nsv_inr_bias_mean: B must be positive and n_samples a power of two");
  if (cfg->n_levels_bias < 1 || cfg->n_levels_bias > 4 || cfg->n_levels_bias > cfg->grid.n_levels || cfg->grid.n_features != 2 ||
      cfg->width != 64 || cfg->depth != 1 || cfg->n_features_slice != 16) {
    set_error("nsv_inr_bias_mean: instantiated for 1 <= n_levels_bias <= 4, F = 2, width 64, depth 1, n_features_slice 16");
    return NSV_EUNSUPPORTED;
  }
  int64_t off[3];
  if (nsv_inr_mlp_layout(cfg, off) < 0) return NSV_EINVAL;
  FusedArgs a{};
  a.cfg = *cfg;
  a.table = (const __half*)prm->table_f16;
  a.mlp = (const __half*)prm->mlp_f16;
  a.axisangle = prm->axisangle;
  a.psf_sigma = prm->psf_sigma;
  a.slice_embedding = prm->slice_embedding;
  a.n_slices = prm->n_slices;
  a.xyz = xyz;
  a.slice_idx = slice_idx;
  a.noise = noise;
  a.seed = seed;
  a.offset = offset;
  a.B = B;
  a.S = S;
  while ((1 << a.log2S) < S) ++a.log2S;
  a.off_bias = off[2];
  return launch_bias_mean(a, out_mean, (cudaStream_t)stream);
}
