// adamw.cu -- fused AdamW over a flat fp32 segment: unscale, moment update, decoupled weight decay,
// parameter update, refresh of the fp16 copy the kernels read, and (optionally) zeroing of the
// gradient for the next iteration -- one streaming pass (16 B read + 12..18 B written per element)
// instead of torch.optim.AdamW's multi-tensor passes + GradScaler.unscale_ + zero_grad
// (nesvor/nesvor/train.py:134-165,190-197).  Arithmetic follows torch.optim.AdamW:
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= lr / (1-b1^t) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
#include "nsv_common.cuh"

namespace nsv {
namespace {

// 4 elements per thread and iteration (16-byte accesses; `n4` vectors, then a scalar tail): the pass is pure streaming
// (16 B read + 12..18 B written per element) and needs wide accesses to approach HBM bandwidth
__device__ __forceinline__ void adamw_update(float g, float& pi, float& mi, float& vi, float lr, float b1, float b2, float eps, float wd,
                                             float step_size, float inv_sqrt_bc2, float unscale) {
  g *= unscale;
  // found-inf guard, element-wise: the reference's GradScaler skips an optimiser step whose gradient holds an inf / NaN
  // (train.py:161-164,195); here an element whose gradient is not finite keeps its parameter and both moments (the fp16
  // backward operands overflow before fp32 does, and one poisoned element must not poison Adam's state for good)
  if (!(fabsf(g) <= 3.0e38f)) return;
  pi *= 1.f - lr * wd;
  mi = b1 * mi + (1.f - b1) * g;
  vi = b2 * vi + (1.f - b2) * g * g;
  const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
  pi -= step_size * (mi / denom);
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, __half* __restrict__ p16, int64_t n, float lr, float b1,
                                                    float b2, float eps, float wd, float step_size, float inv_sqrt_bc2,
                                                    float unscale, int zero_grad) {
  const int64_t n4 = n >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += nth) {
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 pv = reinterpret_cast<const float4*>(p)[i], mv = reinterpret_cast<const float4*>(m)[i], vv = reinterpret_cast<const float4*>(v)[i];
    adamw_update(gv.x, pv.x, mv.x, vv.x, lr, b1, b2, eps, wd, step_size, inv_sqrt_bc2, unscale);
    adamw_update(gv.y, pv.y, mv.y, vv.y, lr, b1, b2, eps, wd, step_size, inv_sqrt_bc2, unscale);
    adamw_update(gv.z, pv.z, mv.z, vv.z, lr, b1, b2, eps, wd, step_size, inv_sqrt_bc2, unscale);
    adamw_update(gv.w, pv.w, mv.w, vv.w, lr, b1, b2, eps, wd, step_size, inv_sqrt_bc2, unscale);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (p16) {
      const __half2 h0 = __floats2half2_rn(pv.x, pv.y), h1 = __floats2half2_rn(pv.z, pv.w);
      reinterpret_cast<uint2*>(p16)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    }
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += nth) {
    float pi = p[i], mi = m[i], vi = v[i];
    adamw_update(g[i], pi, mi, vi, lr, b1, b2, eps, wd, step_size, inv_sqrt_bc2, unscale);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (p16) p16[i] = __float2half_rn(pi);
    if (zero_grad) g[i] = 0.f;
  }
}

// ---- data-parallel step: reduce-scatter + AdamW + all-gather in ONE kernel over NVLink peer memory ----
// Every rank owns a contiguous shard [lo, hi) of the flat parameter vector.  For its shard it loads the gradient of
// every rank (its own from HBM, the others through peer pointers over NVLink / NVSwitch), sums them in rank order
// (the same order on every owner, so all replicas see bit-identical parameters), applies AdamW to its fp32 master,
// moments included, and stores the refreshed fp16 parameters -- the only copy the training kernel reads -- into the
// shadow buffer of every rank.  Per rank and step the links carry (world-1)/world x 4 B/param in and
// (world-1)/world x 2 B/param out, against 2 x (world-1)/world x 4 B for an all-reduce, and the optimiser's HBM traffic
// shrinks by 1/world; nothing is staged, no library collective is launched.  Callers bracket the launch with the
// symmetric-memory barrier (all gradients complete before, all shadows written after).
constexpr int kMaxPeers = 16;
struct PeerPtrs {
  const float* grad[kMaxPeers];
  __half* p16[kMaxPeers];
  float* p32[kMaxPeers];  // optional fp32 mirror of the elements [f32_lo, n): the per-slice parameters kernel A reads in fp32
  // optional in-kernel rendezvous (replaces the two host-launched symmetric-memory barriers around the kernel):
  // flags[r] = rank r's flag block in peer memory, 64 x u64: [0,16) "gradient of rank i complete" (written by rank i),
  // [16,32) "rank i has read every gradient and written every fp16 / fp32 copy", [32] block ticket counter, [33] time-out marker,
  // [34] device-scope "every rank's gradient is complete" (written by block 0 of this rank)
  unsigned long long* flags[kMaxPeers];
  // optional NVSwitch multicast addresses of the same three buffers (one address = the copy on every rank): with them the gradient
  // sum is ONE multimem.ld_reduce per vector (the switch adds the ranks' values: a shard's worth of inbound traffic instead of
  // world - 1 shards) and the refreshed parameters are ONE multimem.st (the switch replicates it)
  const float* mc_grad;
  __half* mc_p16;
  float* mc_p32;
};

__device__ __forceinline__ float4 multimem_sum_f32x4(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_store_f32x4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_store_f16x4(__half* mc, uint2 packed) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f16x2 [%0], {%1, %2};" ::"l"(mc), "r"(packed.x), "r"(packed.y) : "memory");
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// spins on a flag in THIS rank's memory until it reaches `want`; gives up after ~4 s of cycles (a rank that died must not
// hang the others' GPUs) and leaves a marker the host checks
__device__ __forceinline__ void wait_flag(unsigned long long* flags_mine, int slot, unsigned long long want) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flags_mine + slot) < want) {
    if (clock64() - t0 > (8ll << 30)) {
      flags_mine[33] = want;
      break;
    }
    __nanosleep(64);
  }
}

// WORLD > 0: compile-time rank count (all peer loads of an element group are in flight together); 0: run-time loop
template <int WORLD>
__global__ void __launch_bounds__(256) adamw_dp_kernel(float* __restrict__ p, const __grid_constant__ PeerPtrs peers, float* __restrict__ m,
                                                       float* __restrict__ v, int world_rt, int64_t lo, int64_t hi, int64_t f32_lo, float lr,
                                                       float b1, float b2, float eps, float wd, float step_size, float inv_sqrt_bc2,
                                                       float unscale, int rank, unsigned long long epoch, int sync_mode) {
  const int world = WORLD > 0 ? WORLD : world_rt;
  // sync_mode bit 0: rendezvous before the gradients are pulled; bit 1: rendezvous after the copies are written
  const bool rendezvous = peers.flags[0] != nullptr && (sync_mode & 2);
  if (peers.flags[0] != nullptr && (sync_mode & 1)) {
    // (1) this rank's gradient is complete (kernel A precedes this launch on the stream): tell every rank; then every block
    // waits until every rank has said so before it pulls gradients through the peer pointers
    // Only block 0 talks to the other ranks (system-scope loads of flags that peers write over NVLink cost microseconds each,
    // and 1184 blocks x 8 flags of them made the first version slower than a host-launched barrier); it then publishes
    // "everybody is ready" in a device-scope flag the other blocks of this rank wait on.
    if (threadIdx.x == 0) {
      unsigned long long* mine = peers.flags[rank];
      if (blockIdx.x == 0) {
        for (int r = 0; r < world; ++r) st_release_sys(peers.flags[r] + rank, epoch);
        for (int r = 0; r < world; ++r) wait_flag(mine, r, epoch);
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(mine + 34), "l"(epoch) : "memory");
      } else {
        const long long t0 = clock64();
        unsigned long long v;
        do {
          asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + 34) : "memory");
          if (v >= epoch) break;
          __nanosleep(32);
        } while (clock64() - t0 < (9ll << 30));
      }
    }
    __syncthreads();
  }
  // lo is a multiple of 4 (16-byte aligned vectors); the ragged tail of the last shard is handled element-wise
  const int64_t n4 = (hi - lo) >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  auto update = [&](float g, float& pi, float& mi, float& vi) {
    g *= unscale;
    if (!(fabsf(g) <= 3.0e38f)) return;  // found-inf guard (see adamw_update): identical on every owner, the sum is the same everywhere
    pi *= 1.f - lr * wd;
    mi = b1 * mi + (1.f - b1) * g;
    vi = b2 * vi + (1.f - b2) * g * g;
    pi -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  };
  for (int64_t i = tid; i < n4; i += nth) {
    const int64_t e = lo + 4 * i;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (peers.mc_grad) {
      g = multimem_sum_f32x4(peers.mc_grad + e);
    } else if (WORLD > 0) {
      float4 gr[WORLD > 0 ? WORLD : 1];
#pragma unroll
      for (int r = 0; r < WORLD; ++r) gr[r] = *reinterpret_cast<const float4*>(peers.grad[r] + e);
#pragma unroll
      for (int r = 0; r < WORLD; ++r) {
        g.x += gr[r].x;
        g.y += gr[r].y;
        g.z += gr[r].z;
        g.w += gr[r].w;
      }
    } else {
      for (int r = 0; r < world; ++r) {
        const float4 gr = *reinterpret_cast<const float4*>(peers.grad[r] + e);
        g.x += gr.x;
        g.y += gr.y;
        g.z += gr.z;
        g.w += gr.w;
      }
    }
    float4 pv = *reinterpret_cast<const float4*>(p + e), mv = *reinterpret_cast<const float4*>(m + e), vv = *reinterpret_cast<const float4*>(v + e);
    update(g.x, pv.x, mv.x, vv.x);
    update(g.y, pv.y, mv.y, vv.y);
    update(g.z, pv.z, mv.z, vv.z);
    update(g.w, pv.w, mv.w, vv.w);
    *reinterpret_cast<float4*>(p + e) = pv;
    *reinterpret_cast<float4*>(m + e) = mv;
    *reinterpret_cast<float4*>(v + e) = vv;
    const __half2 h0 = __floats2half2_rn(pv.x, pv.y), h1 = __floats2half2_rn(pv.z, pv.w);
    const uint2 packed = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    if (peers.mc_grad) {
      multimem_store_f16x4(peers.mc_p16 + e, packed);
      if (e >= f32_lo) multimem_store_f32x4(peers.mc_p32 + (e - f32_lo), pv);
      continue;
    }
    for (int r = 0; r < world; ++r) *reinterpret_cast<uint2*>(peers.p16[r] + e) = packed;
    if (e >= f32_lo)  // f32_lo and lo are multiples of 4: a vector never straddles the boundary
      for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(peers.p32[r] + (e - f32_lo)) = pv;
  }
  for (int64_t e = lo + 4 * n4 + tid; e < hi; e += nth) {
    float g = 0.f;
    for (int r = 0; r < world; ++r) g += peers.grad[r][e];
    float pi = p[e], mi = m[e], vi = v[e];
    update(g, pi, mi, vi);
    p[e] = pi;
    m[e] = mi;
    v[e] = vi;
    for (int r = 0; r < world; ++r) peers.p16[r][e] = __float2half_rn(pi);
    if (e >= f32_lo)
      for (int r = 0; r < world; ++r) peers.p32[r][e - f32_lo] = pi;
  }
  if (rendezvous) {
    // (2) the last block of this rank to finish announces "all my reads and all my writes into your copies are done" and
    // then waits for the same announcement of every rank: when the kernel ends, this rank's fp16 / fp32 copies are complete
    // (the next kernel A may read them) and nobody reads this rank's gradient any more (it may be cleared)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      unsigned long long* mine = peers.flags[rank];
      const unsigned long long ticket = atomicAdd(mine + 32, 1ull);
      if (ticket == (unsigned long long)gridDim.x - 1ull) {
        __threadfence_system();
        mine[32] = 0ull;
        for (int r = 0; r < world; ++r) st_release_sys(peers.flags[r] + 16 + rank, epoch);
        for (int r = 0; r < world; ++r) wait_flag(mine, 16 + r, epoch);
      }
    }
  }
}

}  // namespace
}  // namespace nsv

extern "C" int nsv_adamw_shard_bounds(int64_t n, int world, int rank, int64_t* lo, int64_t* hi) {
  using namespace nsv;
  NSV_REQUIRE(world >= 1 && rank >= 0 && rank < world && n >= 0 && lo && hi, "nsv_adamw_shard_bounds: bad arguments");
  const int64_t chunk = ((n + world - 1) / world + 3) / 4 * 4;
  *lo = chunk * rank < n ? chunk * rank : n;
  *hi = *lo + chunk < n ? *lo + chunk : n;
  return NSV_OK;
}

static int adamw_step_dp_impl(float* param, const void* const* peer_grads, float* exp_avg, float* exp_avg_sq,
                             void* const* peer_param_f16, int world, int rank, int64_t n, float lr, float beta1, float beta2,
                             float eps, float weight_decay, int step, float grad_unscale, int64_t f32_lo,
                             void* const* peer_param_f32, void* const* peer_flags, uint64_t epoch, int sync_mode, void* stream,
                             const void* mc_grad = nullptr, void* mc_param_f16 = nullptr, void* mc_param_f32 = nullptr) {
  using namespace nsv;
  NSV_REQUIRE(n >= 0 && step >= 1, "nsv_adamw_step_dp: bad n / step");
  NSV_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "nsv_adamw_step_dp: world must be 1..16, rank in [0, world)");
  NSV_REQUIRE(param && peer_grads && exp_avg && exp_avg_sq && peer_param_f16, "nsv_adamw_step_dp: NULL pointer");
  if (!peer_param_f32) f32_lo = n;  // no fp32 mirror
  NSV_REQUIRE(f32_lo >= 0 && f32_lo % 4 == 0, "nsv_adamw_step_dp: f32_lo must be a non-negative multiple of 4");
  PeerPtrs pp;
  pp.mc_grad = (const float*)mc_grad;
  pp.mc_p16 = (__half*)mc_param_f16;
  pp.mc_p32 = (float*)mc_param_f32;
  if (mc_grad) {
    NSV_REQUIRE(mc_param_f16 && (f32_lo >= n || mc_param_f32), "nsv_adamw_step_dp_mc: NULL multicast pointer");
    NSV_REQUIRE(((uintptr_t)mc_grad | (uintptr_t)mc_param_f32) % 16 == 0 && (uintptr_t)mc_param_f16 % 8 == 0,
                "nsv_adamw_step_dp_mc: multicast buffers must be 16-byte aligned (fp16 copy: 8-byte)");
  }
  for (int r = 0; r < kMaxPeers; ++r) {
    pp.grad[r] = r < world ? (const float*)peer_grads[r] : nullptr;
    pp.p16[r] = r < world ? (__half*)peer_param_f16[r] : nullptr;
    pp.p32[r] = (r < world && peer_param_f32) ? (float*)peer_param_f32[r] : nullptr;
    pp.flags[r] = (r < world && peer_flags) ? (unsigned long long*)peer_flags[r] : nullptr;
    NSV_REQUIRE(r >= world || !peer_flags || pp.flags[r], "nsv_adamw_step_dp_sync: NULL flag pointer");
    NSV_REQUIRE(r >= world || (pp.grad[r] && pp.p16[r]), "nsv_adamw_step_dp: NULL peer pointer");
    NSV_REQUIRE(r >= world || f32_lo >= n || (pp.p32[r] && (uintptr_t)pp.p32[r] % 16 == 0), "nsv_adamw_step_dp: NULL / unaligned fp32 mirror pointer");
  }
  int64_t lo = 0, hi = 0;
  nsv_adamw_shard_bounds(n, world, rank, &lo, &hi);
  if (hi <= lo && !peer_flags) return NSV_OK;  // with the in-kernel rendezvous an empty shard still takes part in it
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const int64_t blocks = ((hi - lo) / 4 + 255) / 256 + 1;
  const int grid = (int)(blocks < (int64_t)num_sms() * 8 ? blocks : (int64_t)num_sms() * 8);
#define NSV_DP_LAUNCH(W)                                                                                                            \
  adamw_dp_kernel<W><<<grid, 256, 0, (cudaStream_t)stream>>>(param, pp, exp_avg, exp_avg_sq, world, lo, hi, f32_lo, lr, beta1, beta2, eps, weight_decay, \
                                                             step_size, inv_sqrt_bc2, grad_unscale, rank, (unsigned long long)epoch, sync_mode)
  if (world == 2) NSV_DP_LAUNCH(2);
  else if (world == 4) NSV_DP_LAUNCH(4);
  else if (world == 8) NSV_DP_LAUNCH(8);
  else NSV_DP_LAUNCH(0);
#undef NSV_DP_LAUNCH
  return check_launch("nsv_adamw_step_dp");
}

extern "C" int nsv_adamw_step_dp(float* param, const void* const* peer_grads, float* exp_avg, float* exp_avg_sq,
                                 void* const* peer_param_f16, int world, int rank, int64_t n, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, int step, float grad_unscale, int64_t f32_lo,
                                 void* const* peer_param_f32, void* stream) {
  return adamw_step_dp_impl(param, peer_grads, exp_avg, exp_avg_sq, peer_param_f16, world, rank, n, lr, beta1, beta2, eps, weight_decay,
                            step, grad_unscale, f32_lo, peer_param_f32, nullptr, 0, 0, stream);
}

extern "C" int nsv_adamw_step_dp_mc(float* param, const void* const* peer_grads, float* exp_avg, float* exp_avg_sq,
                                    void* const* peer_param_f16, int world, int rank, int64_t n, float lr, float beta1, float beta2,
                                    float eps, float weight_decay, int step, float grad_unscale, int64_t f32_lo,
                                    void* const* peer_param_f32, const void* mc_grad, void* mc_param_f16, void* mc_param_f32,
                                    void* stream) {
  using namespace nsv;
  NSV_REQUIRE(mc_grad, "nsv_adamw_step_dp_mc: needs the multicast address of the gradient buffers");
  return adamw_step_dp_impl(param, peer_grads, exp_avg, exp_avg_sq, peer_param_f16, world, rank, n, lr, beta1, beta2, eps, weight_decay,
                            step, grad_unscale, f32_lo, peer_param_f32, nullptr, 0, 0, stream, mc_grad, mc_param_f16, mc_param_f32);
}

extern "C" int nsv_adamw_step_dp_sync(float* param, const void* const* peer_grads, float* exp_avg, float* exp_avg_sq,
                                      void* const* peer_param_f16, int world, int rank, int64_t n, float lr, float beta1, float beta2,
                                      float eps, float weight_decay, int step, float grad_unscale, int64_t f32_lo,
                                      void* const* peer_param_f32, void* const* peer_flags, uint64_t epoch, int sync_mode, void* stream) {
  using namespace nsv;
  NSV_REQUIRE(peer_flags && epoch > 0, "nsv_adamw_step_dp_sync: needs the ranks' flag blocks and a positive, strictly increasing epoch");
  NSV_REQUIRE(sync_mode >= 1 && sync_mode <= 3, "nsv_adamw_step_dp_sync: sync_mode is 1 (rendezvous before), 2 (after) or 3 (both)");
  return adamw_step_dp_impl(param, peer_grads, exp_avg, exp_avg_sq, peer_param_f16, world, rank, n, lr, beta1, beta2, eps, weight_decay,
                            step, grad_unscale, f32_lo, peer_param_f32, peer_flags, epoch, sync_mode, stream);
}

extern "C" int nsv_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* param_f16, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int step, float grad_unscale, int zero_grad,
                              void* stream) {
  using namespace nsv;
  NSV_REQUIRE(n >= 0 && step >= 1, "nsv_adamw_step: bad n / step");
  if (n == 0) return NSV_OK;
  NSV_REQUIRE(param && grad && exp_avg && exp_avg_sq, "nsv_adamw_step: NULL pointer");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  NSV_REQUIRE(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0 && (uintptr_t)param_f16 % 8 == 0,
              "nsv_adamw_step: buffers must be 16-byte aligned (fp16 copy: 8-byte)");
  const int64_t blocks = (n / 4 + 255) / 256 + 1;
  const int grid = (int)(blocks < (int64_t)num_sms() * 8 ? blocks : (int64_t)num_sms() * 8);
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, (__half*)param_f16, n, lr, beta1, beta2, eps,
                                                       weight_decay, step_size, inv_sqrt_bc2, grad_unscale, zero_grad);
  return check_launch("nsv_adamw_step");
}
