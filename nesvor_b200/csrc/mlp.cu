// mlp.cu -- standalone fully fused MLP (forward, backward) on fp16 tensor-core fragments.
// C-ABI replacement for tcnn.Network(CutlassMLP) behind build_network's fp16 branch
// (nesvor/nesvor/models.py:28-41): ReLU hidden layers, linear output, no biases.  The reference
// runs one CUTLASS GEMM per layer with activations round-tripping HBM; here a CTA keeps a 256-row
// tile and all weights in shared memory and walks every layer (and, backward, every dgrad+wgrad)
// without leaving the SM.  Weight gradients stay in registers across the persistent tile loop and
// are flushed once per CTA.
#include "mlp_mma.cuh"

namespace nsv {
namespace {

constexpr int kRows = 256, kWarps = 8, kThreadsMlp = 256, kOut = 16;

template <int IN, int W>
struct MlpSmem {
  static constexpr int ldx = IN + kPad, ldh = W + kPad, ldo = kOut + kPad;
  static size_t weights_halves(int nh) { return (size_t)W * ldx + (size_t)(nh - 1) * W * ldh + (size_t)kOut * ldh; }
};

// global [N][cols] fp16 -> shared [kRows][ld]; rows past N are zero
__device__ __forceinline__ void load_tile(__half* dst, int ld, const __half* __restrict__ src, int cols, int64_t row0, int64_t N) {
  const int vpr = cols / 8;
  for (int i = threadIdx.x; i < kRows * vpr; i += blockDim.x) {
    const int r = i / vpr, v = i % vpr;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (row0 + r < N) val = __ldg(reinterpret_cast<const uint4*>(src + (row0 + r) * cols) + v);
    *reinterpret_cast<uint4*>(dst + (size_t)r * ld + v * 8) = val;
  }
}
__device__ __forceinline__ void store_tile(__half* __restrict__ dst, int cols, const __half* src, int ld, int64_t row0, int64_t N) {
  const int vpr = cols / 8;
  for (int i = threadIdx.x; i < kRows * vpr; i += blockDim.x) {
    const int r = i / vpr, v = i % vpr;
    if (row0 + r < N) *(reinterpret_cast<uint4*>(dst + (row0 + r) * cols) + v) = *reinterpret_cast<const uint4*>(src + (size_t)r * ld + v * 8);
  }
}

template <int IN, int W>
__global__ void __launch_bounds__(kThreadsMlp, 1)
    mlp_fwd_kernel(const __half* __restrict__ x, const __half* __restrict__ weights, __half* __restrict__ out,
                   __half* __restrict__ hidden, int64_t N, int nh) {
  using L = MlpSmem<IN, W>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* sW0 = reinterpret_cast<__half*>(smem_raw);
  __half* sWh = sW0 + (size_t)W * L::ldx;
  __half* sWo = sWh + (size_t)(nh - 1) * W * L::ldh;
  __half* sX = sWo + (size_t)kOut * L::ldh;   // [kRows][ldx]
  __half* sA = sX + (size_t)kRows * L::ldx;   // [kRows][ldh] staging of activations / outputs
  stage_weights(sW0, L::ldx, weights, W, IN);
  for (int l = 0; l + 1 < nh; ++l) stage_weights(sWh + (size_t)l * W * L::ldh, L::ldh, weights + (size_t)W * IN + (size_t)l * W * W, W, W);
  stage_weights(sWo, L::ldh, weights + (size_t)W * IN + (size_t)(nh - 1) * W * W, kOut, W);
  const int warp = threadIdx.x >> 5, row0 = warp * 32;
  const int64_t n_tiles = (N + kRows - 1) / kRows;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kRows;
    __syncthreads();
    load_tile(sX, L::ldx, x, IN, base, N);
    __syncthreads();
    uint32_t ain[2][IN / 16][4];
    load_a_frags<IN / 16>(ain, sX, L::ldx, row0);
    float c[2][W / 8][4];
    warp_gemm_fwd<IN / 16, W / 8>(c, ain, sW0, L::ldx);
    uint32_t ah[2][W / 16][4];
    acc_to_a<W / 8, true>(ah, c);
    for (int l = 0; l < nh; ++l) {
      if (hidden) {  // park, then stream the whole tile out coalesced
        store_a_frags<W / 16>(ah, sA, L::ldh, row0);
        __syncthreads();
        store_tile(hidden + (size_t)l * N * W, W, sA, L::ldh, base, N);
        __syncthreads();
      }
      if (l + 1 < nh) {
        warp_gemm_fwd<W / 16, W / 8>(c, ah, sWh + (size_t)l * W * L::ldh, L::ldh);
        acc_to_a<W / 8, true>(ah, c);
      }
    }
    float co[2][kOut / 8][4];
    warp_gemm_fwd<W / 16, kOut / 8>(co, ah, sWo, L::ldh);
    uint32_t ao[2][1][4];
    acc_to_a<kOut / 8, false>(ao, co);
    store_a_frags<1>(ao, sA, L::ldh, row0);
    __syncthreads();
    store_tile(out, kOut, sA, L::ldh, base, N);
  }
}

// NH = number of hidden layers (compile time so that per-layer wgrad accumulators stay in registers)
template <int IN, int W, int NH>
__global__ void __launch_bounds__(kThreadsMlp, 1)
    mlp_bwd_kernel(const __half* __restrict__ x, const __half* __restrict__ weights, const __half* __restrict__ hidden,
                   const __half* __restrict__ grad_out, __half* __restrict__ grad_x, float* __restrict__ grad_w, int64_t N) {
  using L = MlpSmem<IN, W>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* sW0 = reinterpret_cast<__half*>(smem_raw);
  __half* sWh = sW0 + (size_t)W * L::ldx;
  __half* sWo = sWh + (size_t)(NH - 1) * W * L::ldh;
  __half* sX = sWo + (size_t)kOut * L::ldh;      // [kRows][ldx]; later holds dX
  __half* sG = sX + (size_t)kRows * L::ldx;      // [kRows][ldo]
  __half* sH = sG + (size_t)kRows * L::ldo;      // [NH][kRows][ldh]; H_l, overwritten by dZ_l
  stage_weights(sW0, L::ldx, weights, W, IN);
#pragma unroll
  for (int l = 0; l + 1 < NH; ++l) stage_weights(sWh + (size_t)l * W * L::ldh, L::ldh, weights + (size_t)W * IN + (size_t)l * W * W, W, W);
  stage_weights(sWo, L::ldh, weights + (size_t)W * IN + (size_t)(NH - 1) * W * W, kOut, W);

  float acc0[WgradSplit<W, IN, kWarps>::NTW][4] = {};
  float acch[NH > 1 ? NH - 1 : 1][WgradSplit<W, W, kWarps>::NTW][4] = {};
  float acco[WgradSplit<kOut, W, kWarps>::NTW][4] = {};

  const int warp = threadIdx.x >> 5, row0 = warp * 32;
  const int64_t n_tiles = (N + kRows - 1) / kRows;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kRows;
    __syncthreads();
    load_tile(sX, L::ldx, x, IN, base, N);
    load_tile(sG, L::ldo, grad_out, kOut, base, N);
#pragma unroll
    for (int l = 0; l < NH; ++l) load_tile(sH + (size_t)l * kRows * L::ldh, L::ldh, hidden + (size_t)l * N * W, W, base, N);
    __syncthreads();

    // output layer
    __half* sHl = sH + (size_t)(NH - 1) * kRows * L::ldh;
    warp_wgrad<kOut, W, kWarps>(acco, sG, L::ldo, sHl, L::ldh, kRows);
    uint32_t ag[2][1][4];
    load_a_frags<1>(ag, sG, L::ldo, row0);
    float c[2][W / 8][4];
    warp_gemm_dgrad<1, W / 8>(c, ag, sWo, L::ldh);
    relu_mask_acc<W / 8>(c, sHl, L::ldh, row0);
    uint32_t adz[2][W / 16][4];
    acc_to_a<W / 8, false>(adz, c);
    __syncthreads();  // every warp is done reading H_last for wgrad
    store_a_frags<W / 16>(adz, sHl, L::ldh, row0);  // H_last <- dZ_last
    __syncthreads();
#pragma unroll
    for (int l = NH - 1; l >= 1; --l) {
      __half* sDz = sH + (size_t)l * kRows * L::ldh;
      __half* sHp = sH + (size_t)(l - 1) * kRows * L::ldh;
      warp_wgrad<W, W, kWarps>(acch[l - 1], sDz, L::ldh, sHp, L::ldh, kRows);
      warp_gemm_dgrad<W / 16, W / 8>(c, adz, sWh + (size_t)(l - 1) * W * L::ldh, L::ldh);
      relu_mask_acc<W / 8>(c, sHp, L::ldh, row0);
      acc_to_a<W / 8, false>(adz, c);
      __syncthreads();
      store_a_frags<W / 16>(adz, sHp, L::ldh, row0);
      __syncthreads();
    }
    warp_wgrad<W, IN, kWarps>(acc0, sH, L::ldh, sX, L::ldx, kRows);
    if (grad_x) {
      float cx[2][IN / 8][4];
      warp_gemm_dgrad<W / 16, IN / 8>(cx, adz, sW0, L::ldx);
      uint32_t ax[2][IN / 16][4];
      acc_to_a<IN / 8, false>(ax, cx);
      __syncthreads();  // wgrad of layer 0 has consumed X
      store_a_frags<IN / 16>(ax, sX, L::ldx, row0);
      __syncthreads();
      store_tile(grad_x, IN, sX, L::ldx, base, N);
    }
  }
  flush_wgrad<W, IN, kWarps>(acc0, grad_w, IN, 1.f);
#pragma unroll
  for (int l = 0; l + 1 < NH; ++l) flush_wgrad<W, W, kWarps>(acch[l], grad_w + (size_t)W * IN + (size_t)l * W * W, W, 1.f);
  flush_wgrad<kOut, W, kWarps>(acco, grad_w + (size_t)W * IN + (size_t)(NH - 1) * W * W, W, 1.f);
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* name) {
  NSV_REQUIRE(bytes <= 227 * 1024, "%s: configuration needs %zu bytes of shared memory (> 227 KB)", name, bytes);
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    return (int)e;
  }
  return NSV_OK;
}

int check_shape(const char* name, int64_t N, int n_in, int n_out, int width, int nh) {
  NSV_REQUIRE(N >= 0, "%s: N < 0", name);
  if (!((n_in == 32 || n_in == 64) && (width == 32 || width == 64) && n_out == kOut && nh >= 1 && nh <= 4)) {
    set_error("%s: unsupported shape n_in=%d (32|64) width=%d (32|64) n_out=%d (16) n_hidden=%d (1..4)", name, n_in, width, n_out, nh);
    return NSV_EUNSUPPORTED;
  }
  return NSV_OK;
}

template <int IN, int W>
int launch_fwd(const __half* x, const __half* w, __half* out, __half* hidden, int64_t N, int nh, cudaStream_t st) {
  using L = MlpSmem<IN, W>;
  const size_t bytes = (L::weights_halves(nh) + (size_t)kRows * L::ldx + (size_t)kRows * L::ldh) * sizeof(__half);
  if (int e = set_smem(mlp_fwd_kernel<IN, W>, bytes, "nsv_mlp_fwd_f16")) return e;
  const int64_t tiles = (N + kRows - 1) / kRows;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  mlp_fwd_kernel<IN, W><<<grid, kThreadsMlp, bytes, st>>>(x, w, out, hidden, N, nh);
  return check_launch("nsv_mlp_fwd_f16");
}

template <int IN, int W, int NH>
int launch_bwd(const __half* x, const __half* w, const __half* hidden, const __half* go, __half* gx, float* gw, int64_t N,
               cudaStream_t st) {
  using L = MlpSmem<IN, W>;
  const size_t bytes =
      (L::weights_halves(NH) + (size_t)kRows * L::ldx + (size_t)kRows * L::ldo + (size_t)NH * kRows * L::ldh) * sizeof(__half);
  if (int e = set_smem(mlp_bwd_kernel<IN, W, NH>, bytes, "nsv_mlp_bwd_f16")) return e;
  const int64_t tiles = (N + kRows - 1) / kRows;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  mlp_bwd_kernel<IN, W, NH><<<grid, kThreadsMlp, bytes, st>>>(x, w, hidden, go, gx, gw, N);
  return check_launch("nsv_mlp_bwd_f16");
}

template <int IN, int W>
int dispatch_bwd(int nh, const __half* x, const __half* w, const __half* hidden, const __half* go, __half* gx, float* gw, int64_t N,
                 cudaStream_t st) {
  switch (nh) {
    case 1: return launch_bwd<IN, W, 1>(x, w, hidden, go, gx, gw, N, st);
    case 2: return launch_bwd<IN, W, 2>(x, w, hidden, go, gx, gw, N, st);
    case 3: return launch_bwd<IN, W, 3>(x, w, hidden, go, gx, gw, N, st);
    default: return launch_bwd<IN, W, 4>(x, w, hidden, go, gx, gw, N, st);
  }
}

}  // namespace
}  // namespace nsv

extern "C" int nsv_mlp_fwd_f16(const void* x, const void* weights, void* out, void* hidden, int64_t N, int n_in, int n_out,
                               int width, int n_hidden, void* stream) {
  using namespace nsv;
  if (int e = check_shape("nsv_mlp_fwd_f16", N, n_in, n_out, width, n_hidden)) return e;
  NSV_REQUIRE(N == 0 || (x && weights && out), "nsv_mlp_fwd_f16: NULL pointer");
  if (N == 0) return NSV_OK;
  const __half *xh = (const __half*)x, *wh = (const __half*)weights;
  __half *oh = (__half*)out, *hh = (__half*)hidden;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_in == 32 && width == 32) return launch_fwd<32, 32>(xh, wh, oh, hh, N, n_hidden, st);
  if (n_in == 64 && width == 32) return launch_fwd<64, 32>(xh, wh, oh, hh, N, n_hidden, st);
  if (n_in == 32 && width == 64) return launch_fwd<32, 64>(xh, wh, oh, hh, N, n_hidden, st);
  return launch_fwd<64, 64>(xh, wh, oh, hh, N, n_hidden, st);
}

extern "C" int nsv_mlp_bwd_f16(const void* x, const void* weights, const void* hidden, const void* grad_out, void* grad_x,
                               float* grad_weights, int64_t N, int n_in, int n_out, int width, int n_hidden, void* stream) {
  using namespace nsv;
  if (int e = check_shape("nsv_mlp_bwd_f16", N, n_in, n_out, width, n_hidden)) return e;
  NSV_REQUIRE(N == 0 || (x && weights && hidden && grad_out && grad_weights), "nsv_mlp_bwd_f16: NULL pointer");
  if (N == 0) return NSV_OK;
  const __half *xh = (const __half*)x, *wh = (const __half*)weights, *hh = (const __half*)hidden, *gh = (const __half*)grad_out;
  __half* gx = (__half*)grad_x;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_in == 32 && width == 32) return dispatch_bwd<32, 32>(n_hidden, xh, wh, hh, gh, gx, grad_weights, N, st);
  if (n_in == 64 && width == 32) return dispatch_bwd<64, 32>(n_hidden, xh, wh, hh, gh, gx, grad_weights, N, st);
  if (n_in == 32 && width == 64) return dispatch_bwd<32, 64>(n_hidden, xh, wh, hh, gh, gx, grad_weights, N, st);
  return dispatch_bwd<64, 64>(n_hidden, xh, wh, hh, gh, gx, grad_weights, N, st);
}
