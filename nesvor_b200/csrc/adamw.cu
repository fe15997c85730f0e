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

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, __half* __restrict__ p16, int64_t n, float lr, float b1,
                                                    float b2, float eps, float wd, float step_size, float inv_sqrt_bc2,
                                                    float unscale, int zero_grad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * unscale;
    float pi = p[i];
    float mi = m[i], vi = v[i];
    pi *= 1.f - lr * wd;
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (p16) p16[i] = __float2half_rn(pi);
    if (zero_grad) g[i] = 0.f;
  }
}

}  // namespace
}  // namespace nsv

extern "C" int nsv_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* param_f16, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int step, float grad_unscale, int zero_grad,
                              void* stream) {
  using namespace nsv;
  NSV_REQUIRE(n >= 0 && step >= 1, "nsv_adamw_step: bad n / step");
  if (n == 0) return NSV_OK;
  NSV_REQUIRE(param && grad && exp_avg && exp_avg_sq, "nsv_adamw_step: NULL pointer");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const int64_t blocks = (n + 255) / 256;
  const int grid = (int)(blocks < (int64_t)num_sms() * 8 ? blocks : (int64_t)num_sms() * 8);
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, (__half*)param_f16, n, lr, beta1, beta2, eps,
                                                       weight_decay, step_size, inv_sqrt_bc2, grad_unscale, zero_grad);
  return check_launch("nsv_adamw_step");
}
