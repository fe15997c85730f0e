// Micro-benchmark: throughput of random gathers / reductions with kernel A's access pattern
// (lane pair = 2 adjacent entries, 16 pairs per warp instruction, L2-resident working set).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/red_bench tools/red_bench.cu && tools/red_bench
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// mode 0: red.v2.f32 (8 B / lane), 1: red.f16x2 (4 B / lane), 2: ld 4 B, 3: red.v4.f32 by even lanes (pair-combined via shfl)
// 4: red.f32 scalar x2
template <int MODE>
__global__ void __launch_bounds__(512) k(float* gf, __half2* gh, uint32_t n_entries, int iters, float* sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t pair = tid >> 1, xb = tid & 1;
  float acc = 0.f;
  uint32_t s = mix(pair * 2654435761u + 12345u);
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    const uint32_t e = ((mix(s) % (n_entries / 2)) * 2) + xb;  // the two lanes of a pair hit adjacent entries
    if (MODE == 0) {
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gf + 2 * (size_t)e), "f"(1.f), "f"(2.f) : "memory");
    } else if (MODE == 1) {
      const __half2 v = __floats2half2_rn(1.f, 2.f);
      asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(gh + e), "r"(*reinterpret_cast<const uint32_t*>(&v)) : "memory");
    } else if (MODE == 2) {
      const __half2 h = __ldg(gh + e);
      acc += __low2float(h);
    } else if (MODE == 3) {
      float a = 1.f, b = 2.f;
      const float c = __shfl_down_sync(0xffffffffu, a, 1), d = __shfl_down_sync(0xffffffffu, b, 1);
      if (xb == 0) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gf + 2 * (size_t)e), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
    } else {
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(gf + 2 * (size_t)e), "f"(1.f) : "memory");
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(gf + 2 * (size_t)e + 1), "f"(2.f) : "memory");
    }
  }
  if (acc == 123.456f) *sink = acc;
}

template <int MODE>
void run(const char* name, float* gf, __half2* gh, uint32_t n_entries, float* sink) {
  const int blocks = 148 * 4, threads = 512, iters = 256;  // 2 * 148 * 4 * 256 * 256 lanes-ops
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(gf, gh, n_entries, iters, sink);
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(gf, gh, n_entries, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double ops = (double)blocks * threads * iters;
  printf("%-28s entries %9u: %8.3f ms  %7.2f G lane-ops/s  (134.2M lane-ops would take %.3f ms)\n", name, n_entries, best, ops / best * 1e-6,
         134.2e6 / (ops / best));
}

int main() {
  const uint32_t n_max = 5124512;
  float* gf;
  __half2* gh;
  float* sink;
  cudaMalloc(&gf, (size_t)n_max * 8);
  cudaMalloc(&gh, (size_t)n_max * 4);
  cudaMalloc(&sink, 4);
  cudaMemset(gf, 0, (size_t)n_max * 8);
  cudaMemset(gh, 0, (size_t)n_max * 4);
  for (uint32_t n : {5124512u, 524288u, 65536u}) {
    run<0>("red.v2.f32 (pair=16B)", gf, gh, n, sink);
    run<1>("red.f16x2 (pair=8B)", gf, gh, n, sink);
    run<3>("red.v4.f32 even lanes", gf, gh, n, sink);
    run<4>("2x red.f32", gf, gh, n, sink);
    run<2>("ldg half2 (pair=8B)", gf, gh, n, sink);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
