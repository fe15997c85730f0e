// Micro-benchmark: throughput of random gathers / reductions with kernel A's access patterns
// (L2-resident working set, 148 x 4 CTAs x 512 threads).  Every mode reports time per "entry-op"
// (one table entry = F=2 features read or accumulated) so that modes are comparable: kernel A does
// 2^20 x 128 = 134.2 M entry-ops per direction per iteration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/red_bench tools/red_bench.cu && tools/red_bench
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

enum Mode {
  RED_V2F32_PAIR = 0,   // every lane: red.v2.f32 on its entry; lanes of a pair adjacent (today's kernel)
  RED_F16X2_PAIR,       // every lane: red.f16x2 (4 B)
  RED_V4F32_EVEN,       // even lanes: red.v4.f32 covering the pair (16 B)
  RED_2XF32,            // every lane: two scalar red.f32
  RED_V2F16X2_EVEN,     // even lanes: red.v2.f16x2 covering the pair (8 B)
  RED_V2F32_RANDOM,     // every lane its own random entry (no pairing): 32 lines / instruction
  RED_V2F32_QUAD,       // 4 lanes adjacent (32 B sector), 8 sectors / instruction
  RED_V4F32_ALL,        // every lane: red.v4.f32 on its own random aligned pair (32 lines / instr, 2 entries / lane)
  LDG_U32_PAIR,         // every lane: 4-byte load, pair adjacent (today's kernel)
  LDG_U64_EVEN,         // even lanes: 8-byte load covering the pair
  LDG_U32_RANDOM,       // every lane its own random entry
  LDG_U64_ALL,          // every lane: 8-byte load of a random aligned pair (2 entries / lane)
  LDG_U128_ALL,         // every lane: 16-byte load of a random aligned quad (4 entries / lane)
  ATOMS_F32_SPREAD,     // shared-memory red.shared.add.f32 at random words of a 32 KB region
  MIX_LDG_RED_PAIR,     // LDG_U32_PAIR + RED_V2F32_PAIR interleaved (both counted)
  N_MODES
};

template <int MODE>
__global__ void __launch_bounds__(512) k(float* gf, __half2* gh, uint32_t n_entries, int iters, float* sink) {
  __shared__ float sm[MODE == ATOMS_F32_SPREAD ? 8192 : 1];
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t pair = tid >> 1, xb = tid & 1;
  float acc = 0.f;
  uint32_t s = mix(pair * 2654435761u + 12345u);
  uint32_t s1 = mix(tid * 2654435761u + 777u);
  if (MODE == ATOMS_F32_SPREAD) {
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
  }
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    s1 = s1 * 1664525u + 1013904223u;
    const uint32_t ep = ((mix(s) % (n_entries / 2)) * 2);  // pair-shared aligned pair
    const uint32_t e = ep + xb;
    const uint32_t er = mix(s1) % n_entries;               // lane-private random entry
    const uint32_t er2 = er & ~1u, er4 = er & ~3u;
    if (MODE == RED_V2F32_PAIR) {
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gf + 2 * (size_t)e), "f"(1.f), "f"(2.f) : "memory");
    } else if (MODE == RED_F16X2_PAIR) {
      const __half2 v = __floats2half2_rn(1.f, 2.f);
      asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(gh + e), "r"(*reinterpret_cast<const uint32_t*>(&v)) : "memory");
    } else if (MODE == RED_V4F32_EVEN) {
      float a = 1.f + i, b = 2.f;
      const float c = __shfl_down_sync(0xffffffffu, a, 1), d = __shfl_down_sync(0xffffffffu, b, 1);
      if (xb == 0) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gf + 2 * (size_t)ep), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
    } else if (MODE == RED_2XF32) {
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(gf + 2 * (size_t)e), "f"(1.f) : "memory");
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(gf + 2 * (size_t)e + 1), "f"(2.f) : "memory");
    } else if (MODE == RED_V2F16X2_EVEN) {
      const __half2 v = __floats2half2_rn(1.f, 2.f);
      const uint32_t u = *reinterpret_cast<const uint32_t*>(&v);
      if (xb == 0) asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(gh + ep), "r"(u), "r"(u) : "memory");
    } else if (MODE == RED_V2F32_RANDOM) {
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gf + 2 * (size_t)er), "f"(1.f), "f"(2.f) : "memory");
    } else if (MODE == RED_V2F32_QUAD) {
      uint32_t sq = mix((tid >> 2) * 2654435761u + 99u + (uint32_t)i * 0x9E3779B9u);
      const uint32_t e4 = ((sq % (n_entries / 4)) * 4) + (tid & 3);
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gf + 2 * (size_t)e4), "f"(1.f), "f"(2.f) : "memory");
    } else if (MODE == RED_V4F32_ALL) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gf + 2 * (size_t)er2), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
    } else if (MODE == LDG_U32_PAIR) {
      const __half2 h = __ldg(gh + e);
      acc += __low2float(h);
    } else if (MODE == LDG_U64_EVEN) {
      if (xb == 0) {
        const uint2 h = __ldg(reinterpret_cast<const uint2*>(gh + ep));
        acc += __uint_as_float(h.x ^ h.y);
      }
    } else if (MODE == LDG_U32_RANDOM) {
      const __half2 h = __ldg(gh + er);
      acc += __low2float(h);
    } else if (MODE == LDG_U64_ALL) {
      const uint2 h = __ldg(reinterpret_cast<const uint2*>(gh + er2));
      acc += __uint_as_float(h.x ^ h.y);
    } else if (MODE == LDG_U128_ALL) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(gh + er4));
      acc += __uint_as_float(h.x ^ h.y ^ h.z ^ h.w);
    } else if (MODE == ATOMS_F32_SPREAD) {
      asm volatile("red.shared.add.f32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(sm + (er & 8191u))), "f"(1.f) : "memory");
    } else if (MODE == MIX_LDG_RED_PAIR) {
      const __half2 h = __ldg(gh + e);
      acc += __low2float(h);
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gf + 2 * (size_t)(n_entries - 1 - e)), "f"(1.f), "f"(2.f) : "memory");
    }
  }
  if (MODE == ATOMS_F32_SPREAD) {
    __syncthreads();
    acc += sm[threadIdx.x];
  }
  if (acc == 123.456f) *sink = acc;
}

template <int MODE>
void run(const char* name, double entries_per_lane_op, float* gf, __half2* gh, uint32_t n_entries, float* sink) {
  const int blocks = 148 * 4, threads = 512, iters = 256;
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
  const double entry_ops = (double)blocks * threads * iters * entries_per_lane_op;
  printf("%-22s entries %8u: %7.3f ms  %7.2f G entry-ops/s  (134.2M entry-ops: %.3f ms)\n", name, n_entries, best, entry_ops / best * 1e-6,
         134.2e6 / (entry_ops / best));
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
    run<RED_V2F32_PAIR>("red.v2.f32 pair", 1, gf, gh, n, sink);
    run<RED_F16X2_PAIR>("red.f16x2 pair", 1, gf, gh, n, sink);
    run<RED_V4F32_EVEN>("red.v4.f32 even", 1, gf, gh, n, sink);
    run<RED_2XF32>("2x red.f32 pair", 1, gf, gh, n, sink);
    run<RED_V2F16X2_EVEN>("red.v2.f16x2 even", 1, gf, gh, n, sink);
    run<RED_V2F32_RANDOM>("red.v2.f32 random", 1, gf, gh, n, sink);
    run<RED_V2F32_QUAD>("red.v2.f32 quad", 1, gf, gh, n, sink);
    run<RED_V4F32_ALL>("red.v4.f32 all", 2, gf, gh, n, sink);
    run<LDG_U32_PAIR>("ldg.32 pair", 1, gf, gh, n, sink);
    run<LDG_U64_EVEN>("ldg.64 even", 1, gf, gh, n, sink);
    run<LDG_U32_RANDOM>("ldg.32 random", 1, gf, gh, n, sink);
    run<LDG_U64_ALL>("ldg.64 all", 2, gf, gh, n, sink);
    run<LDG_U128_ALL>("ldg.128 all", 4, gf, gh, n, sink);
    run<ATOMS_F32_SPREAD>("red.shared.f32 spread", 0.5, gf, gh, n, sink);
    run<MIX_LDG_RED_PAIR>("ldg.32+red.v2 pair", 2, gf, gh, n, sink);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
