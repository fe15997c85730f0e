/*
 * TEST INFRASTRUCTURE ONLY -- never linked into or imported by the product path.
 *
 * Host shim that lets the reference's *own* CUDA kernel bodies
 * (/root/reference/nesvor/slice_acquisition/slice_acq_cuda_kernel.cu:8-952 and
 *  /root/reference/nesvor/transform/transform_convert_cuda_kernel.cu:8-442, i.e. the anonymous
 * namespace holding the __global__ templates) be compiled by g++ and executed on CPU cores.
 * oracle/build_ref.sh streams those line ranges from /root/reference through g++ together with
 * this header and a driver (.inc) -- no reference source is copied into the repository; only the
 * resulting shared object lands in the git-ignored oracle/_ref/.
 *
 * One "CUDA thread" is emulated per loop iteration: blockDim.x = 1, threadIdx.x = 0,
 * blockIdx.x = flat index.  atomicAdd becomes an OpenMP atomic so drivers may parallelise.
 */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>

struct nsv_ref_dim3 {
  unsigned x, y, z;
};
static thread_local nsv_ref_dim3 blockIdx = {0, 0, 0};
static thread_local nsv_ref_dim3 blockDim = {1, 1, 1};
static thread_local nsv_ref_dim3 threadIdx = {0, 0, 0};

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline

template <typename T>
static inline T atomicAdd(T* addr, T val) {
  T old;
#pragma omp atomic capture
  {
    old = *addr;
    *addr += val;
  }
  return old;
}

using std::floor;
using std::round;
