// slice_acq.cu -- kernel B, bit-exact flavour (namespace nsv::sa_exact): the generic implementation of
// slice_acq_impl.cuh compiled with -fmad=false so that the gather passes reproduce the reference's C arithmetic
// (slice_acq_cuda_kernel.cu:18-950 as built for the CPU by oracle/build_ref.sh) bit for bit.  Selected at run time with
// nsv_set_slice_acq_exact(1) / NSV_SLICE_ACQ_EXACT=1; the C ABI and the fast product kernels live in slice_acq_fast.cu.
#define NSV_SA_NS sa_exact
#include "slice_acq_impl.cuh"
