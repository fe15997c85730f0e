/*
 * oracle/nsv_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the parts of NeSVoR's hot path whose arithmetic lives in the
 * reference repository itself: the slice-acquisition operator family and the rigid-pose
 * converters.  It is the checker for the CUDA kernels in nesvor_b200/csrc and is itself pinned
 * against (a) oracle/_ref = the reference's own kernel bodies compiled for CPU
 * (oracle/build_ref.sh), (b) the reference's golden vectors (11 scipy axis-angle pairs,
 * tests/__init__.py:24-38) and (c) its known-answer test (CG recovery of the 32^3 phantom,
 * tests/slice_acquisition/test_slice_acq.py:76-81) -- see tests/test_oracle_*.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file.  The product (nesvor_b200) never does.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle/nsv_oracle.c -lm
 *        (done by oracle/build.py and __graft_entry__.build()).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define REAL float
#define SUFFIX f32
#include "slice_acq_oracle_impl.h"
#include "transform_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL double
#define SUFFIX f64
#include "slice_acq_oracle_impl.h"
#include "transform_oracle_impl.h"
#undef REAL
#undef SUFFIX
