#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds the reference's OWN CUDA extensions for sm_100a -- slice acquisition and the pose converters -- straight from
# where their source files lie under /root/reference, into oracle/_ref/nesvor_ref_slice_acq_cuda.so and
# oracle/_ref/nesvor_ref_transform_convert_cuda.so (git-ignored, travel with gpurun): the GPU-side cross-check / baseline
# for kernel B and the converters.  It does not run the reference's build system
# (setup.py / torch JIT): g++ for the *_cuda.cpp bindings, nvcc for the *_cuda_kernel.cu files.  The .cu files do not compile
# against torch 2.11 as shipped (10 x AT_DISPATCH_FLOATING_TYPES(x.type(), ...)); the one-token fix
# (.type() -> .scalar_type()) is applied with sed into a scratch file under $TMPDIR that is deleted afterwards --
# nothing from /root/reference is copied into the repository.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${NSV_REFERENCE_ROOT:-/root/reference}"
PY="${NSV_PYTHON:-python}"
NVCC="${NSV_NVCC:-/usr/local/cuda/bin/nvcc}"
CXX="${NSV_CXX:-/usr/bin/g++}"
if [[ ! -f "$REF/nesvor/slice_acquisition/slice_acq_cuda.cpp" || ! -f "$REF/nesvor/transform/transform_convert_cuda.cpp" ]]; then
  echo "build_ref_gpu.sh: reference sources not found under $REF -- skipping (prebuilt files under $HERE/_ref are used if present)" >&2
  exit 0
fi
mkdir -p "$HERE/_ref"
SCRATCH="$(mktemp -d)"
trap 'rm -rf "$SCRATCH"' EXIT
read -r INCS LIBDIR <<<"$($PY - <<'PYEOF'
import os, sysconfig, torch
from torch.utils.cpp_extension import include_paths
incs = include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
print(",".join(incs), os.path.join(os.path.dirname(torch.__file__), "lib"))
PYEOF
)"
IFLAGS=""
IFS=',' read -ra ARR <<<"$INCS"
for i in "${ARR[@]}"; do IFLAGS="$IFLAGS -isystem $i"; done

# build_one <module name> <binding .cpp> <kernel .cu>
build_one() {
  local NAME="$1" CPP="$2" CU="$3" OUT="$HERE/_ref/$1.so"
  local DEFS="-DTORCH_EXTENSION_NAME=$NAME -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1"
  sed -E 's/AT_DISPATCH_FLOATING_TYPES\(([A-Za-z_]+)\.type\(\)/AT_DISPATCH_FLOATING_TYPES(\1.scalar_type()/' "$CU" > "$SCRATCH/$NAME.cu"
  $CXX -O2 -fPIC -std=c++17 -w $DEFS $IFLAGS -c "$CPP" -o "$SCRATCH/$NAME.binding.o" &
  $NVCC -O3 -std=c++17 -w -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $DEFS $IFLAGS \
    -c "$SCRATCH/$NAME.cu" -o "$SCRATCH/$NAME.kernel.o"
  wait
  $CXX -shared -o "$OUT" "$SCRATCH/$NAME.binding.o" "$SCRATCH/$NAME.kernel.o" -L"$LIBDIR" -Wl,-rpath,"$LIBDIR" -lc10 -ltorch_cpu -ltorch \
    -ltorch_python -lc10_cuda -ltorch_cuda -L/usr/local/cuda/lib64 -lcudart
  echo "built $OUT"
}
WHAT="${1:-all}"
PIDS=()
if [[ "$WHAT" == "all" || "$WHAT" == "slice_acq" ]]; then
  ( build_one nesvor_ref_slice_acq_cuda "$REF/nesvor/slice_acquisition/slice_acq_cuda.cpp" "$REF/nesvor/slice_acquisition/slice_acq_cuda_kernel.cu" ) &
  PIDS+=($!)
fi
if [[ "$WHAT" == "all" || "$WHAT" == "transform" ]]; then
  ( build_one nesvor_ref_transform_convert_cuda "$REF/nesvor/transform/transform_convert_cuda.cpp" "$REF/nesvor/transform/transform_convert_cuda_kernel.cu" ) &
  PIDS+=($!)
fi
for p in "${PIDS[@]}"; do wait "$p"; done   # the two extensions build side by side; any failure fails the script (set -e)
