#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds the reference's OWN slice-acquisition CUDA extension for sm_100a, straight from where its two source files lie
# under /root/reference, into oracle/_ref/nesvor_ref_slice_acq_cuda.so (git-ignored, travels with gpurun) -- the GPU-side
# cross-check / baseline for kernel B.  It does not run the reference's build system
# (setup.py / torch JIT): g++ for slice_acq_cuda.cpp, nvcc for slice_acq_cuda_kernel.cu.  The .cu file does not compile
# against torch 2.11 as shipped (6 x AT_DISPATCH_FLOATING_TYPES(x.type(), ...)); the one-token fix
# (.type() -> .scalar_type()) is applied with sed into a scratch file under $TMPDIR that is deleted afterwards --
# nothing from /root/reference is copied into the repository.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${NSV_REFERENCE_ROOT:-/root/reference}"
CPP="$REF/nesvor/slice_acquisition/slice_acq_cuda.cpp"
CU="$REF/nesvor/slice_acquisition/slice_acq_cuda_kernel.cu"
OUT="$HERE/_ref/nesvor_ref_slice_acq_cuda.so"
if [[ ! -f "$CPP" || ! -f "$CU" ]]; then
  echo "build_ref_gpu.sh: reference sources not found under $REF -- skipping (a prebuilt $OUT is used if present)" >&2
  exit 0
fi
PY="${NSV_PYTHON:-python}"
NVCC="${NSV_NVCC:-/usr/local/cuda/bin/nvcc}"
CXX="${NSV_CXX:-/usr/bin/g++}"
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
DEFS="-DTORCH_EXTENSION_NAME=nesvor_ref_slice_acq_cuda -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1"
sed -E 's/AT_DISPATCH_FLOATING_TYPES\(([A-Za-z_]+)\.type\(\)/AT_DISPATCH_FLOATING_TYPES(\1.scalar_type()/' "$CU" > "$SCRATCH/kernel.cu"
$CXX -O2 -fPIC -std=c++17 -w $DEFS $IFLAGS -c "$CPP" -o "$SCRATCH/binding.o" &
$NVCC -O3 -std=c++17 -w -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $DEFS $IFLAGS \
  -c "$SCRATCH/kernel.cu" -o "$SCRATCH/kernel.o"
wait
$CXX -shared -o "$OUT" "$SCRATCH/binding.o" "$SCRATCH/kernel.o" -L"$LIBDIR" -Wl,-rpath,"$LIBDIR" -lc10 -ltorch_cpu -ltorch -ltorch_python \
  -lc10_cuda -ltorch_cuda -L/usr/local/cuda/lib64 -lcudart
echo "built $OUT"
