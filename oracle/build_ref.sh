#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own kernel bodies for CPU, straight from where they lie under
# /root/reference, into oracle/_ref/libnesvor_ref_cpu.so (git-ignored, travels with gpurun).
# Nothing from /root/reference is copied into the repository: the kernel namespaces are streamed
# (awk) into g++'s stdin between oracle/ref_cpu_shim.h and the oracle/ref_driver_*.inc drivers.
#
# It does NOT run the reference's build system (setup.py / torch cpp_extension); the two .cu files
# need torch + a GPU as shipped, so the anonymous-namespace kernel templates are the only part used.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${NSV_REFERENCE_ROOT:-/root/reference}"
SA="$REF/nesvor/slice_acquisition/slice_acq_cuda_kernel.cu"
TC="$REF/nesvor/transform/transform_convert_cuda_kernel.cu"
if [[ ! -f "$SA" || ! -f "$TC" ]]; then
  echo "build_ref.sh: reference sources not found under $REF -- skipping (prebuilt oracle/_ref is used if present)" >&2
  exit 0
fi
mkdir -p "$HERE/_ref"
CXX="${NSV_CXX:-/usr/bin/g++}"   # not $CXX: this image presets it to a wrapper without libgomp
# -ffp-contract=off: no FMA contraction, so the result is the literal C arithmetic of the source.
FLAGS="-O2 -fPIC -fopenmp -ffp-contract=off -std=c++17 -w"
extract() { awk '/^namespace \{/{on=1} on{print} /^\} \/\/ namespace/{on=0}' "$1"; }

{ extract "$SA"; cat "$HERE/ref_driver_slice_acq.inc"; } |
  $CXX $FLAGS -x c++ -include "$HERE/ref_cpu_shim.h" -c -o "$HERE/_ref/ref_slice_acq.o" -
{ extract "$TC"; cat "$HERE/ref_driver_transform.inc"; } |
  $CXX $FLAGS -x c++ -include "$HERE/ref_cpu_shim.h" -c -o "$HERE/_ref/ref_transform.o" -
$CXX -shared -fopenmp -o "$HERE/_ref/libnesvor_ref_cpu.so" "$HERE/_ref/ref_slice_acq.o" "$HERE/_ref/ref_transform.o"
rm -f "$HERE/_ref/ref_slice_acq.o" "$HERE/_ref/ref_transform.o"
echo "built $HERE/_ref/libnesvor_ref_cpu.so"
