set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -q -s -k "full_size" 2>&1 | grep -v "^$" | tail -30 | tee gpurun_out/r1d_fullsize.log
