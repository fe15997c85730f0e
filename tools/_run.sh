set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r1_pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 2>&1 | tail -3 | tee gpurun_out/r1_bench.log
