set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_pose.py tests/test_gpu_e2e.py -q 2>&1 | tail -4 | tee gpurun_out/r1i_tests.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r1i_smoke.log
timeout 300 python bench.py --steps 300 --warmup 5 2>&1 | tail -1 | tee gpurun_out/r1i_bench.log
