set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 200 python -m pytest tests/test_gpu_fused.py -q -s -k "bias or rejects" 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/r1b_bias_tests.log
timeout 120 python -m pytest "tests/test_gpu_e2e.py::test_train_with_fused_bias_field_head" -q -s 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/r1b_bias_e2e.log
timeout 100 python tools/bench_kernel_a.py --cfg 5 --variants 1048576:1 --reps 10 2>&1 | tail -2 | tee gpurun_out/r1b_cfg5_kernel.log
timeout 100 python tools/bench_kernel_a.py --cfg 3 --variants 1048576:1 --reps 10 2>&1 | tail -2 | tee gpurun_out/r1b_cfg3_kernel.log
timeout 60 python tools/run_kernel_b.py 2>&1 | tail -2 | tee gpurun_out/r1b_kernelB_times.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"forward_kernel|backward_kernel" -c 5 -f -o gpurun_out/r1b_kernelB python tools/run_kernel_b.py --reps 0 2>&1 | tail -3
timeout 200 python -m pytest tests/test_gpu_fused.py -q -k "not bias and not rejects" 2>&1 | tail -5 | tee gpurun_out/r1b_fused_rest.log
ls -la gpurun_out
