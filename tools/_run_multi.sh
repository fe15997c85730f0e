# scratch driver for one multi-GPU gpurun session (round 2): DP tests, weak / strong scaling lines, config 5 on 4 GPUs
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -rfs --tb=short 2>&1 | tail -40) > gpurun_out/r02_pytest_dp.log
(timeout 300 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 200 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_bench_8gpu_weak.json
(NSV_DP_SYNC=host timeout 300 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --steps 200 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_bench_8gpu_weak_hostsync.json
(timeout 300 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --steps 200 --warmup 5 --scaling strong 2>&1 | tail -1) > gpurun_out/r02_bench_8gpu_strong.json
(timeout 600 $TR --nproc-per-node 4 --master-port 29614 tools/cfg5_workload.py --iters 5000 2>&1 | tail -1) > gpurun_out/r02_cfg5_4gpu.json
(timeout 300 $TR --nproc-per-node 2 --master-port 29615 bench.py --gpus 2 --steps 200 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_bench_2gpu_weak.json
(timeout 300 $TR --nproc-per-node 4 --master-port 29616 bench.py --gpus 4 --steps 200 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_bench_4gpu_weak.json
(timeout 300 $TR --nproc-per-node 4 --master-port 29617 bench.py --gpus 4 --steps 200 --warmup 5 --scaling strong 2>&1 | tail -1) > gpurun_out/r02_bench_4gpu_strong.json
(timeout 300 $TR --nproc-per-node 2 --master-port 29618 bench.py --gpus 2 --steps 200 --warmup 5 --scaling strong 2>&1 | tail -1) > gpurun_out/r02_bench_2gpu_strong.json
tail -6 gpurun_out/r02_pytest_dp.log
for f in gpurun_out/r02_bench_*gpu_*.json; do echo $f; cut -c1-260 $f; echo; done
