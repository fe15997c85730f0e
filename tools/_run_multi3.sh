# scratch driver, 8-GPU session 3 (round 2): hybrid (optimised in-kernel prologue) vs host synchronisation, interleaved, + the scaling series
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 600 $TR --nproc-per-node 2 --master-port 29641 tools/dp_check.py 2>&1 | tail -2) > gpurun_out/r02_dp_check_2gpu.log
for rep in a b; do for mode in hybrid host; do
(NSV_DP_SYNC=$mode timeout 300 $TR --nproc-per-node 8 --master-port 29642 bench.py --gpus 8 --steps 300 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_s3_8gpu_weak_${mode}_$rep.json
done; done
for mode in hybrid host; do
(NSV_DP_SYNC=$mode timeout 300 $TR --nproc-per-node 8 --master-port 29643 bench.py --gpus 8 --steps 300 --warmup 5 --scaling strong 2>&1 | tail -1) > gpurun_out/r02_s3_8gpu_strong_$mode.json
(NSV_DP_SYNC=$mode timeout 300 $TR --nproc-per-node 4 --master-port 29644 bench.py --gpus 4 --steps 300 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_s3_4gpu_weak_$mode.json
(NSV_DP_SYNC=$mode timeout 300 $TR --nproc-per-node 2 --master-port 29645 bench.py --gpus 2 --steps 300 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_s3_2gpu_weak_$mode.json
done
