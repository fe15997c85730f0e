# scratch driver, 8-GPU session 2 (round 2): the three synchronisation schemes of the data-parallel optimiser, weak + strong
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 600 $TR --nproc-per-node 2 --master-port 29631 tools/dp_check.py 2>&1 | tail -3) > gpurun_out/r02_dp_check_2gpu.log
for mode in hybrid host; do
(NSV_DP_SYNC=$mode timeout 300 $TR --nproc-per-node 8 --master-port 29632 bench.py --gpus 8 --steps 300 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_bench_8gpu_weak_$mode.json
(NSV_DP_SYNC=$mode timeout 300 $TR --nproc-per-node 8 --master-port 29633 bench.py --gpus 8 --steps 300 --warmup 5 --scaling strong 2>&1 | tail -1) > gpurun_out/r02_bench_8gpu_strong_$mode.json
done
(NSV_DP_SYNC=hybrid timeout 300 $TR --nproc-per-node 4 --master-port 29634 bench.py --gpus 4 --steps 300 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_bench_4gpu_weak_hybrid.json
(NSV_DP_SYNC=hybrid timeout 300 $TR --nproc-per-node 2 --master-port 29635 bench.py --gpus 2 --steps 300 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_bench_2gpu_weak_hybrid.json
cat gpurun_out/r02_dp_check_2gpu.log | cut -c1-400
for f in gpurun_out/r02_bench_*gpu_*_h*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], d['dp']['exchange_ms'])"; done
