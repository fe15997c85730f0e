set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py 2>&1 | grep -v "^$" | tail -6 | tee gpurun_out/r1f_dp_check.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/dp_bias_check.py 2>&1 | grep -v "^$" | tail -6 | tee gpurun_out/r1f_dp_bias_check.log
