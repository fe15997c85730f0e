set -x
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/dp_train_check.py 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/r1g_dp_train_check.log
