"""Multi-GPU check of the fused data-parallel optimiser (nsv_adamw_step_dp: reduce-scatter + AdamW + all-gather over
NVLink peer memory) against NCCL all-reduce + nsv_adamw_step.  Needs >= 2 GPUs on the node; skipped otherwise (the
host-side sharding logic is covered on CPU by tests/test_distributed_cpu.py with gloo)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_memory_optimizer_matches_allreduce(native_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0


def test_bias_field_head_data_parallel(native_lib):
    """Config-5 heads over 2 ranks: the global mean(log_bias) (biasReg) and the rank-averaged gradient must equal one
    launch over the concatenated batch (tools/dp_bias_check.py)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tools", "dp_bias_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0


def test_train_is_data_parallel_under_a_process_group(native_lib):
    """`train(slices, args)` inside a 2-rank NCCL group (tools/dp_train_check.py): replicas identical, phantom reconstructed."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29535", os.path.join(ROOT, "tools", "dp_train_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0
