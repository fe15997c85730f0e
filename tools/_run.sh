# scratch driver for gpurun sessions: full GPU suite, smoke, bench (what the round-end run does)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.log
