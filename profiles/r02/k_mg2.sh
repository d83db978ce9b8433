#!/bin/bash
set -x
O=gpurun_out/r02k; mkdir -p $O
python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -s > $O/pytest_multigpu.log 2>&1; tail -25 $O/pytest_multigpu.log
timeout 600 compute-sanitizer --tool memcheck python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/r02/sanitize_run.py > $O/memcheck_2gpu.log 2>&1; tail -8 $O/memcheck_2gpu.log
