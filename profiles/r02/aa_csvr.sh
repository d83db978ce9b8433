#!/bin/bash
set -x
O=gpurun_out/r02aa; mkdir -p $O
python -m pytest tests/test_langevin_gpu.py tests/test_hooks_gpu.py tests/test_md_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; tail -25 $O/pytest.log
