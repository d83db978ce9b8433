#!/bin/bash
set -x
O=gpurun_out/r02u; mkdir -p $O
python -m pytest tests/test_hooks_gpu.py tests/test_md_gpu.py tests/test_langevin_gpu.py tests/test_dropin_gpu.py -m gpu -x -q > $O/pytest_hooks.log 2>&1; tail -25 $O/pytest_hooks.log
