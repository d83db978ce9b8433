#!/bin/bash
set -x
O=gpurun_out/r02z2; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
python bench.py --ensemble nvt --grid 64 --no-cpu-baseline > $O/bench_nvt_64.json 2>/dev/null; python -c "import json; d=json.load(open('$O/bench_nvt_64.json')); print('64^3 nvt %.4f ms %.3e' % (d['ms_per_step'], d['value']), d['config']['kernel_tiling'])"
python bench.py --no-cpu-baseline --no-e2e --steps 100 --warmup 10 > $O/bench_npt_256.json 2>/dev/null; python -c "import json; d=json.load(open('$O/bench_npt_256.json')); print('256^3 npt %.4f ms %.3e' % (d['ms_per_step'], d['value']), d['config']['kernel_tiling'])"
