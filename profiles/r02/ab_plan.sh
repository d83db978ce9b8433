#!/bin/bash
set -x
O=gpurun_out/r02ab; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
for ens in npt nve; do
  python bench.py --ensemble $ens --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_$ens.json 2> $O/bench_$ens.err
  python -c "import json; d=json.load(open('$O/bench_$ens.json')); r=d['roofline']; print('$ens ms/step %.4f %.3e kern %.4f force %.4f' % (d['ms_per_step'], d['value'], r['kernel_ms'], (r.get('force_only_kernel') or {}).get('kernel_ms', 0)), d['config']['kernel_tiling'], d['check']['epot'])"
done
python bench.py --ensemble npt --chunk 43 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_npt_uniform43.json 2>/dev/null
python -c "import json; d=json.load(open('$O/bench_npt_uniform43.json')); print('npt uniform chunk 43: ms/step %.4f' % d['ms_per_step'], d['config']['kernel_tiling'])"
