#!/bin/bash
set -x
O=gpurun_out/r02ad; mkdir -p $O
for ens in npt nve; do for pl in 1 0; do
  python bench.py --ensemble $ens --plan $pl --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_${ens}_$pl.json 2>/dev/null
  python -c "import json; d=json.load(open('$O/bench_${ens}_$pl.json')); r=d['roofline']; t=d['config']['kernel_tiling']; print('N=1 $ens plan $pl ms/step %.4f %.3e kern %.4f force %.4f blocks %d eff %.3f' % (d['ms_per_step'], d['value'], r['kernel_ms'], (r.get('force_only_kernel') or {}).get('kernel_ms', 0), t['blocks'], t['schedule_efficiency']))"
done; done
