#!/bin/bash
set -x
O=gpurun_out/r02z; mkdir -p $O
for g in 64 96 128 160 192; do for ens in nvt npt; do for wv in 0 1; do
  python bench.py --grid $g --ensemble $ens --wrap $wv --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > $O/b_${g}_${ens}_$wv.json 2>/dev/null
  python -c "import json; d=json.load(open('$O/b_${g}_${ens}_$wv.json')); print('grid $g $ens wrap $wv ms/step %.4f  %.3e' % (d['ms_per_step'], d['value']))"
done; done; done
