#!/bin/bash
# round 2, call AH: repeat of the ghost-fill overlap A/B (NPT, alternating) - call AG's box gave erratic whole-step times
O=gpurun_out/r02ah; mkdir -p $O
for rep in 1 2; do for ov in 1 0; do
  MICMEC_B200_HALO_OVERLAP=$ov timeout 60 python bench.py --ensemble npt --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_npt_ov${ov}_$rep.json 2>/dev/null
  python -c "import json; d=json.load(open('$O/bench_npt_ov${ov}_$rep.json')); print('OV npt overlap $ov rep $rep ms/step %.4f kern %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
done; done
