#!/bin/bash
# round 2, call AI: last check of the final build (ghost fill back in stream order, NVML initialised before the timed region)
O=gpurun_out/r02ai; mkdir -p $O
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
for rep in 1 2; do
  timeout 30 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_npt_$rep.json 2>/dev/null
  python -c "import json; d=json.load(open('$O/bench_npt_$rep.json')); print('FINAL npt rep $rep ms/step %.4f kern %.4f clocks %s epot %.13e' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'], d['check']['epot']))"
done
