#!/bin/bash
# round 2, call AJ: bench.py with the graph capture moved out of the timed region (warm-up split into two calls)
O=gpurun_out/r02aj; mkdir -p $O
for run in npt_1 npt_2 nve_1; do
  timeout 25 python bench.py --ensemble ${run%_*} --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_$run.json 2>/dev/null
  python -c "import json; d=json.load(open('$O/bench_$run.json')); print('FINAL $run ms/step %.4f value %.4e kern %.4f clocks %s epot %.13e steps_total %d' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['clocks'], d['check']['epot'], d['check']['steps_total']))"
done
