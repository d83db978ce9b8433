#!/bin/bash
# round 2, call Y: launch list of the slab-default path (images on load + one tail launch per marching launch) on one GPU,
# and ncu --set full of the default-model (eight-corner) kernels
set -x
O=gpurun_out/r02y; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 36 --csv --log-file $O/launches_npt_images_tail.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --wrap 1 > $O/launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_march2 -s 10 -c 3 -o /tmp/npt_defmodel \
    python bench.py --model default --ensemble npt --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_defmodel.log 2>&1; tail -2 $O/ncu_defmodel.log | cut -c1-200
ncu -i /tmp/npt_defmodel.ncu-rep --page raw --csv > $O/npt_default_model_raw.csv 2>/dev/null
ls -la $O
