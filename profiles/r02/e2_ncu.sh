#!/bin/bash
set -x
O=gpurun_out/r02e; mkdir -p $O
for wv in 0 1; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_march2 -s 6 -c 1 -o /tmp/nve_wrap$wv \
    python bench.py --ensemble nve --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --wrap $wv --tail 0 > $O/ncu_nve_wrap$wv.log 2>&1; tail -2 $O/ncu_nve_wrap$wv.log | cut -c1-200
ncu -i /tmp/nve_wrap$wv.ncu-rep --page raw --csv > $O/nve_wrap${wv}_raw.csv 2>/dev/null
ncu -i /tmp/nve_wrap$wv.ncu-rep --page source --csv > $O/nve_wrap${wv}_source.csv 2>/dev/null
done
ls -la $O
