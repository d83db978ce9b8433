#!/bin/bash
set -x
O=gpurun_out/r02o; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scalar -s 12 -c 4 -o /tmp/scalar \
    python bench.py --ensemble npt --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_scalar.log 2>&1; tail -2 $O/ncu_scalar.log | cut -c1-200
ncu -i /tmp/scalar.ncu-rep --page raw --csv > $O/scalar_raw.csv 2>/dev/null
ncu -i /tmp/scalar.ncu-rep --page source --csv > $O/scalar_source.csv 2>/dev/null
python bench.py --config5 3000 > $O/config5_3000.json 2> $O/config5_3000.err; cat $O/config5_3000.json; tail -3 $O/config5_3000.err
