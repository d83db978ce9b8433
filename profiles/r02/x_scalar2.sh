#!/bin/bash
set -x
O=gpurun_out/r02x; mkdir -p $O
python -m pytest tests/test_md_gpu.py tests/test_large_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file $O/launches_npt.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > $O/launches_npt.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02x/launches_npt.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value')
print(' '.join('%s:%.0f' % (r[kn].split('(')[0].replace('void ','')[:10], float(r[mv].replace(',',''))/1000) for r in rows[start:] if len(r)>mv and 'scalar' in r[kn]))
PY
for i in 1 2; do python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > $O/bench_$i.json 2>/dev/null; python -c "import json; d=json.load(open('$O/bench_$i.json')); print('ms/step %.4f' % d['ms_per_step'])"; done
