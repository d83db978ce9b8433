#!/bin/bash
# Round-1 (third session) evidence run on one B200: GPU parity tests, bench lines, ncu launch list, one full capture.
set -u
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_npt.json 2> gpurun_out/bench_npt.err
timeout 200 python bench.py --ensemble nve --no-cpu-baseline > gpurun_out/bench_nve.json 2> gpurun_out/bench_nve.err
timeout 200 python bench.py --ensemble nvt --no-cpu-baseline --no-e2e > gpurun_out/bench_nvt.json 2> gpurun_out/bench_nvt.err
timeout 200 python bench.py --grid 64 --ensemble nvt --no-cpu-baseline > gpurun_out/bench_nvt64.json 2> gpurun_out/bench_nvt64.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_npt.csv \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_march -s 6 -c 1 -f -o gpurun_out/prof_march_v14_nve \
    python bench.py --ensemble nve --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
for f in gpurun_out/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(' ', '%.3e'%d['value'], 'ms/step', d['ms_per_step'], 'frac', (d.get('roofline') or {}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), 'launches', d.get('gpu_launches'))
" || tail -3 ${f%.json}.err; done
ls -la gpurun_out
