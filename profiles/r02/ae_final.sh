#!/bin/bash
# round 2, call AE: final state with the two-class block schedule - full GPU suite, smoke(), default bench line, NVE / NVT lines,
# reference arm
set -x
O=gpurun_out/r02ae; mkdir -p $O
python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu.log 2>&1; tail -10 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
python bench.py > $O/bench_npt_256.json 2> $O/bench_npt_256.err; cut -c1-600 $O/bench_npt_256.json
for ens in nve nvt; do
  python bench.py --ensemble $ens --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_${ens}_256.json 2>/dev/null
  python -c "import json; d=json.load(open('$O/bench_${ens}_256.json')); r=d['roofline']; print('$ens ms/step %.4f %.3e kern %.4f frac %.3f fp64 %.3f' % (d['ms_per_step'], d['value'], r['kernel_ms'], r['frac'], r['fp64']['frac']))"
done
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-300 $O/bench_ref.json
