#!/bin/bash
# round 2, call AG: ghost fill beside the scalar kernel (side stream) A/B on one GPU, then the full GPU suite and smoke() on it
set -x
O=gpurun_out/r02ag; mkdir -p $O
for ens in npt nve nvt; do for ov in 1 0; do
  MICMEC_B200_HALO_OVERLAP=$ov timeout 120 python bench.py --ensemble $ens --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_${ens}_ov$ov.json 2>$O/bench_${ens}_ov$ov.err
  python -c "import json; d=json.load(open('$O/bench_${ens}_ov$ov.json')); r=d['roofline']; print('OV $ens overlap $ov ms/step %.4f %.4e kern %.4f launches %d epot %.13e econs %.13e' % (d['ms_per_step'], d['value'], r['kernel_ms'], d['gpu_launches'], d['check']['epot'], d['check']['econs']))"
done; done
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
