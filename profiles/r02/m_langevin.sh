#!/bin/bash
set -x
O=gpurun_out/r02m; mkdir -p $O
python -m pytest tests/test_langevin_gpu.py tests/test_hooks_gpu.py tests/test_md_gpu.py -m gpu -x -q > $O/pytest_langevin.log 2>&1; tail -25 $O/pytest_langevin.log
python - > $O/langevin_bench.log 2>&1 <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import bench
from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
from micmec_b200.sampling.verlet import VerletIntegrator
from micmec_b200.sampling.nvt import LangevinThermostat
from micmec_b200.units import femtosecond
import torch
for grid in (64, 256):
    system, vel0 = bench.make_state(grid)
    part = ForcePartMechanical(system, device=0)
    mmf = MicMecForceField(system, [part])
    np.random.seed(0)
    verlet = VerletIntegrator(mmf, timestep=10*femtosecond, hooks=[LangevinThermostat(300.0, timecon=100*femtosecond)], vel0=vel0)
    assert verlet.device_mode
    verlet.run(10); torch.cuda.synchronize()
    t0 = time.perf_counter(); verlet.run(100); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("langevin NVT %d^3: %.4f ms/step (incl. final state read-back), %.3e node-steps/s, T=%.2f" % (grid, 10*dt, grid**3*100/dt, verlet.temp), flush=True)
    del verlet, mmf, part
PY
cat $O/langevin_bench.log | tail -4
