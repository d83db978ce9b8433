#!/bin/bash
set -x
O=gpurun_out/r02q; mkdir -p $O
python -m pytest tests/test_md_gpu.py -m gpu -x -q > $O/pytest_md.log 2>&1; tail -5 $O/pytest_md.log
python - > $O/stream_bench.log 2>&1 <<'PY'
import sys, time, shutil, os
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
from micmec_b200.sampling.verlet import VerletIntegrator
from micmec_b200.sampling.trajectory import DeviceRawWriter, RawWriter, load_raw
from micmec_b200.units import femtosecond
system, vel0 = bench.make_state(256)
for cls, kw in ((None, {}), (DeviceRawWriter, dict(fields=("pos",))), (RawWriter, dict(keys=("pos",)))):
    d = "/tmp/traj_%s" % (cls.__name__ if cls else "none")
    shutil.rmtree(d, ignore_errors=True)
    part = ForcePartMechanical(system, device=0)
    mmf = MicMecForceField(system, [part])
    hooks = [cls(d, start=5, step=5, **kw)] if cls else []
    verlet = VerletIntegrator(mmf, timestep=10 * femtosecond, hooks=hooks, vel0=vel0)
    verlet.run(5); torch.cuda.synchronize()
    t0 = time.perf_counter(); verlet.run(40); torch.cuda.synchronize(); t1 = time.perf_counter()
    if cls is DeviceRawWriter: hooks[0].close()
    t2 = time.perf_counter()
    frames = load_raw(d)["pos"].shape[0] if cls else 0
    print("%-16s 40 NVE steps at 256^3 with a 403 MB frame every 5 steps: %.3f s in run() (+ %.3f s until the last frame is on disk), %d frames" % (
        cls.__name__ if cls else "no writer", t1 - t0, t2 - t1, frames), flush=True)
    del verlet, mmf, part
    shutil.rmtree(d, ignore_errors=True)
PY
tail -4 $O/stream_bench.log
