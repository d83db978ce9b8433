#!/bin/bash
set -x
O=gpurun_out/r02v; mkdir -p $O
python -m pytest tests/test_batchopt_gpu.py tests/test_force_gpu.py -m gpu -x -q > $O/pytest_qn.log 2>&1; tail -25 $O/pytest_qn.log
python - > $O/qn_bench.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from test_batchopt_cpu import replicas
from micmec_b200.replicas import ReplicaBatch
from micmec_b200.sampling.batchopt import ReplicaQNOptimizer, DeviceReplicaQNOptimizer
names = ["3x3x3_conf%d" % i for i in (0, 3, 9)]
for per in (128, 1024):
    systems = replicas(names, per, 0.5)
    pos0 = np.stack([s.pos for s in systems]); rvecs0 = np.stack([np.array(s.domain.rvecs) for s in systems])
    batch = ReplicaBatch(systems)
    t0 = time.perf_counter()
    dev = DeviceReplicaQNOptimizer(batch, pos0, rvecs0, gpos_rms=1e-7, dpos_rms=1e-5)
    sweeps = dev.run(1000)
    dt = time.perf_counter() - t0
    st = dev._fetch()
    print("device QN: %d replicas, %d sweeps, %.3f s -> %.0f replicas/s; converged %d failed %d; mean iterations %.1f" % (
        len(systems), sweeps, dt, len(systems) / dt, st["converged"].sum(), st["failed"].sum(), st["iterations"].mean()), flush=True)
    t0 = time.perf_counter()
    devs = DeviceReplicaQNOptimizer(ReplicaBatch(replicas(names, per, 0.3, 0.02)), *(lambda ss: (np.stack([s.pos for s in ss]), np.stack([np.array(s.domain.rvecs) for s in ss])))(replicas(names, per, 0.3, 0.02)),
                                    dof="strain", gpos_rms=1e-8, dpos_rms=1e-6, grvecs_rms=1e-8, drvecs_rms=1e-6)
    ssw = devs.run(2000)
    dts = time.perf_counter() - t0
    sts = devs._fetch()
    print("device QN, strain DOF: %d replicas, %d sweeps, %.3f s -> %.0f replicas/s; converged %d failed %d" % (
        len(systems), ssw, dts, len(systems) / dts, sts["converged"].sum(), sts["failed"].sum()), flush=True)
    if per == 128:
        t0 = time.perf_counter()
        host = ReplicaQNOptimizer(ReplicaBatch(replicas(names, per, 0.5)), pos0, rvecs0, dof="cartesian", gpos_rms=1e-7, dpos_rms=1e-5)
        hs = host.run(1000)
        dt = time.perf_counter() - t0
        print("host-driven lockstep: %d replicas, %d sweeps, %.3f s -> %.0f replicas/s" % (len(systems), hs, dt, len(systems) / dt), flush=True)
PY
cat $O/qn_bench.log | tail -8
