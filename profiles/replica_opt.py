#!/usr/bin/env python
"""Config 5 (BASELINE.json configs[4]): geometry optimisation of 10 defect configurations x 1024 perturbed replicas of
the 27-node 3x3x3 system, in lockstep (micmec_b200.sampling.batchopt.ReplicaQNOptimizer on ReplicaBatch).

    python profiles/replica_opt.py [--nrep 10240] [--dof cartesian|strain] [--cpu-sample 6]

Prints one JSON line: sweeps (= batched force evaluations), wall time, the split between the batched force evaluation on
the GPU and the per-replica trust-radius algebra, converged fraction, and - as the reported CPU baseline - the same
optimisation of a few replicas with the single-system QNOptimizer on the CPU oracle (test infrastructure, timed only).
Three of the type maps are the reference's conf0 / conf3 / conf9 (tests/golden), the others random two-type maps."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class TimedBatch(object):
    def __init__(self, batch):
        self.batch, self.seconds, self.calls = batch, 0.0, 0

    def compute(self, *args, **kwargs):
        t0 = time.perf_counter()
        out = self.batch.compute(*args, **kwargs)
        self.seconds += time.perf_counter() - t0
        self.calls += 1
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nrep", type=int, default=10240)
    ap.add_argument("--dof", default="cartesian", choices=["cartesian", "strain"])
    ap.add_argument("--cpu-sample", type=int, default=6)
    ap.add_argument("--max-sweeps", type=int, default=200)
    args = ap.parse_args()

    import goldenio as gio
    from micmec_b200.system import System
    from micmec_b200.replicas import ReplicaBatch
    from micmec_b200.sampling.batchopt import ReplicaQNOptimizer

    recs = [gio.system_from(gio.load("force_3x3x3_conf%d" % c)) for c in (0, 3, 9)]
    rec = recs[0]
    rng = np.random.default_rng(0)
    confs = []
    for c in range(10):
        types = np.asarray(recs[c].types).astype(int).ravel() if c < 3 else rng.integers(1, 3, size=27)
        confs.append(System(rec.pos, rec.masses, rec.rvecs, rec.surrounding_cells, rec.surrounding_nodes,
                            grid=np.asarray(rec.grid), types=types, params=rec.params))
    nrep = args.nrep
    systems = [confs[r % 10] for r in range(nrep)]
    pos0 = np.stack([rec.pos] * nrep) + 0.3 * rng.standard_normal((nrep, 27, 3))
    strain = np.eye(3) + (0.01 * rng.standard_normal((nrep, 3, 3)) if args.dof == "strain" else np.zeros((nrep, 3, 3)))
    strain = 0.5 * (strain + strain.transpose(0, 2, 1))
    rvecs0 = np.stack([rec.rvecs] * nrep) @ strain
    pos0 = pos0 @ strain
    kwargs = dict(gpos_rms=1e-7, dpos_rms=1e-5) if args.dof == "cartesian" else dict(
        gpos_rms=1e-8, dpos_rms=1e-6, grvecs_rms=1e-8, drvecs_rms=1e-6)

    batch = TimedBatch(ReplicaBatch(systems))
    t0 = time.perf_counter()
    opt = ReplicaQNOptimizer(batch, pos0, rvecs0, dof=args.dof, **kwargs)
    sweeps = opt.run(args.max_sweeps)
    wall = time.perf_counter() - t0

    cpu = None
    if args.cpu_sample > 0:
        from oraclepart import OracleForcePart
        from micmec_b200.pes.mmff import MicMecForceField
        from micmec_b200.sampling.dof import CartesianDOF, StrainCellDOF
        from micmec_b200.sampling.opt import QNOptimizer

        t1 = time.perf_counter()
        worst = 0.0
        for r in range(args.cpu_sample):
            s = confs[r % 10]
            system = System(pos0[r].copy(), s.masses, rvecs0[r].copy(), s.surrounding_cells, s.surrounding_nodes,
                            grid=s.grid, types=s.types, params=s.params)
            mmf = MicMecForceField(system, [OracleForcePart(system)])
            dof = (CartesianDOF if args.dof == "cartesian" else StrainCellDOF)(mmf, **kwargs)
            single = QNOptimizer(dof)
            single.run(args.max_sweeps)
            worst = max(worst, abs(single.f - opt.f[r]))
        dt = time.perf_counter() - t1
        cpu = {"replicas": args.cpu_sample, "seconds": dt, "replicas_per_s": args.cpu_sample / dt, "cores": 1,
               "kind": "port (single-system QNOptimizer on the CPU oracle)", "max_abs_energy_diff_vs_batch": worst}

    print(json.dumps({
        "workload": "geometry optimisation, %d replicas x 27 nodes (10 type maps), %s DOF" % (nrep, args.dof),
        "sweeps": sweeps, "force_evaluations": batch.calls, "wall_s": wall, "replicas_per_s": nrep / wall,
        "force_s": batch.seconds, "algebra_s": wall - batch.seconds, "ms_per_batched_force_eval": 1e3 * batch.seconds / batch.calls,
        "converged_fraction": float(opt.converged.mean()), "failed": int(opt.failed.sum()),
        "accepted_steps_mean": float(opt.iterations.mean()), "accepted_steps_max": int(opt.iterations.max()),
        "final_energy_max": float(opt.f.max()), "cpu_baseline": cpu,
    }))


if __name__ == "__main__":
    main()
