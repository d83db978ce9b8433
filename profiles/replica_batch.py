#!/usr/bin/env python
"""Config 5 in full size: 10 defect configurations x 1024 perturbed, strained replicas of the 27-node 3x3x3 system
evaluated as one batch (energy + gradient + virial per replica).  Needs tests/golden/force_3x3x3_conf*.npz for three of
the configurations; the other type maps are random two-type maps with the same parameters."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import goldenio as gio
    from micmec_b200.system import System
    from micmec_b200.replicas import ReplicaBatch

    d = gio.load("force_3x3x3_conf3")
    rec = gio.system_from(d)
    rng = np.random.default_rng(0)
    confs = []
    for c in range(10):
        types = rng.integers(1, 3, size=27) if c not in (0,) else np.ones(27, dtype=int)
        confs.append(System(rec.pos, rec.masses, rec.rvecs, rec.surrounding_cells, rec.surrounding_nodes, grid=types.reshape(3, 3, 3),
                            types=types, params=rec.params))
    nrep = 10240
    systems = [confs[r % 10] for r in range(nrep)]
    pos = np.stack([rec.pos] * nrep) + 0.3 * rng.standard_normal((nrep, 27, 3))
    rvecs = np.stack([rec.rvecs] * nrep) * (1.0 + 0.01 * rng.standard_normal((nrep, 1, 1)))
    t0 = time.perf_counter()
    batch = ReplicaBatch(systems)
    t1 = time.perf_counter()
    batch.compute(pos, rvecs)
    reps = 20
    t2 = time.perf_counter()
    for _ in range(reps):
        e, g, v = batch.compute(pos, rvecs)
    t3 = time.perf_counter()
    per = (t3 - t2) / reps
    print("replica batch: %d replicas x 27 nodes; setup %.2f s; %.3f ms per batched compute (host arrays in/out) = %.3e replica-evaluations/s"
          % (nrep, t1 - t0, 1e3 * per, nrep / per))
    print("reference: ~3.9 ms per 27-cell compute on one core (BASELINE.md) -> %.1f s for the same batch" % (3.9e-3 * nrep))


if __name__ == "__main__":
    main()
