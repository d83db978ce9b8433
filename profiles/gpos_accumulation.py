#!/usr/bin/env python
"""Force-only evaluation of a G^3 fcu grid with the three gpos accumulation schemes (run under ncu for counters):

    gather   k_cells (24 gradient doubles per cell to HBM) + k_gather (node-centric, deterministic)
    scatter  k_cells_scatter (cell-centric, warp-aggregated RED.ADD.F64 into gpos)
    march    k_march<FORCE> (structured grid: separable in-tile gather, nothing per cell in HBM)

Prints the CUDA-event time of `reps` evaluations with positions resident on the device."""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from micmec_b200 import _lib
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.pes.mmff import ForcePartMechanical

    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    lib = _lib.load()
    system = System.periodic_grid((args.grid,) * 3, TYPE_FCU, explicit=False)
    rng = np.random.default_rng(0)
    pos = torch.from_numpy(system.pos + 0.1 * rng.standard_normal(system.pos.shape)).cuda()
    gpos = torch.zeros_like(pos)
    rvecs = np.array(system.domain.rvecs)
    ref = None
    for name, kw in (("gather", dict(structured=False)), ("scatter", dict(structured=False, scatter=True)),
                     ("march", dict(structured=True))):
        part = ForcePartMechanical(system, **kw)
        stream = torch.cuda.Stream()
        _lib.check(lib.mm_set_stream(part.handle, ctypes.c_void_p(stream.cuda_stream)))
        e, _ = part.compute_device(pos, rvecs, gpos=gpos, vtens=True)
        e, _ = part.compute_device(pos, rvecs, gpos=gpos, vtens=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.reps):
            e, _ = part.compute_device(pos, rvecs, gpos=gpos, vtens=True)
        ev1.record(stream)
        torch.cuda.synchronize()
        g = gpos.cpu().numpy()
        if ref is None:
            ref = g.copy()
        dev = float(np.max(np.abs(g - ref)) / np.sqrt(np.mean(ref ** 2)))
        print("%-8s %8.3f ms / evaluation   E = %.12e   max|dgpos|/rms vs gather = %.1e" % (
            name, ev0.elapsed_time(ev1) / args.reps, e, dev), flush=True)
        del part


if __name__ == "__main__":
    main()
