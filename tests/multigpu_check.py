#!/usr/bin/env python
"""Multi-GPU parity check of the z-slab decomposition (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

Every rank integrates its slab of a perturbed 40 x 24 x 32 fcu grid ($MULTIGPU_SHAPE; 40 x 24 x 64 for 8 slabs) (NVE, NVT and NPT, 30 steps); rank 0
additionally runs the WHOLE grid on its own GPU with the same kernels.  Gathered positions / velocities and all
scalars must agree to 1e-12 (the only difference is the order of the 16-double reductions).
Prints "multigpu ok" on success.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(system, ens, vel0, slab=None, device=0, ndof=None):
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import NHCThermostat
    from micmec_b200.sampling.npt import MTKBarostat, TBCombination
    from micmec_b200.units import femtosecond, pascal

    part = ForcePartMechanical(system, model=os.environ.get("MULTIGPU_MODEL", "original"), device=device, structured=True, slab=slab)
    mmf = MicMecForceField(system, [part])
    hooks = []
    cvel = np.array([1e-4, -2e-4, 5e-5])
    vp0 = 1e-7 * np.array([[1.0, 0.2, -0.1], [0.2, -0.5, 0.3], [-0.1, 0.3, 0.8]])
    if ens in ("nvt", "npt"):
        thermo = NHCThermostat(300.0, timecon=100 * femtosecond, chain_vel0=cvel, chain_pos0=np.zeros(3), restart=True)
        hooks = [thermo]
    if ens == "npt":
        baro = MTKBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e5 * femtosecond, vel_press0=vp0, restart=True)
        hooks = [TBCombination(thermo, baro)]
    return part, mmf, hooks, dict(timestep=10 * femtosecond, vel0=vel0, ndof=ndof)


def main():
    import torch
    import torch.distributed as dist
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.system import System
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200 import slab as slabmod

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    shape = tuple(int(v) for v in os.environ.get("MULTIGPU_SHAPE", "40,24,32").split(","))
    rng = np.random.default_rng(11)  # same stream on every rank: everybody knows the global arrays
    full = System.periodic_grid(shape, TYPE_FCU, explicit=False)
    pos = full.pos + 0.3 * rng.standard_normal(full.pos.shape)
    vel = 2e-5 * rng.standard_normal(full.pos.shape)
    vel -= vel.mean(axis=0)
    nglobal = full.nnodes
    layout = slabmod.SlabLayout(shape, rank, world)
    worst = 0.0
    for ens in ("nve", "nvt", "npt"):
        ndof = 3 * nglobal - (3 if ens != "nve" else 0)
        # ---- decomposed run -----------------------------------------------------------------------------------
        lsys = slabmod.local_system(layout, TYPE_FCU, pos=pos)
        part, mmf, hooks, kw = build(lsys, ens, layout.take(vel), slab=layout.slab_arg(), device=local_rank, ndof=ndof)
        slabmod.init_comm(part, layout)
        g0, v0 = np.zeros(lsys.pos.shape), np.zeros((3, 3))
        e_dec = mmf.compute(g0, v0)  # plugin API on a slab: global energy / virial, local gradient
        verlet = VerletIntegrator(mmf, hooks=hooks, **kw)
        verlet.run(30)
        gp = slabmod.gather_nodes(layout, verlet.pos)
        gv = slabmod.gather_nodes(layout, verlet.vel)
        gg = slabmod.gather_nodes(layout, g0)
        scal = np.array([verlet.epot, verlet.ekin, verlet.econs, verlet.temp, verlet.press, verlet.rmsd_delta,
                         verlet.rmsd_gpos, e_dec] + list(verlet.vtens.ravel()) + list(np.asarray(verlet.rvecs).ravel()))
        # ---- the same grid on one GPU -------------------------------------------------------------------------
        if rank == 0:
            fsys = System.periodic_grid(shape, TYPE_FCU, explicit=False)
            fsys.pos[:] = pos
            fpart, fmmf, fhooks, fkw = build(fsys, ens, vel, device=local_rank, ndof=ndof)
            gf, vf = np.zeros(pos.shape), np.zeros((3, 3))
            e_one = fmmf.compute(gf, vf)
            one = VerletIntegrator(fmmf, hooks=fhooks, **fkw)
            one.run(30)
            ref = np.array([one.epot, one.ekin, one.econs, one.temp, one.press, one.rmsd_delta, one.rmsd_gpos, e_one]
                           + list(one.vtens.ravel()) + list(np.asarray(one.rvecs).ravel()))

            def rel(a, b):
                return float(np.max(np.abs(a - b)) / np.sqrt(np.mean(b * b)))

            names = ["epot", "ekin", "econs", "temp", "press", "rmsd_delta", "rmsd_gpos", "e_compute"] + \
                    ["vtens%d" % i for i in range(9)] + ["rvecs%d" % i for i in range(9)]
            # the pressure is a small difference of the kinetic and virial terms: compare it on the scale of those
            pscale = float(np.sqrt(np.mean(one.vtens ** 2)) / one.mmf.system.domain.volume)
            scale = np.maximum(np.abs(ref), 1e-300)
            scale[4] = max(scale[4], pscale)
            scale[8:17] = np.sqrt(np.mean(one.vtens ** 2))  # virial entries on the scale of the tensor (1e-9 of RMS)
            serr = np.abs(scal - ref) / scale
            errs = dict(pos=rel(gp, one.pos), vel=rel(gv, one.vel), gpos0=rel(gg, gf), vtens0=rel(v0, vf),
                        scalars=float(np.max(serr)))
            print("   worst scalar:", names[int(np.argmax(serr))], flush=True)
            print(ens, {k: "%.2e" % v for k, v in errs.items()}, flush=True)
            # arrays to 1e-10; reduced scalars (pressure, virial: differences of large sums whose order changes with
            # the decomposition, fed back through the barostat) to 1e-8 - the north_star's trajectory tolerance
            worst = max(worst, errs["pos"], errs["vel"], errs["gpos0"], errs["vtens0"], 1e-2 * errs["scalars"])
        del verlet, mmf, part
    flag = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if float(flag.item()) > 1e-10:
        print("multigpu FAILED: worst relative deviation %.3e" % float(flag.item()))
        sys.exit(1)
    if rank == 0:
        print("multigpu ok (%d slabs, worst relative deviation %.2e)" % (world, float(flag.item())))


if __name__ == "__main__":
    main()
