#!/usr/bin/env python
"""Small structured-grid workload for compute-sanitizer (racecheck / memcheck / synccheck): a 40 x 24 x 32 fcu grid,
one compute(), then NVE, NVT and NPT steps through the device-resident integrator (k_march STEP / FORCE variants, halo
kernels, k_scalar, CUDA-graph replay of the lean step pairs).  Under torchrun the same grid runs as z-slabs.

    compute-sanitizer --tool racecheck python profiles/r02/sanitize_run.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.system import System
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import NHCThermostat
    from micmec_b200.sampling.npt import MTKBarostat, TBCombination
    from micmec_b200.units import femtosecond, pascal
    from micmec_b200 import slab as slabmod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps = int(os.environ.get("SAN_STEPS", "7"))
    shape = (40, 24, 32)
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rng = np.random.default_rng(5)
    full = System.periodic_grid(shape, TYPE_FCU, explicit=False)
    pos = full.pos + 0.3 * rng.standard_normal(full.pos.shape)
    vel = 2e-5 * rng.standard_normal(full.pos.shape)
    vel -= vel.mean(axis=0)
    layout = slabmod.SlabLayout(shape, rank, world)
    for ens in ("nve", "nvt", "npt"):
        if world > 1:
            system = slabmod.local_system(layout, TYPE_FCU, pos=pos)
            part = ForcePartMechanical(system, device=local_rank, structured=True, slab=layout.slab_arg())
        else:
            system = System.periodic_grid(shape, TYPE_FCU, explicit=False)
            system.pos[:] = pos
            part = ForcePartMechanical(system, device=local_rank, structured=True)
        mmf = MicMecForceField(system, [part])
        if world > 1:
            slabmod.init_comm(part, layout)
        gpos, vtens = np.zeros(system.pos.shape), np.zeros((3, 3))
        energy = mmf.compute(gpos, vtens)
        hooks = []
        cvel = np.array([1e-4, -2e-4, 5e-5])
        vp0 = 1e-7 * np.array([[1.0, 0.2, -0.1], [0.2, -0.5, 0.3], [-0.1, 0.3, 0.8]])
        if ens in ("nvt", "npt"):
            thermo = NHCThermostat(300.0, timecon=100 * femtosecond, chain_vel0=cvel, chain_pos0=np.zeros(3), restart=True)
            hooks = [thermo]
        if ens == "npt":
            baro = MTKBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e5 * femtosecond, vel_press0=vp0, restart=True)
            hooks = [TBCombination(thermo, baro)]
        ndof = 3 * full.nnodes - (3 if ens != "nve" else 0)
        verlet = VerletIntegrator(mmf, timestep=10 * femtosecond, hooks=hooks, vel0=layout.take(vel), ndof=ndof)
        verlet.run(steps)      # direct launches
        verlet.run(steps + 2)  # steps_done >= 2: the middle pairs replay as CUDA graphs
        if rank == 0:
            print("%s: E0=%.12e econs=%.12e launches=%d" % (ens, energy, verlet.econs, part.launches), flush=True)
        del verlet, mmf, part
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("sanitize_run done")


if __name__ == "__main__":
    main()
