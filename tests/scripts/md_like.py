#!/usr/bin/env python
"""An MD driver written against the REFERENCE's module names and call pattern (what simulations/md.py does: NVE,
Langevin NVT, Langevin NPT with an HDF5 trajectory and a screen log).  It imports ``micmec.*`` and ``molmod.units``
only; tests run it through ``python -m micmec_b200.dropin`` to show that such a script needs no edits."""
import argparse

import h5py

from micmec.system import System
from micmec.pes.mmff import MicMecForceField, ForcePartMechanical
from micmec.sampling.verlet import VerletIntegrator, VerletScreenLog
from micmec.sampling.trajectory import HDF5Writer
from micmec.sampling.nvt import LangevinThermostat
from micmec.sampling.npt import TBCombination, LangevinBarostat
from molmod.units import kelvin, pascal, femtosecond

parser = argparse.ArgumentParser()
parser.add_argument("input_fn")
parser.add_argument("output_fn")
parser.add_argument("-steps", type=int, default=20)
parser.add_argument("-temp", type=float, default=None)
parser.add_argument("-press", type=float, default=None)
args = parser.parse_args()

system = System.from_file(args.input_fn)
mmf = MicMecForceField(system, [ForcePartMechanical(system)])
dt = 10 * femtosecond
with h5py.File(args.output_fn, mode="w") as f:
    hooks = [HDF5Writer(f, step=5), VerletScreenLog(step=10)]
    if args.temp is not None:
        thermo = LangevinThermostat(temp=args.temp * kelvin, timecon=100 * dt)
        if args.press is None:
            hooks.append(thermo)
        else:
            baro = LangevinBarostat(mmf, temp=args.temp * kelvin, press=args.press * 1e6 * pascal, timecon=1e4 * dt)
            hooks.append(TBCombination(thermo, baro))
    verlet = VerletIntegrator(mmf, timestep=dt, hooks=hooks, temp0=(args.temp or 300.0) * kelvin)
    verlet.run(args.steps)
print("md_like done: device_mode=%s counter=%d temp=%.3f epot=%.10e" % (verlet.device_mode, verlet.counter, verlet.temp, verlet.epot))
