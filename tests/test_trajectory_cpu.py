"""Trajectory hooks (micmec_b200/sampling/trajectory.py) driven by the host-driven integrator on the oracle-backed
force part: rows written at the right iterations, state items complete, system dumped once."""
import numpy as np

from fakeh5 import Group
from oraclepart import OracleForcePart


def test_hdf5_and_xyz_writers(tmp_path):
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.pes.mmff import MicMecForceField
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import LangevinThermostat
    from micmec_b200.sampling.trajectory import HDF5Writer, XYZWriter
    from micmec_b200.units import femtosecond, angstrom

    system = System.periodic_grid((2, 2, 2), TYPE_FCU, explicit=True)
    mmf = MicMecForceField(system, [OracleForcePart(system)])
    f = Group()
    xyz = str(tmp_path / "traj.xyz")
    np.random.seed(1)
    verlet = VerletIntegrator(mmf, timestep=10 * femtosecond, temp0=300.0,
                              hooks=[HDF5Writer(f, step=2), XYZWriter(xyz, step=5), LangevinThermostat(300.0)])
    verlet.run(10)
    assert np.array_equal(f["system"]["pos"][:].shape, (8, 3)) and "masses" in f["system"]
    traj = f["trajectory"]
    assert traj["pos"].shape == (6, 8, 3) and traj["vel"].shape == (6, 8, 3) and traj["domain"].shape == (6, 3, 3)
    assert list(traj["counter"][:]) == [0, 2, 4, 6, 8, 10]
    assert np.allclose(traj["pos"][-1], verlet.pos) and traj["epot"][-1] == verlet.epot
    assert traj.attrs["ndof"] == verlet.ndof
    assert np.all(np.diff(traj["time"][:]) > 0)
    lines = open(xyz).read().splitlines()
    assert len(lines) == 3 * (8 + 2) and lines[0].strip() == "8" and lines[2].split()[0] == "Cs"
    last = np.array([[float(v) for v in line.split()[1:]] for line in lines[-8:]])
    assert np.allclose(last * angstrom, verlet.pos, atol=1e-5)
