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


def test_raw_streaming_writer_round_trip(tmp_path):
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.pes.mmff import MicMecForceField
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.trajectory import RawWriter, load_raw
    from micmec_b200.units import femtosecond

    system = System.periodic_grid((3, 2, 2), TYPE_FCU, explicit=True)
    mmf = MicMecForceField(system, [OracleForcePart(system)])
    directory = str(tmp_path / "traj")
    np.random.seed(2)
    frames = []

    class Keep(object):
        def expects_call(self, counter):
            return counter % 3 == 0

        def __call__(self, it):
            frames.append((it.counter, it.pos.copy(), it.vel.copy(), it.epot))

    verlet = VerletIntegrator(mmf, timestep=10 * femtosecond, temp0=300.0, hooks=[RawWriter(directory, step=3), Keep()])
    verlet.run(9)
    traj = load_raw(directory)
    assert traj["pos"].shape == (4, 12, 3) and traj["domain"].shape == (4, 3, 3) and traj["attrs"]["ndof"] == verlet.ndof
    for row, (counter, pos, vel, epot) in enumerate(frames):
        assert traj["counter"][row] == counter and traj["epot"][row] == epot
        assert np.array_equal(traj["pos"][row], pos) and np.array_equal(traj["vel"][row], vel)
    # an interrupted frame (bytes beyond the last complete one) is ignored
    with open(directory + "/pos.bin", "ab") as handle:
        handle.write(b"\0" * 40)
    assert load_raw(directory, mmap=False)["pos"].shape == (4, 12, 3)
    only = RawWriter(str(tmp_path / "few"), keys=["epot", "time"])
    only(verlet)
    assert sorted(k for k in load_raw(str(tmp_path / "few")) if k != "attrs") == ["epot", "time"]
