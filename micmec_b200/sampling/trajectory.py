"""Trajectory writers: drop-in for ``micmec.sampling.trajectory`` (trajectory.py:31-128).

Conventional hooks: the device-resident integrator stops at the iterations where they fire, refreshes its host
mirrors (one D2H copy of pos / vel / gpos) and calls them with the same ``iterative.state`` items as the reference.
``HDF5Writer`` takes an open ``h5py.File`` (or anything with the same group / dataset interface); ``XYZWriter`` writes
plain XYZ frames itself (the reference delegates to ``molmod.io.XYZWriter``), nodes shown as caesium atoms, in angstrom.
"""
from ..units import angstrom
from .iterative import Hook

__all__ = ["HDF5Writer", "XYZWriter"]


class BaseHDF5Writer(Hook):
    def __init__(self, f, start=0, step=1):
        self.f = f
        Hook.__init__(self, start, step)

    @staticmethod
    def _skip(item):
        return item.value is None or (len(item.shape) > 0 and min(item.shape) == 0)

    def __call__(self, iterative):
        if "trajectory" not in self.f:
            self.init_trajectory(iterative)
        tgrp = self.f["trajectory"]
        # a row that was only partly written by an interrupted run is reused (trajectory.py:55-57)
        row = min(tgrp[key].shape[0] for key in iterative.state if key in tgrp.keys())
        for key, item in iterative.state.items():
            if self._skip(item):
                continue
            ds = tgrp[key]
            if ds.shape[0] <= row:
                ds.resize(row + 1, axis=0)
            ds[row] = item.value

    def dump_system(self, system, grp):
        system.to_hdf5(grp)

    def init_trajectory(self, iterative):
        tgrp = self.f.create_group("trajectory")
        for key, item in iterative.state.items():
            if self._skip(item):
                continue
            tgrp.create_dataset(key, (0,) + item.shape, maxshape=(None,) + item.shape, dtype=item.dtype)
            for name, value in item.iter_attrs(iterative):
                tgrp.attrs[name] = value


class HDF5Writer(BaseHDF5Writer):
    def __call__(self, iterative):
        if "system" not in self.f:
            self.dump_system(iterative.mmf.system, self.f)
        BaseHDF5Writer.__call__(self, iterative)


class XYZWriter(Hook):
    def __init__(self, fn_xyz, select=None, start=0, step=1):
        self.fn_xyz = fn_xyz
        self.select = select
        self.frames = 0
        Hook.__init__(self, start, step)

    def __call__(self, iterative):
        pos = iterative.mmf.system.pos
        if self.select is not None:
            pos = pos[self.select]
        with open(self.fn_xyz, "w" if self.frames == 0 else "a") as handle:
            handle.write("%5i\n%7i E_pot = %.10f     \n" % (len(pos), iterative.counter, iterative.epot))
            for x, y, z in pos / angstrom:
                handle.write("%2s %12.6f %12.6f %12.6f\n" % ("Cs", x, y, z))
        self.frames += 1
