"""Trajectory writers: drop-in for ``micmec.sampling.trajectory`` (trajectory.py:31-128).

Conventional hooks: the device-resident integrator stops at the iterations where they fire, refreshes its host
mirrors (one D2H copy of pos / vel / gpos) and calls them with the same ``iterative.state`` items as the reference.
``HDF5Writer`` takes an open ``h5py.File`` (or anything with the same group / dataset interface); ``XYZWriter`` writes
plain XYZ frames itself (the reference delegates to ``molmod.io.XYZWriter``), nodes shown as caesium atoms, in angstrom.

``RawWriter`` has no counterpart in the reference: a streaming writer for grids whose frames are hundreds of megabytes
(256^3: pos, vel, gpos are 403 MB each) and for machines without h5py.  Every state item is appended to its own flat binary
file, so a frame costs one sequential write per item and nothing is ever re-read or resized; ``load_raw`` memory-maps the
files back as ``[frames, *shape]`` arrays.
"""
from ..units import angstrom
from .iterative import Hook

import json
import os

import numpy as np

__all__ = ["HDF5Writer", "XYZWriter", "RawWriter", "load_raw"]


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


class RawWriter(Hook):
    """Append the integrator's state items to ``<directory>/<key>.bin`` (C order, native dtype), one frame per call;
    ``<directory>/meta.json`` records shape, dtype and the number of complete frames.  ``keys`` restricts what is written
    (default: every state item with a value)."""

    def __init__(self, directory, keys=None, start=0, step=1):
        self.directory = directory
        self.keys = None if keys is None else set(keys)
        self.meta = None
        Hook.__init__(self, start, step)

    def _open(self, iterative):
        os.makedirs(self.directory, exist_ok=True)
        self.meta = {"frames": 0, "items": {}, "attrs": {}}
        for key, item in iterative.state.items():
            if item.value is None or (self.keys is not None and key not in self.keys):
                continue
            value = np.asarray(item.value)
            self.meta["items"][key] = {"shape": list(value.shape), "dtype": value.dtype.str}
            open(os.path.join(self.directory, key + ".bin"), "wb").close()
            for name, attr in item.iter_attrs(iterative):
                self.meta["attrs"][name] = attr.tolist() if isinstance(attr, np.ndarray) else attr
        system = iterative.mmf.system
        np.save(os.path.join(self.directory, "system_masses.npy"), np.asarray(system.masses))

    def __call__(self, iterative):
        if self.meta is None:
            self._open(iterative)
        for key, spec in self.meta["items"].items():
            value = np.ascontiguousarray(iterative.state[key].value, dtype=np.dtype(spec["dtype"]))
            with open(os.path.join(self.directory, key + ".bin"), "ab") as handle:
                value.tofile(handle)
        self.meta["frames"] += 1
        tmp = os.path.join(self.directory, "meta.json.tmp")
        with open(tmp, "w") as handle:
            json.dump(self.meta, handle, default=lambda o: o.decode() if isinstance(o, bytes) else str(o))
        os.replace(tmp, os.path.join(self.directory, "meta.json"))  # a crash never leaves a half-written index


def load_raw(directory, mmap=True):
    """Read a ``RawWriter`` directory: dict of ``[frames, *shape]`` arrays (memory-mapped unless ``mmap=False``) plus
    ``"attrs"``.  Frames beyond the last complete one (an interrupted write) are ignored."""
    with open(os.path.join(directory, "meta.json")) as handle:
        meta = json.load(handle)
    out = {"attrs": meta["attrs"]}
    for key, spec in meta["items"].items():
        shape = (meta["frames"],) + tuple(spec["shape"])
        path = os.path.join(directory, key + ".bin")
        if mmap and meta["frames"] > 0:
            out[key] = np.memmap(path, dtype=np.dtype(spec["dtype"]), mode="r", shape=shape)
        else:
            count = int(np.prod(shape))
            out[key] = np.fromfile(path, dtype=np.dtype(spec["dtype"]), count=count).reshape(shape)
    return out
