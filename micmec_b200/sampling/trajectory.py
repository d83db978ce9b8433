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

__all__ = ["HDF5Writer", "XYZWriter", "RawWriter", "DeviceRawWriter", "load_raw"]


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


class DeviceRawWriter(Hook):
    """Streaming writer for the device-resident integrator at benchmark sizes (the reference's
    ``HDF5Writer`` copies the full ``pos`` array on every call, trajectory.py:31-85).

    Same files as ``RawWriter`` (``load_raw`` reads them), but the frame never passes through the integrator's host
    mirrors: ``mm_md_get_state`` exports positions (and velocities / gradients on request) straight from the device
    layout into one of two pinned host buffers, and a background thread appends that buffer to ``<key>.bin`` while the
    integrator is already running the next steps.  The integrator does not refresh ``iterative.pos / vel / gpos`` for
    this hook (``wants_arrays = False``): a 256^3 frame costs one 403 MB device-to-host copy per field and nothing else
    on the critical path.  Scalars (``time``, ``epot``, ``ekin``, ``temp``, ``econs``, ``counter``) go to ``scalars.bin``.
    """

    wants_arrays = False
    SCALARS = ("counter", "time", "epot", "ekin", "temp", "etot", "econs", "press", "volume")

    def __init__(self, directory, fields=("pos",), start=0, step=1):
        import queue
        import threading

        if not set(fields) <= {"pos", "vel", "gpos"}:
            raise ValueError("DeviceRawWriter writes pos, vel and / or gpos")
        self.directory, self.fields = directory, tuple(fields)
        self.meta = None
        self._buffers, self._turn = None, 0
        self._queue = queue.Queue()
        self._free = [threading.Semaphore(1), threading.Semaphore(1)]
        self._thread = None
        self._error = None
        Hook.__init__(self, start, step)

    @staticmethod
    def _pinned(shape):
        try:  # pinned memory makes the device-to-host copy asynchronous-capable and twice as fast; torch is plumbing only
            import torch

            tensor = torch.empty(shape, dtype=torch.float64).pin_memory()
            return tensor.numpy(), tensor
        except Exception:
            return np.empty(shape), None

    def _writer_loop(self):
        while True:
            job = self._queue.get()
            if job is None:
                return
            slot, nframes, scal = job
            try:
                for key in self.fields:
                    with open(os.path.join(self.directory, key + ".bin"), "ab") as handle:
                        self._buffers[slot][key][0].tofile(handle)
                with open(os.path.join(self.directory, "scalars.bin"), "ab") as handle:
                    np.asarray(scal, dtype=float).tofile(handle)
                self.meta["frames"] = nframes
                tmp = os.path.join(self.directory, "meta.json.tmp")
                with open(tmp, "w") as handle:
                    json.dump(self.meta, handle)
                os.replace(tmp, os.path.join(self.directory, "meta.json"))
            except Exception as exc:  # surfaced by the next call / close()
                self._error = exc
            finally:
                self._free[slot].release()

    def _open(self, iterative):
        import threading

        os.makedirs(self.directory, exist_ok=True)
        shape = tuple(iterative.pos.shape)
        self.meta = {"frames": 0, "attrs": {"scalars": list(self.SCALARS)},
                     "items": {key: {"shape": list(shape), "dtype": "<f8"} for key in self.fields}}
        self.meta["items"]["scalars"] = {"shape": [len(self.SCALARS)], "dtype": "<f8"}
        for key in list(self.fields) + ["scalars"]:
            open(os.path.join(self.directory, key + ".bin"), "wb").close()
        np.save(os.path.join(self.directory, "system_masses.npy"), np.asarray(iterative.mmf.system.masses))
        self._buffers = [{key: self._pinned(shape) for key in self.fields} for _ in range(2)]
        self._thread = threading.Thread(target=self._writer_loop, daemon=True)
        self._thread.start()
        self._count = 0
        import atexit

        atexit.register(self.close)  # frames still queued when the interpreter exits are written, not dropped

    def __call__(self, iterative):
        from .. import _lib

        if self._error is not None:
            raise self._error
        if not getattr(iterative, "device_mode", False):
            raise RuntimeError("DeviceRawWriter needs the device-resident integrator; use RawWriter in host-driven mode")
        if self.meta is None:
            self._open(iterative)
        slot = self._turn
        self._free[slot].acquire()  # the background thread has finished with this buffer
        bufs = self._buffers[slot]
        ptrs = [(_lib.ptr(bufs[key][0]) if key in bufs else None) for key in ("pos", "vel", "gpos")]
        _lib.check(iterative._lib.mm_md_get_state(iterative._md, ptrs[0], ptrs[1], ptrs[2], _lib.MM_HOST, None, None, None, None))
        self._count += 1
        scal = [float(getattr(iterative, name, 0.0) if name != "volume" else iterative.mmf.system.domain.volume) for name in self.SCALARS]
        self._queue.put((slot, self._count, scal))
        self._turn ^= 1

    def close(self):
        """Wait until every frame is on disk."""
        if self._thread is not None:
            self._queue.put(None)
            self._thread.join()
            self._thread = None
        if self._error is not None:
            raise self._error


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
