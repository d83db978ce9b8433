"""``System`` and ``Domain``: the input side of the force-field boundary.

``System`` mirrors the constructor and attributes of ``micmec.system.System`` (micmec/system.py:38-111) that the
force field and the integrators read; ``Domain`` mirrors ``micmec.pes.ext.Domain`` (micmec/pes/ext.pyx:36-123,
micmec/pes/domain.c) and is backed by the native ``mm_domain`` entry point of libmicmec_b200.so.
A reference ``System`` object can be passed to ``ForcePartMechanical`` just as well - only attributes are read.
"""
import ctypes

import numpy as np

from . import _lib
from .chk import load_chk, dump_chk
from .topology import periodic_grid_arrays

__all__ = ["System", "Domain"]


class Domain(object):
    """Periodic boundary conditions: ``rvecs`` (rows a, b, c), ``gvecs``, ``volume``, ``nvec``."""

    def __init__(self, rvecs):
        self.update_rvecs(rvecs)

    def update_rvecs(self, rvecs):
        if rvecs is None or np.size(rvecs) == 0:
            self._rvecs = np.zeros((0, 3), float)
            self._gvecs = np.zeros((0, 3), float)
            self._volume = 0.0
            return
        rvecs = np.asarray(rvecs, dtype=float)
        if rvecs.ndim != 2 or rvecs.shape[0] > 3 or rvecs.shape[1] != 3:
            # ext.pyx:60-61
            raise TypeError("rvecs must be a C-contiguous array with three columns and at most three rows.")
        rvecs = np.ascontiguousarray(rvecs)
        gvecs = np.zeros_like(rvecs)
        volume = ctypes.c_double()
        _lib.check(_lib.load().mm_domain(_lib.ptr(rvecs), rvecs.shape[0], ctypes.byref(volume), _lib.ptr(gvecs)))
        self._rvecs, self._gvecs, self._volume = rvecs.copy(), gvecs, volume.value

    nvec = property(lambda self: self._rvecs.shape[0])
    volume = property(lambda self: self._volume)

    @property
    def rvecs(self):
        out = self._rvecs.copy()
        out.setflags(write=False)
        return out

    @property
    def gvecs(self):
        out = self._gvecs.copy()
        out.setflags(write=False)
        return out

    def _full(self, reciprocal):
        """3x3 completion of the cell (ext.pyx:56-72): identity for nvec = 0, the cell itself for nvec = 3, else the
        periodic vectors followed by an orthonormal complement (reciprocal: the matching dual basis)."""
        nvec = self.nvec
        if nvec == 3:
            return (self._gvecs if reciprocal else self._rvecs).copy()
        if nvec == 0:
            return np.identity(3)
        up, sp, vt = np.linalg.svd(self._rvecs, full_matrices=True)
        sing, u = np.ones(3), np.identity(3)
        sing[:nvec], u[:nvec, :nvec] = sp, up
        if reciprocal:
            return np.dot(u / sing, vt)
        full = np.dot(u * sing, vt)
        full[:nvec] = self._rvecs
        return full

    def _get_rvecs(self, full=False):
        out = self._full(False) if full else self._rvecs.copy()
        out.setflags(write=False)
        return out

    def _get_gvecs(self, full=False):
        out = self._full(True) if full else self._gvecs.copy()
        out.setflags(write=False)
        return out

    @property
    def parameters(self):
        """Lengths and angles (ext.pyx:108-121)."""
        rv = self._rvecs
        tmp = rv @ rv.T
        lengths = np.sqrt(np.diag(tmp))
        tmp = tmp / lengths / lengths.reshape(-1, 1)
        if len(rv) < 2:
            cosines = np.array([])
        elif len(rv) == 2:
            cosines = np.array([tmp[0, 1]])
        else:
            cosines = np.array([tmp[1, 2], tmp[2, 0], tmp[0, 1]])
        return lengths, np.arccos(np.clip(cosines, -1, 1))


class System(object):
    """A micromechanical system (same positional arguments as ``micmec.system.System``)."""

    def __init__(self, pos, masses, rvecs, surrounding_cells, surrounding_nodes, boundary_nodes=None, grid=None,
                 types=None, params=None, structured_shape=None):
        self.pos = pos
        self.masses = masses
        self.domain = Domain(rvecs)
        self.grid = grid
        self.types = types
        self.params = params
        self.surrounding_cells = surrounding_cells
        self.surrounding_nodes = surrounding_nodes
        self.boundary_nodes = boundary_nodes
        # full periodic grids built by ``periodic_grid`` may leave the index arrays implicit
        self.structured_shape = structured_shape
        if structured_shape is not None:
            self.nnodes = self.ncells = int(np.prod(structured_shape))
        else:
            self.nnodes = len(self.surrounding_cells)
            self.ncells = len(self.surrounding_nodes)

    def update_params(self, new_params, type_=1):
        for key, val in new_params.items():
            self.params["type%d/%s" % (int(type_), key)] = val

    @classmethod
    def from_file(cls, fn, **user_kwargs):
        """Load a ``.chk`` system (micmec/system.py:214-259)."""
        if not fn.endswith(".chk"):
            raise IOError("Cannot read from file '%s': only .chk files are supported here." % fn)
        allowed = ["pos", "masses", "rvecs", "surrounding_cells", "surrounding_nodes", "boundary_nodes", "grid", "types"]
        kwargs, params = {}, {}
        for key, value in load_chk(fn).items():
            if key in allowed:
                kwargs[key] = value
            elif key.startswith("type"):
                params[key] = value
        kwargs["params"] = params
        kwargs.update(user_kwargs)
        return cls(**kwargs)

    def to_file(self, fn):
        if not fn.endswith(".chk"):
            raise IOError("Cannot write to file '%s': only .chk files are supported here." % fn)
        sn, sc = self.surrounding_nodes, self.surrounding_cells
        if sn is None:
            sn, sc, _ = periodic_grid_arrays(self.structured_shape)
        output = {"pos": self.pos, "masses": self.masses, "rvecs": self.domain.rvecs, "surrounding_cells": sc,
                  "surrounding_nodes": sn, "boundary_nodes": self.boundary_nodes, "grid": self.grid,
                  "types": self.types}
        output.update(self.params)
        dump_chk(fn, {k: v for k, v in output.items() if v is not None})

    def to_hdf5(self, f):
        """Write ``system/pos`` and ``system/masses`` into an open, writable h5py file (system.py:352-370)."""
        if "system" in f:
            raise ValueError("The HDF5 file already contains a system description.")
        sgrp = f.create_group("system")
        sgrp.create_dataset("pos", data=self.pos)
        if self.masses is not None:
            sgrp.create_dataset("masses", data=self.masses)

    @classmethod
    def periodic_grid(cls, shape, type_params, explicit=None):
        """Full periodic ``nx x ny x nz`` grid of ONE cell type at rest, in closed form.

        Follows the conventions of ``build_system`` (micmec/utils.py:164-263): node (k, l, m) has id
        ``(k*ny + l)*nz + m`` and sits at ``(k, l, m) * diag(h0)``; ``rvecs = shape * diag(h0)``; every node mass is
        8 * 1/8 of the type mass.  ``type_params`` holds ``cell``, ``elasticity``, ``free_energy``,
        ``effective_temp`` and ``mass``.  Index arrays are materialised only when ``explicit`` (default: up to 64^3).
        """
        nx, ny, nz = (int(s) for s in shape)
        n = nx * ny * nz
        h0 = np.asarray(type_params["cell"], dtype=float).reshape(-1, 3, 3)[0]
        diag = np.diag(h0)
        k, l, m = np.meshgrid(np.arange(nx, dtype=float), np.arange(ny, dtype=float), np.arange(nz, dtype=float),
                              indexing="ij")
        pos = np.stack([k.ravel() * diag[0], l.ravel() * diag[1], m.ravel() * diag[2]], axis=1)
        rvecs = np.diag(np.array([nx, ny, nz], dtype=float) * diag)
        masses = np.full(n, float(type_params["mass"]))
        params = {"type1/" + key: type_params[key] for key in ("cell", "elasticity", "free_energy", "effective_temp", "mass")}
        if explicit is None:
            explicit = n <= 64 ** 3
        sn = sc = bn = None
        if explicit:
            sn, sc, bn = periodic_grid_arrays((nx, ny, nz))
        return cls(pos, masses, rvecs, sc, sn, boundary_nodes=bn, grid=np.ones((nx, ny, nz), dtype=np.int64),
                   types=np.ones(n, dtype=np.int64), params=params, structured_shape=(nx, ny, nz))
