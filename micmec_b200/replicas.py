"""Batches of independent replicas (BASELINE.json config 5: defect configurations x perturbations).

The reference evaluates one small system per Python call.  ``ReplicaBatch`` concatenates R systems of identical size
(e.g. the 27-node ``3x3x3_confN`` systems, which share cell-type parameters and differ in their type maps) into one
device-resident batch: one kernel sequence evaluates all of them, each with its OWN positions and domain vectors,
and returns per-replica energies, gradients and virials.  Replicas never communicate, so a multi-GPU run simply
gives each rank its share of the replicas ("replicas only").
"""
import ctypes

import numpy as np

from . import _lib
from .units import boltzmann
from .topology import cell_shifts, type_tables

__all__ = ["ReplicaBatch"]


class ReplicaBatch(object):
    """R independent systems evaluated together.

    Parameters
    ----------
    systems : list of System
        Systems with the same number of nodes and cells and the same ``params`` (type maps may differ).
    model, device : as for ``ForcePartMechanical``.
    """

    def __init__(self, systems, model="original", device=0):
        first = systems[0]
        self.nrep = len(systems)
        self.nnodes, self.ncells = first.nnodes, first.ncells
        if any(s.nnodes != self.nnodes or s.ncells != self.ncells for s in systems):
            raise ValueError("All replicas of a batch must have the same number of nodes and cells.")
        pbc = first.domain.rvecs.shape[0] > 0
        types = np.concatenate([np.asarray(s.types).ravel() for s in systems])
        tab = type_tables(first.params, types)
        sn = np.concatenate([np.asarray(s.surrounding_nodes, dtype=np.int64) + r * self.nnodes for r, s in enumerate(systems)])
        sc_parts = []
        for r, s in enumerate(systems):
            sc = np.asarray(s.surrounding_cells, dtype=np.int64).copy()
            sc[sc >= 0] += r * self.ncells
            sc_parts.append(sc)
        sc = np.concatenate(sc_parts)
        shift = np.concatenate([cell_shifts(s.grid, s.ncells, pbc) for s in systems])
        sn, sc, shift = (np.ascontiguousarray(a) for a in (sn, sc, shift))
        desc = _lib.Desc()
        desc.nnodes, desc.ncells = self.nrep * self.nnodes, self.nrep * self.ncells
        desc.surrounding_nodes, desc.surrounding_cells, desc.shift = _lib.ptr(sn), _lib.ptr(sc), _lib.ptr(shift)
        desc.cell_type = _lib.ptr(tab["cell_type"])
        desc.ntypes = len(tab["type_nstates"])
        desc.type_nstates = _lib.ptr(tab["type_nstates"])
        desc.h0, desc.elasticity = _lib.ptr(tab["h0"]), _lib.ptr(tab["elasticity"])
        desc.free_energy, desc.effective_temp = _lib.ptr(tab["free_energy"]), _lib.ptr(tab["effective_temp"])
        desc.boltzmann = boltzmann
        desc.model = _lib.MODELS[model]
        desc.device = int(device)
        desc.nreplicas = self.nrep
        self._lib = _lib.load()
        self._handle = ctypes.c_void_p()
        _lib.check(self._lib.mm_create(ctypes.byref(desc), ctypes.byref(self._handle)))
        self.masses = np.stack([np.asarray(s.masses, dtype=float) for s in systems])

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle is not None and handle.value:
            self._lib.mm_destroy(handle)
            self._handle = ctypes.c_void_p()

    @property
    def launches(self):
        return int(self._lib.mm_launch_count(self._handle))

    def compute(self, pos, rvecs, gpos=True, vtens=True):
        """Evaluate every replica.

        pos [R, nnodes, 3] and rvecs [R, 3, 3] (host arrays).  Returns ``(energies [R], gpos [R, nnodes, 3] or None,
        vtens [R, 3, 3] or None)``; raises ``ValueError`` when any replica produced a NaN, as the reference would.
        """
        pos = np.ascontiguousarray(pos, dtype=float).reshape(self.nrep * self.nnodes, 3)
        rvecs = np.ascontiguousarray(rvecs, dtype=float).reshape(self.nrep, 9)
        _lib.check(self._lib.mm_set_rvecs_batch(self._handle, _lib.ptr(rvecs)))
        _lib.check(self._lib.mm_set_pos(self._handle, _lib.ptr(pos), _lib.MM_HOST))
        g = np.zeros((self.nrep * self.nnodes, 3)) if gpos else None
        total = ctypes.c_double()
        vsum = np.zeros((3, 3))
        _lib.check(self._lib.mm_compute(self._handle, ctypes.byref(total), _lib.ptr(g), _lib.MM_HOST, _lib.ptr(vsum)))
        energies = np.zeros(self.nrep)
        vt = np.zeros((self.nrep, 3, 3)) if vtens else None
        _lib.check(self._lib.mm_get_replica_results(self._handle, _lib.ptr(energies), _lib.ptr(vt)))
        return energies, (g.reshape(self.nrep, self.nnodes, 3) if gpos else None), vt
