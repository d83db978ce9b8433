"""The micromechanical force field on a B200: drop-in for ``micmec.pes.mmff``.

Same classes, constructor arguments, attributes and error behaviour as the reference module
(micmec/pes/mmff.py): ``ForcePart`` (:39-153), ``MicMecForceField`` (:156-201) and ``ForcePartMechanical``
(:204-323).  ``ForcePartMechanical._internal_compute`` runs on the GPU through the C ABI of
libmicmec_b200.so; there is no CPU code path.  Everything in ``micmec.sampling`` and the ``simulations/*.py``
scripts only ever touch ``mmf.system``, ``mmf.update_pos``, ``mmf.update_rvecs``, ``mmf.compute`` and
``mmf.parts``, which behave as in the reference.
"""
import ctypes
import os

import numpy as np

from .. import _lib
from ..units import boltzmann
from ..topology import cell_shifts, type_tables

__all__ = ["MicMecForceField", "ForcePart", "ForcePartMechanical"]


class ForcePart(object):
    """Base class of anything that computes an energy (and optionally gradient / virial) of a ``System``."""

    def __init__(self, name, system):
        self.name = name
        self.energy = 0.0
        self.gpos = np.zeros((system.nnodes, 3), float)
        self.vtens = np.zeros((3, 3), float)
        self.clear()

    def clear(self):
        """Mark the cached results invalid (filled with ``nan``), mmff.py:61-65."""
        self.energy = np.nan
        self.gpos[:] = np.nan
        self.vtens[:] = np.nan

    def update_rvecs(self, rvecs):
        self.clear()

    def update_pos(self, pos):
        self.clear()

    def compute(self, gpos=None, vtens=None):
        """Energy of this part; results are ADDED to ``gpos`` / ``vtens`` when given (mmff.py:87-149)."""
        my_gpos = my_vtens = None
        if gpos is not None:
            my_gpos = self.gpos
            my_gpos[:] = 0.0
        if vtens is not None:
            my_vtens = self.vtens
            my_vtens[:] = 0.0
        self.energy = self._internal_compute(my_gpos, my_vtens)
        if np.isnan(self.energy):
            raise ValueError("The energy is not-a-number (``nan``).")
        if gpos is not None:
            if np.isnan(my_gpos).any():
                raise ValueError("Some ``gpos`` element(s) is/are not-a-number (``nan``).")
            gpos += my_gpos
        if vtens is not None:
            if np.isnan(my_vtens).any():
                raise ValueError("Some ``vtens`` element(s) is/are not-a-number (``nan``).")
            vtens += my_vtens
        return self.energy

    def _internal_compute(self, gpos, vtens):
        raise NotImplementedError


class MicMecForceField(ForcePart):
    """A complete micromechanical force field: the sum of its parts (mmff.py:156-201)."""

    def __init__(self, system, parts):
        ForcePart.__init__(self, "all", system)
        self.system = system
        self.parts = []
        for part in parts:
            self.add_part(part)

    def add_part(self, part):
        self.parts.append(part)
        name = "part_%s" % part.name
        if name in self.__dict__:
            raise ValueError("The part %s occurs twice in the micromechanical force field." % name)
        self.__dict__[name] = part

    def update_rvecs(self, rvecs):
        ForcePart.update_rvecs(self, rvecs)
        self.system.domain.update_rvecs(rvecs)

    def update_pos(self, pos):
        ForcePart.update_pos(self, pos)
        self.system.pos[:] = pos

    def _internal_compute(self, gpos, vtens):
        return sum([part.compute(gpos, vtens) for part in self.parts])


class ForcePartMechanical(ForcePart):
    """The micromechanical part of the force field, evaluated on one B200.

    Parameters
    ----------
    system : System
        ``micmec_b200.system.System`` or the reference's ``micmec.system.System`` (only attributes are read).
    model : {"original", "default"}, optional
        Per-cell energy model: ``"original"`` is what the reference's ``mmff.py:27`` imports
        (``nanocell_original``: one averaged cell matrix); ``"default"`` is ``nanocell.py`` (eight corner
        matrices).  Defaults to ``$MICMEC_MODEL`` or ``"original"``.
    device : int, optional
        CUDA device ordinal.
    structured : bool, optional
        Full periodic grids built without index arrays (``System.periodic_grid``) can run on the structured-grid
        kernels (fused force + Verlet, no per-cell data in HBM).  ``None`` lets the library decide (grids of at
        least 4096 nodes), ``True`` / ``False`` force the choice; systems that do not qualify ignore it.
    """

    def __init__(self, system, model=None, device=0, structured=None, slab=None, scatter=False):
        ForcePart.__init__(self, "micmec", system)
        self.system = system
        # slab = (rank, count, nnodes_global): `system` is one z-slab of a periodic grid spread over `count` GPUs
        # (micmec_b200/slab.py); call init_comm() on every rank before the first compute
        self.slab = slab
        self.model = model or os.environ.get("MICMEC_MODEL", "original")
        if self.model not in _lib.MODELS:
            raise ValueError("Unknown per-cell model %r (expected 'original' or 'default')." % (self.model,))
        self.device = int(device)
        self.pbc = self.get_pbc(self.system.domain.rvecs)
        self._lib = _lib.load()
        self._handle = ctypes.c_void_p()
        self._keep = self._create()
        if structured is not None:
            _lib.check(self._lib.mm_set_option(self._handle, b"structured", int(bool(structured))))
        # kernel selection / tuning knobs, e.g. MICMEC_B200_MARCH2=0
        for opt in ("march2", "wrap_on_load", "tail", "tail_in_kernel", "plan"):
            env = os.environ.get("MICMEC_B200_" + opt.upper())
            if env is not None:
                _lib.check(self._lib.mm_set_option(self._handle, opt.encode(), int(env)))
        if scatter:
            # indexed kernels only: accumulate gpos with warp-aggregated fp64 atomics instead of the deterministic
            # per-cell-gradient + node-gather pair (DESIGN.md section 7 has the measured comparison)
            _lib.check(self._lib.mm_set_option(self._handle, b"scatter", 1))

    @staticmethod
    def get_pbc(rvecs):
        nper = rvecs.shape[0]
        pbc = nper > 0
        if pbc and nper != 3:  # mmff.py:249-254
            raise ValueError(
                "Attribute `rvecs` only supports finite systems or 3D periodic systems, "
                f"not {nper}D periodic systems."
            )
        return pbc

    def _create(self):
        system = self.system
        tab = type_tables(system.params, system.types)
        desc = _lib.Desc()
        desc.nnodes, desc.ncells = system.nnodes, system.ncells
        keep = [tab]
        shape = getattr(system, "structured_shape", None)
        if shape is not None and self.pbc and system.surrounding_nodes is None:
            desc.nx, desc.ny, desc.nz = (int(s) for s in shape)
            if self.slab is not None:
                desc.slab_rank, desc.slab_count, desc.nnodes_global = (int(v) for v in self.slab)
        else:
            sn = np.ascontiguousarray(system.surrounding_nodes, dtype=np.int64)
            sc = np.ascontiguousarray(system.surrounding_cells, dtype=np.int64)
            shift = np.ascontiguousarray(cell_shifts(system.grid, system.ncells, self.pbc))
            keep += [sn, sc, shift]
            desc.surrounding_nodes, desc.surrounding_cells, desc.shift = _lib.ptr(sn), _lib.ptr(sc), _lib.ptr(shift)
        desc.cell_type = _lib.ptr(tab["cell_type"])
        desc.ntypes = len(tab["type_nstates"])
        desc.type_nstates = _lib.ptr(tab["type_nstates"])
        desc.h0, desc.elasticity = _lib.ptr(tab["h0"]), _lib.ptr(tab["elasticity"])
        desc.free_energy, desc.effective_temp = _lib.ptr(tab["free_energy"]), _lib.ptr(tab["effective_temp"])
        desc.boltzmann = boltzmann
        desc.model = _lib.MODELS[self.model]
        desc.device = self.device
        _lib.check(self._lib.mm_create(ctypes.byref(desc), ctypes.byref(self._handle)))
        return keep

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle is not None and handle.value:
            self._lib.mm_destroy(handle)
            self._handle = ctypes.c_void_p()

    def init_comm(self, unique_id, nccl_path=None):
        """Join the NCCL communicator of the slab decomposition (``unique_id``: 128 bytes from rank 0, see slab.py)."""
        path = nccl_path.encode() if nccl_path else None
        _lib.check(self._lib.mm_comm_init(self._handle, path, bytes(unique_id)))

    @property
    def handle(self):
        return self._handle

    @property
    def structured(self):
        """True when the structured-grid (marching) kernels evaluate this part."""
        return int(self._lib.mm_get_option(self._handle, b"structured")) == 1

    @property
    def launches(self):
        """Number of CUDA kernels launched by this part so far."""
        return int(self._lib.mm_launch_count(self._handle))

    def _rvecs9(self):
        rv = np.zeros((3, 3))
        cur = np.asarray(self.system.domain.rvecs)
        rv[: cur.shape[0]] = cur
        return rv

    def _internal_compute(self, gpos, vtens):
        """mmff.py:288-297 on the device: H2D positions + cell, kernels, D2H energy / gpos / vtens."""
        pos = np.ascontiguousarray(self.system.pos, dtype=float)
        rv = self._rvecs9()
        _lib.check(self._lib.mm_set_rvecs(self._handle, _lib.ptr(rv)))
        _lib.check(self._lib.mm_set_pos(self._handle, _lib.ptr(pos), _lib.MM_HOST))
        energy = ctypes.c_double()
        _lib.check(self._lib.mm_compute(self._handle, ctypes.byref(energy), _lib.ptr(gpos), _lib.MM_HOST, _lib.ptr(vtens)))
        self._cells = None
        return energy.value

    def compute_device(self, pos, rvecs, gpos=None, vtens=False):
        """Zero-copy variant for torch CUDA tensors: ``pos`` [nnodes,3] float64 in, ``gpos`` (optional) out.

        Returns ``(energy, vtens or None)``; nothing but the ten result scalars crosses PCIe.
        """
        rv = np.zeros((3, 3))
        rvecs = np.asarray(rvecs, dtype=float)
        rv[: rvecs.shape[0]] = rvecs
        _lib.check(self._lib.mm_set_rvecs(self._handle, _lib.ptr(rv)))
        _lib.check(self._lib.mm_set_pos(self._handle, _lib.ptr(pos), _lib.MM_DEVICE))
        energy = ctypes.c_double()
        vt = np.zeros((3, 3)) if vtens else None
        _lib.check(self._lib.mm_compute(self._handle, ctypes.byref(energy), _lib.ptr(gpos), _lib.MM_DEVICE, _lib.ptr(vt)))
        return energy.value, vt

    # per-cell caches of the last evaluation (mmff.py:243-245, 290-292), fetched lazily from the device
    def _fetch_cells(self):
        if getattr(self, "_cells", None) is None:
            e = np.zeros(self.system.ncells)
            g = np.zeros((self.system.ncells, 8, 3))
            _lib.check(self._lib.mm_get_cell_cache(self._handle, _lib.ptr(e), _lib.ptr(g)))
            self._cells = (e, g)
        return self._cells

    epot_cells = property(lambda self: self._fetch_cells()[0])
    gpos_cells = property(lambda self: self._fetch_cells()[1])
