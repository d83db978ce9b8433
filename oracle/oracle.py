"""ctypes front-end of the CPU oracle (``oracle/liboracle.so``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product package never imports it.

The arithmetic lives in ``micmec_oracle.c`` (a literal restatement of the reference, pinned against
golden vectors recorded from the unmodified reference).  This file only marshals arrays and restates
the two pieces of *host bookkeeping* the reference does in Python:

* the minimum-image table ``mic`` (micmec/pes/mmff.py:259-286), restricted to the (vertex 0, vertex k)
  pairs ``deformation`` actually looks up (mmff.py:347-371) so that it is O(ncells) instead of O(nnodes^2);
* the parameter dictionaries keyed by integer type (mmff.py:219-231).
"""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "liboracle.so")

BOLTZMANN = 3.1668154051341965e-06  # molmod.boltzmann (see micmec_b200/units.py for provenance)
MAX_CHAIN = 16

MODELS = {"original": 0, "default": 1}


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(HERE, "micmec_oracle.c")
    hdr = os.path.join(HERE, "micmec_oracle.h")
    if (
        force
        or not os.path.exists(LIBPATH)
        or os.path.getmtime(LIBPATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    ):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return LIBPATH


class _System(ctypes.Structure):
    _fields_ = [
        ("nnodes", ctypes.c_int64),
        ("ncells", ctypes.c_int64),
        ("surrounding_nodes", ctypes.c_void_p),
        ("surrounding_cells", ctypes.c_void_p),
        ("shift", ctypes.c_void_p),
        ("cell_type", ctypes.c_void_p),
        ("ntypes", ctypes.c_int32),
        ("type_nstates", ctypes.c_void_p),
        ("type_offset", ctypes.c_void_p),
        ("h0", ctypes.c_void_p),
        ("C", ctypes.c_void_p),
        ("efree", ctypes.c_void_p),
        ("temp_eff", ctypes.c_void_p),
        ("boltzmann", ctypes.c_double),
        ("model", ctypes.c_int32),
        ("nthreads", ctypes.c_int32),
    ]


class _Chain(ctypes.Structure):
    _fields_ = [
        ("length", ctypes.c_int32),
        ("timestep", ctypes.c_double),
        ("temp", ctypes.c_double),
        ("timecon", ctypes.c_double),
        ("ndof", ctypes.c_double),
        ("pos", ctypes.c_double * MAX_CHAIN),
        ("vel", ctypes.c_double * MAX_CHAIN),
        ("masses", ctypes.c_double * MAX_CHAIN),
    ]


class _Baro(ctypes.Structure):
    _fields_ = [
        ("temp", ctypes.c_double),
        ("press", ctypes.c_double),
        ("timecon", ctypes.c_double),
        ("timestep", ctypes.c_double),
        ("mass_press", ctypes.c_double),
        ("anisotropic", ctypes.c_int32),
        ("vol_constraint", ctypes.c_int32),
        ("dim", ctypes.c_int32),
        ("baro_ndof", ctypes.c_int32),
        ("vel_press", ctypes.c_double * 9),
    ]


class _MD(ctypes.Structure):
    _fields_ = [
        ("nnodes", ctypes.c_int64),
        ("pos", ctypes.c_void_p),
        ("vel", ctypes.c_void_p),
        ("gpos", ctypes.c_void_p),
        ("masses", ctypes.c_void_p),
        ("posold", ctypes.c_void_p),
        ("delta", ctypes.c_void_p),
        ("rvecs", ctypes.c_double * 9),
        ("vtens", ctypes.c_double * 9),
        ("ptens", ctypes.c_double * 9),
        ("timestep", ctypes.c_double),
        ("time", ctypes.c_double),
        ("ndof", ctypes.c_double),
        ("counter", ctypes.c_int64),
        ("epot", ctypes.c_double),
        ("ekin", ctypes.c_double),
        ("temp", ctypes.c_double),
        ("etot", ctypes.c_double),
        ("econs", ctypes.c_double),
        ("cons_err", ctypes.c_double),
        ("press", ctypes.c_double),
        ("rmsd_gpos", ctypes.c_double),
        ("rmsd_delta", ctypes.c_double),
        ("ce_counter", ctypes.c_int64),
        ("ce_ekin_m", ctypes.c_double),
        ("ce_ekin_s", ctypes.c_double),
        ("ce_econs_m", ctypes.c_double),
        ("ce_econs_s", ctypes.c_double),
        ("has_thermo", ctypes.c_int32),
        ("has_baro", ctypes.c_int32),
        ("chain", _Chain),
        ("baro", _Baro),
        ("econs_correction", ctypes.c_double),
        ("work", ctypes.c_void_p),
        ("nforce", ctypes.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIBPATH)
        _lib.orc_compute.restype = ctypes.c_double
        _lib.orc_volume.restype = ctypes.c_double
        _lib.orc_work_size.restype = ctypes.c_int64
        _lib.orc_chain_econs.restype = ctypes.c_double
    return _lib


def _ptr(arr):
    return arr.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------------------------------------------
# host bookkeeping restated from the reference
# ----------------------------------------------------------------------------------------------------------------

_NEIGHBOR_CELLS = np.array(
    [(0, 0, 0), (-1, 0, 0), (0, -1, 0), (0, 0, -1), (-1, -1, 0), (-1, 0, -1), (0, -1, -1), (-1, -1, -1)]
)


def grid_nodes(grid, pbc=True):
    """Node grid coordinates in reference order (micmec/utils.py:113-137, fully periodic or fully open)."""
    grid = np.asarray(grid)
    nx, ny, nz = grid.shape
    per = bool(pbc)
    shape = (nx + 1 - per, ny + 1 - per, nz + 1 - per)
    valid = np.zeros(shape, dtype=bool)
    kk, ll, mm = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    for off in _NEIGHBOR_CELLS:
        ka, la, mu = kk + off[0], ll + off[1], mm + off[2]
        if per:
            # python negative indices wrap, exactly as grid[(kappa, lambda_, mu)] does at utils.py:135
            ok = np.ones(shape, dtype=bool)
            ka, la, mu = ka % nx, la % ny, mu % nz
        else:
            ok = (ka >= 0) & (ka < nx) & (la >= 0) & (la < ny) & (mu >= 0) & (mu < nz)
            ka, la, mu = np.clip(ka, 0, nx - 1), np.clip(la, 0, ny - 1), np.clip(mu, 0, nz - 1)
        valid |= ok & (grid[ka, la, mu] != 0)
    nodes = np.argwhere(valid)  # C order == the kk, ll, mm loop nest
    return nodes, shape


def cell_shifts(grid, surrounding_nodes, pbc=True):
    """``mic[v0, vk, :]`` for every cell and vertex, following micmec/pes/mmff.py:259-286 pair by pair."""
    sn = np.asarray(surrounding_nodes, dtype=np.int64)
    ncells = sn.shape[0]
    shift = np.zeros((ncells, 8, 3), dtype=np.int8)
    if not pbc:
        return shift  # mmff.py:262-263
    nodes, shape = grid_nodes(grid, True)
    maxs = np.array(shape) - 1  # kij_max, lij_max, mij_max (mmff.py:267-269)
    boundary = ((nodes == 0) | (nodes == maxs)).any(axis=1)  # utils.py:139-148
    p = np.repeat(sn[:, :1], 8, axis=1)  # vertex 0
    q = sn
    i = np.minimum(p, q)
    j = np.maximum(p, q)
    both = boundary[i] & boundary[j] & (p != q)
    kij = nodes[j] - nodes[i]  # (ncells, 8, 3)
    p_is_i = (p == i)[..., None]
    sign = 2 * (kij > 0) - 1
    hit = (np.abs(kij) == maxs) & both[..., None]
    general = np.where(p_is_i, -sign, sign)  # mic[i, j] = -sign, mic[j, i] = +sign
    special = np.where(p_is_i, kij == -1, kij == 1).astype(np.int64)  # 2-wide grids, mmff.py:278-285
    value = np.where(maxs != 1, general, special)
    shift[:] = np.where(hit, value, 0)
    return shift


def periodic_grid_system(shape, type_params):
    """Full periodic ``nx x ny x nz`` grid of one cell type at rest, NumPy only (conventions of ``build_system``,
    micmec/utils.py:164-263: node (k, l, m) has id ``(k*ny + l)*nz + m`` and sits at ``(k, l, m) * diag(h0)``,
    ``rvecs = shape * diag(h0)``, node mass = 8 * 1/8 of the type mass).  Returns the keyword arguments of
    ``Oracle`` plus ``pos``, ``masses``, ``rvecs``; used by the CPU arm of ``bench.py``, which must not load
    anything of the product library."""
    nx, ny, nz = (int(s) for s in shape)
    k, l, m = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    k, l, m = k.ravel(), l.ravel(), m.ravel()
    offsets = -_NEIGHBOR_CELLS  # vertex offsets (utils.py:43-52)

    def ident(a, b, c):
        return ((a % nx) * ny + (b % ny)) * nz + (c % nz)

    sn = np.stack([ident(k + d[0], l + d[1], m + d[2]) for d in offsets], axis=1).astype(np.int64)
    sc = np.stack([ident(k - d[0], l - d[1], m - d[2]) for d in offsets], axis=1).astype(np.int64)
    # image shifts: vertex offset bit along an axis where the cell sits on the last layer
    last = np.stack([k == nx - 1, l == ny - 1, m == nz - 1], axis=1)
    shift = (last[:, None, :] & (offsets[None, :, :] == 1)).astype(np.int8)
    h0 = np.asarray(type_params["cell"], dtype=float).reshape(-1, 3, 3)[0]
    diag = np.diag(h0)
    pos = np.stack([k * diag[0], l * diag[1], m * diag[2]], axis=1).astype(float)
    n = nx * ny * nz
    params = {"type1/" + key: type_params[key] for key in ("cell", "elasticity", "free_energy", "effective_temp", "mass")}
    arrays = dict(surrounding_nodes=sn, surrounding_cells=sc, shift=shift, grid=np.ones((nx, ny, nz), dtype=np.int64),
                  types=np.ones(n, dtype=np.int64), params=params, pbc=True)
    return arrays, pos, np.full(n, float(type_params["mass"])), np.diag(np.array([nx, ny, nz], dtype=float) * diag)


def type_tables(params, types):
    """Flatten ``system.params`` the way mmff.py:219-231 + :374-379 read it.  Returns compact arrays."""
    type_ids = sorted({int(t) for t in np.asarray(types).ravel()})
    index = {t: n for n, t in enumerate(type_ids)}
    nstates, offset, h0, C, efree, temp = [], [], [], [], [], []
    total = 0
    for t in type_ids:
        cell = np.asarray(params["type%d/cell" % t], dtype=float).reshape(-1, 3, 3)
        elas = np.asarray(params["type%d/elasticity" % t], dtype=float).reshape(-1, 3, 3, 3, 3)
        free = np.asarray(params["type%d/free_energy" % t], dtype=float).reshape(-1)
        ns = min(len(cell), len(elas), len(free))  # zip() semantics, mmff.py:377-379
        nstates.append(ns)
        offset.append(total)
        total += ns
        h0.append(cell[:ns])
        C.append(elas[:ns])
        efree.append(free[:ns])
        temp.append(float(params["type%d/effective_temp" % t]))
    cell_type = np.array([index[int(t)] for t in np.asarray(types).ravel()], dtype=np.int32)
    return dict(
        cell_type=cell_type,
        type_nstates=np.array(nstates, dtype=np.int32),
        type_offset=np.array(offset, dtype=np.int32),
        h0=np.ascontiguousarray(np.concatenate(h0)),
        C=np.ascontiguousarray(np.concatenate(C)),
        efree=np.ascontiguousarray(np.concatenate(efree)),
        temp_eff=np.array(temp, dtype=float),
    )


class Oracle(object):
    """CPU oracle for one system (anything with the attributes of ``micmec.system.System``)."""

    def __init__(self, system=None, model="original", nthreads=1, boltzmann=BOLTZMANN, **arrays):
        if system is not None:
            rv = np.asarray(system.domain.rvecs if hasattr(system, "domain") else system.rvecs)
            arrays = dict(
                surrounding_nodes=system.surrounding_nodes,
                surrounding_cells=system.surrounding_cells,
                grid=system.grid,
                types=system.types,
                params=system.params,
                pbc=rv.shape[0] > 0,
            )
        self.sn = np.ascontiguousarray(arrays["surrounding_nodes"], dtype=np.int64)
        self.sc = np.ascontiguousarray(arrays["surrounding_cells"], dtype=np.int64)
        self.nnodes = self.sc.shape[0]
        self.ncells = self.sn.shape[0]
        self.pbc = bool(arrays.get("pbc", True))
        if "shift" in arrays and arrays["shift"] is not None:
            self.shift = np.ascontiguousarray(arrays["shift"], dtype=np.int8)
        else:
            self.shift = np.ascontiguousarray(cell_shifts(arrays["grid"], self.sn, self.pbc))
        self.tab = type_tables(arrays["params"], arrays["types"])
        if int(self.tab["type_nstates"].max()) > MAX_CHAIN:
            raise ValueError("too many metastable states for the oracle")
        self.model = model
        self.c = _System(
            self.nnodes,
            self.ncells,
            _ptr(self.sn),
            _ptr(self.sc),
            _ptr(self.shift),
            _ptr(self.tab["cell_type"]),
            len(self.tab["type_nstates"]),
            _ptr(self.tab["type_nstates"]),
            _ptr(self.tab["type_offset"]),
            _ptr(self.tab["h0"]),
            _ptr(self.tab["C"]),
            _ptr(self.tab["efree"]),
            _ptr(self.tab["temp_eff"]),
            boltzmann,
            MODELS[model],
            int(nthreads),
        )
        self.boltzmann = boltzmann
        self.work = np.zeros(lib().orc_work_size(ctypes.byref(self.c)), dtype=float)

    # -- per-cell -------------------------------------------------------------------------------------------------
    @staticmethod
    def cell_state(model, verts, h0, C):
        verts = np.ascontiguousarray(verts, dtype=float)
        h0 = np.ascontiguousarray(h0, dtype=float)
        C = np.ascontiguousarray(C, dtype=float)
        e = ctypes.c_double()
        g = np.zeros((8, 3))
        lib().orc_cell_state(MODELS[model], _ptr(verts), _ptr(h0), _ptr(C), ctypes.byref(e), _ptr(g))
        return e.value, g

    def deformation(self, pos, rvecs):
        pos = np.ascontiguousarray(pos, dtype=float)
        rv = np.zeros((3, 3))
        rvecs = np.asarray(rvecs, dtype=float)
        rv[: rvecs.shape[0]] = rvecs
        e = np.zeros(self.ncells)
        g = np.zeros((self.ncells, 8, 3))
        v = np.zeros((self.ncells, 8, 3))
        lib().orc_deformation(ctypes.byref(self.c), _ptr(pos), _ptr(rv), _ptr(e), _ptr(g), _ptr(v))
        return e, g, v

    def compute(self, pos, rvecs, gpos=False, vtens=False):
        """Return ``(energy, gpos or None, vtens or None)`` - fresh arrays, not accumulated."""
        pos = np.ascontiguousarray(pos, dtype=float)
        rv = np.zeros((3, 3))
        rvecs = np.asarray(rvecs, dtype=float)
        rv[: rvecs.shape[0]] = rvecs
        g = np.zeros((self.nnodes, 3)) if gpos else None
        v = np.zeros((3, 3)) if vtens else None
        e = lib().orc_compute(
            ctypes.byref(self.c), _ptr(pos), _ptr(rv), _ptr(g) if gpos else None, _ptr(v) if vtens else None,
            _ptr(self.work),
        )
        return e, g, v

    # -- MD -------------------------------------------------------------------------------------------------------
    def md(self, pos, vel, masses, rvecs, timestep, **kwargs):
        return OracleMD(self, pos, vel, masses, rvecs, timestep, **kwargs)


class OracleMD(object):
    """Velocity Verlet (+NHC, +MTK) on the oracle; state lives in NumPy arrays owned by this object.

    ``thermo = dict(temp=, timecon=, chainlength=, chain_vel0=, chain_pos0=None)``;
    ``baro = dict(temp=, press=, timecon=, anisotropic=True, vol_constraint=False, vel_press0=)``.
    Random initial chain / barostat velocities are the caller's business (the reference draws them from
    the global legacy ``np.random`` state, nvt.py:402-408, sampling/utils.py:450-475).
    """

    def __init__(self, oracle, pos, vel, masses, rvecs, timestep, thermo=None, baro=None, ndof=None):
        self.oracle = oracle
        n = oracle.nnodes
        self.pos = np.array(pos, dtype=float, order="C")
        self.vel = np.array(vel, dtype=float, order="C")
        self.masses = np.array(masses, dtype=float, order="C")
        self.gpos = np.zeros((n, 3))
        self.posold = np.zeros((n, 3))
        self.delta = np.zeros((n, 3))
        md = _MD()
        md.nnodes = n
        md.pos, md.vel, md.gpos, md.masses = _ptr(self.pos), _ptr(self.vel), _ptr(self.gpos), _ptr(self.masses)
        md.posold, md.delta = _ptr(self.posold), _ptr(self.delta)
        md.rvecs[:] = list(np.asarray(rvecs, dtype=float).reshape(9))
        md.timestep = timestep
        md.time = 0.0
        md.counter = 0
        md.ndof = -1.0 if ndof is None else float(ndof)
        md.work = _ptr(oracle.work)
        md.has_thermo = int(thermo is not None)
        md.has_baro = int(baro is not None)
        if thermo is not None:
            length = int(thermo.get("chainlength", 3))
            md.chain.length = length
            md.chain.temp = thermo["temp"]
            md.chain.timecon = thermo["timecon"]
            md.chain.vel[:length] = list(np.asarray(thermo["chain_vel0"], dtype=float))
            if thermo.get("chain_pos0") is not None:
                md.chain.pos[:length] = list(np.asarray(thermo["chain_pos0"], dtype=float))
        if baro is not None:
            md.baro.temp = baro["temp"]
            md.baro.press = baro["press"]
            md.baro.timecon = baro["timecon"]
            md.baro.anisotropic = int(baro.get("anisotropic", True))
            md.baro.vol_constraint = int(baro.get("vol_constraint", False))
            md.baro.dim = 3
            ndof_b = 6 if md.baro.anisotropic else 1  # sampling/utils.py:478-501
            if md.baro.vol_constraint:
                ndof_b -= 1
            md.baro.baro_ndof = ndof_b
            vp = np.zeros(9)
            vp0 = np.asarray(baro["vel_press0"], dtype=float).reshape(-1)
            vp[: vp0.size] = vp0
            md.baro.vel_press[:] = list(vp)
        self.c = md
        lib().orc_md_initialize(ctypes.byref(oracle.c), ctypes.byref(md))

    def run(self, nsteps):
        for _ in range(nsteps):
            lib().orc_md_step(ctypes.byref(self.oracle.c), ctypes.byref(self.c))

    def __getattr__(self, name):
        c = self.__dict__.get("c")
        if c is not None and name in (
            "epot", "ekin", "temp", "etot", "econs", "cons_err", "press", "rmsd_gpos", "rmsd_delta", "time",
            "counter", "ndof", "nforce", "econs_correction",
        ):
            return getattr(c, name)
        raise AttributeError(name)

    @property
    def rvecs(self):
        return np.array(list(self.c.rvecs)).reshape(3, 3)

    @property
    def vtens(self):
        return np.array(list(self.c.vtens)).reshape(3, 3)

    @property
    def ptens(self):
        return np.array(list(self.c.ptens)).reshape(3, 3)

    @property
    def chain_vel(self):
        return np.array(list(self.c.chain.vel)[: self.c.chain.length])

    @property
    def chain_pos(self):
        return np.array(list(self.c.chain.pos)[: self.c.chain.length])

    @property
    def vel_press(self):
        vp = np.array(list(self.c.baro.vel_press)).reshape(3, 3)
        return vp if self.c.baro.anisotropic else vp[0, 0]
