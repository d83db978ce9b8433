"""Topology compiler: from a MicMec ``System`` to the flat arrays ``mm_create`` takes.

Replaces the O(nnodes^2) host work of ``ForcePartMechanical.__init__`` (micmec/pes/mmff.py:207-286): the dense
``mic[nnodes, nnodes, 3]`` table (1.6 TB at 64^3) becomes three wrap flags per cell, because the only pairs
``deformation`` looks up are (vertex 0, vertex k) of one cell (mmff.py:347-371) and for those
``mic[v0, vk, a] = d_ka * [kappa_a == n_a - 1]`` (d = vertex offset bit, kappa = cell grid coordinate).

Also the closed-form generator for full periodic N^3 grids (the reference's ``build_system``,
micmec/utils.py:164-263, does a linear search per vertex and cannot build the 64^3 / 256^3 configurations).
"""
import numpy as np

__all__ = ["cell_shifts", "type_tables", "periodic_grid_arrays", "NEIGHBOR_NODES"]

# micmec/utils.py:43-52
NEIGHBOR_NODES = np.array(
    [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)], dtype=np.int64
)


def cell_shifts(grid, ncells, pbc=True):
    """int8 [ncells][8][3]: the minimum-image integers of every (cell, vertex) pair."""
    shift = np.zeros((ncells, 8, 3), dtype=np.int8)
    if not pbc:
        return shift  # mmff.py:262-263
    grid = np.asarray(grid)
    kappa = np.argwhere(grid != 0)  # C order == the enumeration of micmec/utils.py:150-161
    if len(kappa) != ncells:
        raise ValueError("The grid has %d non-empty cells but the system has %d." % (len(kappa), ncells))
    wrap = kappa == (np.array(grid.shape) - 1)  # (ncells, 3)
    shift[:] = NEIGHBOR_NODES[None, :, :] * wrap[:, None, :]
    return shift


def type_tables(params, types):
    """Flatten ``system.params`` (keys ``typeN/cell|elasticity|free_energy|effective_temp``, mmff.py:219-231)."""
    flat = np.asarray(types).ravel()
    type_ids = sorted({int(t) for t in flat})  # int(type_), mmff.py:376 (types may be stored as floats)
    index = {t: n for n, t in enumerate(type_ids)}
    nstates, h0, C, efree, temp = [], [], [], [], []
    for t in type_ids:
        cell = np.asarray(params["type%d/cell" % t], dtype=float).reshape(-1, 3, 3)
        elas = np.asarray(params["type%d/elasticity" % t], dtype=float).reshape(-1, 3, 3, 3, 3)
        free = np.asarray(params["type%d/free_energy" % t], dtype=float).reshape(-1)
        ns = min(len(cell), len(elas), len(free))  # zip() at mmff.py:377-379
        nstates.append(ns)
        h0.append(cell[:ns])
        C.append(elas[:ns])
        efree.append(free[:ns])
        temp.append(float(params["type%d/effective_temp" % t]))
    lut = np.full(max(type_ids) + 1, -1, dtype=np.int32)
    for t, n in index.items():
        lut[t] = n
    cell_type = lut[flat.astype(np.int64)]
    return dict(
        cell_type=np.ascontiguousarray(cell_type, dtype=np.int32),
        type_nstates=np.array(nstates, dtype=np.int32),
        h0=np.ascontiguousarray(np.concatenate(h0)),
        elasticity=np.ascontiguousarray(np.concatenate(C)),
        free_energy=np.ascontiguousarray(np.concatenate(efree)),
        effective_temp=np.array(temp, dtype=float),
    )


def periodic_grid_arrays(shape):
    """Index arrays of a full periodic grid in reference order, id = (k*ny + l)*nz + m  (O(N^3), vectorised).

    Returns ``surrounding_nodes`` [ncells][8], ``surrounding_cells`` [nnodes][8] and ``boundary_nodes``
    exactly as micmec/utils.py:205-233 would produce them.
    """
    nx, ny, nz = (int(s) for s in shape)
    k, l, m = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    k, l, m = k.ravel(), l.ravel(), m.ravel()

    def ident(a, b, c):
        return ((a % nx) * ny + (b % ny)) * nz + (c % nz)

    sn = np.stack([ident(k + d[0], l + d[1], m + d[2]) for d in NEIGHBOR_NODES], axis=1)
    sc = np.stack([ident(k - d[0], l - d[1], m - d[2]) for d in NEIGHBOR_NODES], axis=1)
    boundary = np.nonzero((k == 0) | (k == nx - 1) | (l == 0) | (l == ny - 1) | (m == 0) | (m == nz - 1))[0]
    return sn.astype(np.int64), sc.astype(np.int64), boundary.astype(np.int64)
