"""z-slab decomposition of a full periodic grid over the GPUs of one box (one process per GPU).

The reference is single-process; this module is the host-side plumbing of the decomposition the
library implements in ``csrc/mm_comm.cu``:

* ``SlabLayout``   which z planes a rank owns, and the maps between global reference order
                   ``id = (k*ny + l)*nz + m`` and the rank-local order ``(k*ny + l)*nzl + (m - m0)``;
* ``local_system`` the rank's ``System`` (a grid of shape (nx, ny, nzl) whose domain vectors are the GLOBAL ones);
* ``init_comm``    NCCL bootstrap: rank 0 draws the unique id, ``torch.distributed`` broadcasts it, every rank joins;
* ``gather_nodes`` collect a per-node array of all slabs on rank 0 in global reference order (trajectory output).

``torch.distributed`` is only the bootstrap / gather plumbing (any backend: NCCL on GPUs, gloo in the CPU tests);
the per-step halo planes and the 16-double all-reduce travel inside ``mm_md_run`` / ``mm_compute``.
"""
import ctypes

import numpy as np

from . import _lib
from .system import System

__all__ = ["SlabLayout", "local_system", "init_comm", "gather_nodes", "nccl_library_path"]


class SlabLayout(object):
    """Planes ``[m0, m1)`` of an ``nx x ny x nz`` grid owned by ``rank`` out of ``count`` equal slabs."""

    def __init__(self, shape, rank, count):
        self.shape = tuple(int(s) for s in shape)
        self.rank, self.count = int(rank), int(count)
        nx, ny, nz = self.shape
        if count < 1 or not 0 <= rank < count:
            raise ValueError("rank %d out of range for %d slabs" % (rank, count))
        if nz % count != 0:
            raise ValueError("the number of z planes (%d) must be a multiple of the number of slabs (%d)" % (nz, count))
        self.nzl = nz // count
        if self.nzl < 2 and count > 1:
            raise ValueError("a slab needs at least 2 planes")
        self.m0, self.m1 = self.rank * self.nzl, (self.rank + 1) * self.nzl
        self.local_shape = (nx, ny, self.nzl)
        self.nnodes_local = nx * ny * self.nzl
        self.nnodes_global = nx * ny * nz
        self.up, self.down = (self.rank + 1) % count, (self.rank - 1) % count

    def global_ids(self):
        """Global reference ids of the local nodes, in local order."""
        nx, ny, nz = self.shape
        k, l, m = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(self.m0, self.m1), indexing="ij")
        return ((k * ny + l) * nz + m).ravel()

    def take(self, array):
        """Local part of a global per-node (or per-cell) array."""
        return np.ascontiguousarray(np.asarray(array)[self.global_ids()])

    def slab_arg(self):
        """The ``slab=`` argument of ``ForcePartMechanical``."""
        return (self.rank, self.count, self.nnodes_global)


def local_system(layout, type_params, pos=None, masses=None):
    """The rank's slab of a one-type periodic grid.  ``pos`` / ``masses``: GLOBAL arrays to cut from, or ``None`` for
    the rest lattice generated directly in slab form (no rank ever holds the whole 256^3 grid)."""
    nx, ny, nz = layout.shape
    h0 = np.asarray(type_params["cell"], dtype=float).reshape(-1, 3, 3)[0]
    diag = np.diag(h0)
    if pos is None:
        k, l, m = np.meshgrid(np.arange(nx, dtype=float), np.arange(ny, dtype=float),
                              np.arange(layout.m0, layout.m1, dtype=float), indexing="ij")
        lpos = np.stack([k.ravel() * diag[0], l.ravel() * diag[1], m.ravel() * diag[2]], axis=1)
    else:
        lpos = layout.take(pos)
    lmass = np.full(layout.nnodes_local, float(type_params["mass"])) if masses is None else layout.take(masses)
    rvecs = np.diag(np.array([nx, ny, nz], dtype=float) * diag)  # GLOBAL domain
    params = {"type1/" + key: type_params[key] for key in ("cell", "elasticity", "free_energy", "effective_temp", "mass")}
    return System(lpos, lmass, rvecs, None, None, grid=np.ones(layout.local_shape, dtype=np.int64),
                  types=np.ones(layout.nnodes_local, dtype=np.int64), params=params, structured_shape=layout.local_shape)


def nccl_library_path():
    """The libnccl.so.2 PyTorch ships (the one already mapped into the process once torch.cuda.nccl is used)."""
    import os
    try:
        import nvidia.nccl as pkg

        path = os.path.join(os.path.dirname(pkg.__file__), "lib", "libnccl.so.2")
        if os.path.exists(path):
            return path
    except Exception:
        pass
    return None


def init_comm(part, layout, group=None):
    """Create the slab communicator of ``part`` (a ``ForcePartMechanical`` built with ``slab=layout.slab_arg()``)."""
    import torch.distributed as dist

    if layout.count == 1:
        return
    path = nccl_library_path()
    ident = [None]
    if dist.get_rank(group) == 0:
        buf = ctypes.create_string_buffer(128)
        _lib.check(_lib.load().mm_comm_unique_id(path.encode() if path else None, buf))
        ident[0] = bytes(buf.raw)
    dist.broadcast_object_list(ident, src=0, group=group)
    part.init_comm(ident[0], path)


def gather_nodes(layout, local, group=None, dst=0):
    """Per-node array of all slabs in GLOBAL reference order on rank ``dst`` (``None`` elsewhere)."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local)
    if layout.count == 1:
        return local.copy()
    tensor = torch.from_numpy(local)
    if dist.get_backend(group) == "nccl":
        tensor = tensor.cuda()
    rank = dist.get_rank(group)
    pieces = [torch.empty_like(tensor) for _ in range(layout.count)] if rank == dst else None
    dist.gather(tensor, pieces, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.empty((layout.nnodes_global,) + local.shape[1:], dtype=local.dtype)
    for r, piece in enumerate(pieces):
        out[SlabLayout(layout.shape, r, layout.count).global_ids()] = piece.cpu().numpy()
    return out
