"""Helpers shared by the tests: load golden fixtures and rebuild system descriptions from them."""
import glob
import os
import types

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def force_fixtures():
    return sorted(os.path.basename(p)[len("force_"):-4] for p in glob.glob(os.path.join(GOLDEN, "force_*.npz")))


def traj_fixtures():
    return sorted(os.path.basename(p)[len("traj_"):-4] for p in glob.glob(os.path.join(GOLDEN, "traj_*.npz")))


class Rec(types.SimpleNamespace):
    """Plain attribute bag shaped like ``micmec.system.System`` (pos, masses, rvecs, topology, params)."""


def system_from(d, prefix=""):
    params = {}
    for key, val in d.items():
        if key.startswith(prefix + "params:"):
            name = key[len(prefix) + len("params:"):]
            params[name] = float(val) if val.ndim == 0 else val
    g = lambda k: d[prefix + k]  # noqa: E731
    return Rec(
        pos=g("pos").copy(), masses=g("masses").copy(), rvecs=g("rvecs").copy(),
        surrounding_cells=g("surrounding_cells"), surrounding_nodes=g("surrounding_nodes"),
        boundary_nodes=g("boundary_nodes"), grid=g("grid"), types=g("types"), params=params,
        nnodes=g("surrounding_cells").shape[0], ncells=g("surrounding_nodes").shape[0],
    )


def rel_rms(a, b):
    """max |a-b| relative to rms(b) - the north_star's force/virial metric."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    rms = np.sqrt(np.mean(b * b))
    if rms == 0.0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b)) / rms)


def virial_noise(gpos_cells, verts_cells):
    """Rounding-noise floor of the reference's own virial.  It sums g_v (x) r_v over ABSOLUTE vertex positions
    (micmec/pes/mmff.py:320-323): terms of size |g||r| that cancel down to |g||h0|, and each g_v already carries
    a relative error of eps/|strain| from the cancellation in 0.5 (G G^T - I).  1e-12 of the absolute-term sum
    (~4500 eps) bounds what a different but equally valid evaluation order may change."""
    g = np.abs(np.asarray(gpos_cells)).reshape(-1, 3)
    r = np.abs(np.asarray(verts_cells)).reshape(-1, 3)
    return float(1e-12 * np.max(g.T @ r))


def virial_close(v, vref, tol, noise):
    v, vref = np.asarray(v), np.asarray(vref)
    rms = np.sqrt(np.mean(vref * vref))
    return bool(np.max(np.abs(v - vref)) <= tol * rms + noise)
