"""Degrees of freedom for the optimisers: drop-in for ``micmec.sampling.dof`` (dof.py:36-697).

A DOF object maps a flat vector ``x`` onto the geometry of the force field (positions and, for the cell DOFs, the
domain vectors), evaluates ``mmf.compute`` there - one GPU force evaluation through ``ForcePartMechanical`` - and
turns ``gpos`` / ``vtens`` into the gradient with respect to ``x``.  Convergence bookkeeping (``check_convergence``,
``conv_val``, ``conv_worst``, ``conv_count``, ``converged``) follows the reference so that ``QNOptimizer`` and user
scripts (``simulations/optimisation.py``, ``stress_strain.py``, ``relaxed_scan.py``) behave the same.
The batched counterparts for replica ensembles live in ``micmec_b200.sampling.batchopt``.
"""
import numpy as np

__all__ = ["DOF", "CartesianDOF", "BaseCellDOF", "FullCellDOF", "StrainCellDOF"]


def _norm_stats(rows):
    """(max, rms) of the Euclidean norms of the rows."""
    sq = (rows ** 2).sum(axis=1)
    return np.sqrt(sq.max()), np.sqrt(sq.mean()), sq


class DOF(object):
    def __init__(self, mmf):
        self.mmf = mmf
        self.x0 = None
        self._init_initial()
        self._gx = np.zeros(self.ndof, float)

    ndof = property(lambda self: len(self.x0))

    def _init_initial(self):
        raise NotImplementedError

    def _update(self, x):
        raise NotImplementedError

    def reset(self):
        self._update(self.x0)

    def check_delta(self, x=None, eps=1e-4, zero=None):
        """Finite-difference test of the gradient along 100 random small displacements (dof.py:60-68): returns the
        worst mismatch between f(x + dx) - f(x) and the trapezoid estimate 0.5 (g(x) + g(x + dx)) . dx, relative to the
        size of the terms of that dot product (0.5 |g(x) + g(x + dx)| . |dx|)."""
        x = self.x0 if x is None else x
        dxs = np.random.uniform(-eps, eps, (100, len(x)))
        if zero is not None:
            dxs[:, zero] = 0.0
        f0, g0 = self.fun(x, True)
        worst = 0.0
        for dx in dxs:
            f1, g1 = self.fun(x + dx, True)
            expected = 0.5 * np.dot(g0 + g1, dx)
            scale = 0.5 * np.dot(abs(g0 + g1), abs(dx))
            worst = max(worst, abs((f1 - f0) - expected) / max(scale, 1e-300))
        self._update(x)
        return worst

    def _finish_convergence(self, conv_vals):
        if len(conv_vals) == 0:
            raise RuntimeError("At least one convergence criterion must be present.")
        self.conv_val, self.conv_worst = max(conv_vals)
        self.conv_count = sum(int(v >= 1) for v, n in conv_vals)
        self.converged = self.conv_count == 0

    def _first_convergence_call(self):
        self.converged = False
        self.conv_val = 2
        self.conv_worst = "first_step"
        self.conv_count = -1

    def log(self):
        pass


class CartesianDOF(DOF):
    """Cartesian node coordinates (optionally of a selection); the domain is left alone (dof.py:75-193)."""

    def __init__(self, mmf, gpos_rms=1e-5, dpos_rms=1e-3, select=None):
        self.th_gpos_rms = gpos_rms
        self.th_dpos_rms = dpos_rms
        self.select = select
        DOF.__init__(self, mmf)
        self._last_pos = None

    def _pick(self, arr):
        return arr if self.select is None else arr[self.select]

    def _init_initial(self):
        pos = self.mmf.system.pos
        self.x0 = self._pick(pos).ravel().copy()
        self._pos = pos.copy()
        self._dpos = np.zeros(pos.shape, float)
        self._gpos = np.zeros(pos.shape, float)

    def _update(self, x):
        if self.select is None:
            self._pos[:] = x.reshape(-1, 3)
        else:
            self._pos[self.select] = x.reshape(-1, 3)
        self.mmf.update_pos(self._pos[:])

    def fun(self, x, do_gradient=False):
        self._update(x)
        if not do_gradient:
            return self.mmf.compute()
        self._gpos[:] = 0.0
        value = self.mmf.compute(self._gpos)
        self._gx[:] = self._pick(self._gpos).ravel()
        return value, self._gx.copy()

    def check_convergence(self):
        if self._last_pos is None:
            self._last_pos = self._pos.copy()
            return self._first_convergence_call()
        self.gpos_max, self.gpos_rms, _ = _norm_stats(self._pick(self._gpos))
        self._dpos[:] = self._pos
        self._dpos -= self._last_pos
        self.dpos_max, self.dpos_rms, _ = _norm_stats(self._pick(self._dpos))
        conv_vals = []
        if self.th_gpos_rms is not None:
            conv_vals.append((self.gpos_rms / self.th_gpos_rms, "gpos_rms"))
            conv_vals.append((self.gpos_max / (self.th_gpos_rms * 3), "gpos_max"))
        if self.th_dpos_rms is not None:
            conv_vals.append((self.dpos_rms / self.th_dpos_rms, "dpos_rms"))
            conv_vals.append((self.dpos_max / (self.th_dpos_rms * 3), "dpos_max"))
        self._finish_convergence(conv_vals)
        self._last_pos[:] = self._pos[:]


class BaseCellDOF(DOF):
    """Cell variables followed by fractional coordinates (dof.py:199-489).  ``x = [celldofs, frac.ravel()]``; with
    ``do_frozen`` the fractional coordinates stay at their initial values; ``freemask`` selects the free cell variables.
    The gradient needs the virial: d E / d rvecs = gvecs . vtens."""

    def __init__(self, mmf, gpos_rms=1e-5, dpos_rms=1e-3, grvecs_rms=1e-5, drvecs_rms=1e-3, do_frozen=False, freemask=None):
        if freemask is not None and not (isinstance(freemask, np.ndarray) and issubclass(freemask.dtype.type, np.bool_)
                                         and freemask.ndim == 1 and freemask.sum() > 0):
            raise TypeError("When given, freemask must be a vector of booleans.")
        self.th_gpos_rms, self.th_dpos_rms = gpos_rms, dpos_rms
        self.th_grvecs_rms, self.th_drvecs_rms = grvecs_rms, drvecs_rms
        self.do_frozen = do_frozen
        self.freemask = freemask
        DOF.__init__(self, mmf)
        self._last_pos = None
        self._last_rvecs = None

    ncellvar = property(lambda self: len(self.domainvars0))
    ncelldof = property(lambda self: len(self.domainvars0) if self.freemask is None else self.freemask.sum())

    def _reduce_cellvars(self, cellvars):
        return cellvars if self.freemask is None else cellvars[self.freemask]

    def _expand_celldofs(self, celldofs):
        if self.freemask is None:
            return celldofs
        cellvars = self.domainvars0.copy()
        cellvars[self.freemask] = celldofs
        return cellvars

    def _isfree(self, icellvar):
        return True if self.freemask is None else bool(self.freemask[icellvar])

    def _init_initial(self):
        system = self.mmf.system
        self.domainvars0 = self._get_initial_cellvars()
        if self.freemask is not None and len(self.freemask) != self.ncellvar:
            raise TypeError("The length of the freemask vector (%i) does not match the number of cellvars (%i)."
                            % (len(self.freemask), len(self.domainvars0)))
        celldofs0 = self._reduce_cellvars(self.domainvars0)
        frac = np.dot(system.pos, system.domain._get_gvecs(full=True).T)
        if self.do_frozen:
            self.x0 = celldofs0
            self._frac0 = frac
        else:
            self.x0 = np.concatenate([celldofs0, frac.ravel()])
        self._pos = system.pos.copy()
        self._dpos = np.zeros(system.pos.shape, float)
        self._gpos = np.zeros(system.pos.shape, float)
        self._rvecs = np.array(system.domain.rvecs)
        self._dcell = np.zeros(self._rvecs.shape, float)
        self._vtens = np.zeros((3, 3), float)
        self._grvecs = np.zeros(self._rvecs.shape, float)

    def _update(self, x):
        self._rvecs = self._cellvars_to_rvecs(self._expand_celldofs(x[:self.ncelldof]))
        self.mmf.update_rvecs(np.ascontiguousarray(self._rvecs))
        frac = self._frac0 if self.do_frozen else x[self.ncelldof:].reshape(-1, 3)
        self._pos[:] = np.dot(frac, self.mmf.system.domain._get_rvecs(full=True))
        self.mmf.update_pos(self._pos[:])

    def fun(self, x, do_gradient=False):
        self._update(x)
        if not do_gradient:
            return self.mmf.compute()
        self._gpos[:] = 0.0
        self._vtens[:] = 0.0
        value = self.mmf.compute(self._gpos, self._vtens)
        self._grvecs[:] = np.dot(self.mmf.system.domain.gvecs, self._vtens)
        jacobian = self._get_celldofs_jacobian(x[:self.ncelldof])
        assert jacobian.shape == (self._grvecs.size, self.ncelldof)
        self._gx[:self.ncelldof] = np.dot(self._grvecs.ravel(), jacobian)
        # keep only the part of the cell gradient the free cell variables can act on (used by check_convergence)
        u = np.linalg.svd(jacobian, full_matrices=False)[0]
        self._grvecs[:] = np.dot(u, np.dot(u.T, self._grvecs.ravel())).reshape(-1, 3)
        if not self.do_frozen:
            self._gx[self.ncelldof:] = np.dot(self._gpos, self._rvecs.T).ravel()
        return value, self._gx.copy()

    def check_convergence(self):
        if self._last_pos is None:
            self._last_pos = self._pos.copy()
            self._last_rvecs = self._rvecs.copy()
            return self._first_convergence_call()
        if not self.do_frozen:
            self.gpos_max, self.gpos_rms, gpossq = _norm_stats(self._gpos)
            self.gpos_indmax = gpossq.argmax()
            self._dpos[:] = self._pos
            self._dpos -= self._last_pos
        self.dpos_max, self.dpos_rms, _ = _norm_stats(self._dpos)
        self.grvecs_max, self.grvecs_rms, _ = _norm_stats(self._grvecs)
        self._dcell[:] = self._rvecs
        self._dcell -= self._last_rvecs
        self.drvecs_max, self.drvecs_rms, _ = _norm_stats(self._dcell)
        conv_vals = []
        if not self.do_frozen and self.th_gpos_rms is not None:
            conv_vals.append((self.gpos_rms / self.th_gpos_rms, "gpos_rms"))
            conv_vals.append((self.gpos_max / (self.th_gpos_rms * 3), "gpos_max(%i)" % self.gpos_indmax))
        if self.th_dpos_rms is not None:
            conv_vals.append((self.dpos_rms / self.th_dpos_rms, "dpos_rms"))
            conv_vals.append((self.dpos_max / (self.th_dpos_rms * 3), "dpos_max"))
        if self.th_grvecs_rms is not None:
            conv_vals.append((self.grvecs_rms / self.th_grvecs_rms, "grvecs_rms"))
            conv_vals.append((self.grvecs_max / (self.th_grvecs_rms * 3), "grvecs_max"))
        if self.th_drvecs_rms is not None:
            conv_vals.append((self.drvecs_rms / self.th_drvecs_rms, "drvecs_rms"))
            conv_vals.append((self.drvecs_max / (self.th_drvecs_rms * 3), "drvecs_max"))
        self._finish_convergence(conv_vals)
        self._last_pos[:] = self._pos[:]
        self._last_rvecs[:] = self._rvecs[:]

    def _get_initial_cellvars(self):
        raise NotImplementedError

    def _cellvars_to_rvecs(self, cellvars):
        raise NotImplementedError

    def _get_celldofs_jacobian(self, x):
        """d rvecs.ravel() / d celldofs: rows = cell vector components, columns = free cell variables."""
        raise NotImplementedError


class FullCellDOF(BaseCellDOF):
    """All components of the domain vectors, made dimensionless with volume^(1/nvec) (dof.py:492-519)."""

    def _get_initial_cellvars(self):
        cell = self.mmf.system.domain
        if cell.nvec == 0:
            raise ValueError("A cell optimization requires a system that is periodic.")
        self._rvecs_scale = cell.volume ** (1.0 / cell.nvec)
        return np.array(cell.rvecs).ravel() / self._rvecs_scale

    def _cellvars_to_rvecs(self, cellvars):
        return cellvars.reshape(-1, 3) * self._rvecs_scale

    def _get_celldofs_jacobian(self, x):
        jac = np.identity(self.ncellvar) * self._rvecs_scale
        return jac if self.freemask is None else jac[:, self.freemask]


# (row, column) of the symmetric deformation matrix each strain variable fills, by number of periodic directions;
# off-diagonal variables are twice the matrix element (dof.py:522-531)
_STRAIN_SLOTS = {
    3: [(0, 0), (1, 1), (2, 2), (1, 2), (2, 0), (0, 1)],
    2: [(0, 0), (1, 1), (0, 1)],
    1: [(0, 0)],
}


class StrainCellDOF(BaseCellDOF):
    """A symmetric deformation A applied to the initial cell, rvecs = A . rvecs0: cell rotations are eliminated and
    six variables [A00, A11, A22, 2 A12, 2 A20, 2 A01] remain in 3D (dof.py:522-697)."""

    def _get_initial_cellvars(self):
        cell = self.mmf.system.domain
        if cell.nvec == 0:
            raise ValueError("A cell optimization requires a system that is periodic.")
        self.rvecs0 = np.array(cell.rvecs)
        nvec = cell.nvec
        return np.array([1.0] * nvec + [0.0] * (nvec * (nvec - 1) // 2))

    def _deformation_basis(self, k):
        """d A / d (variable k) as an nvec x nvec matrix."""
        nvec = self.rvecs0.shape[0]
        i, j = _STRAIN_SLOTS[nvec][k]
        basis = np.zeros((nvec, nvec))
        if i == j:
            basis[i, i] = 1.0
        else:
            basis[i, j] = basis[j, i] = 0.5
        return basis

    def _cellvars_to_rvecs(self, x):
        nvec = self.rvecs0.shape[0]
        deform = np.zeros((nvec, nvec))
        for k, (i, j) in enumerate(_STRAIN_SLOTS[nvec]):
            if i == j:
                deform[i, i] = x[k]
            else:
                deform[i, j] = deform[j, i] = 0.5 * x[k]
        return np.dot(deform, self.rvecs0)

    def _get_celldofs_jacobian(self, x):
        nvar = len(_STRAIN_SLOTS[self.rvecs0.shape[0]])
        cols = [np.dot(self._deformation_basis(k), self.rvecs0).ravel() for k in range(nvar) if self._isfree(k)]
        return np.array(cols).T
