"""Lockstep quasi-Newton optimisation of a batch of independent replicas (BASELINE.json config 5: defect
configurations x perturbations).

The reference optimises one small system per Python process: ``QNOptimizer`` (opt.py:258-393) over a ``CartesianDOF``
or ``StrainCellDOF`` (dof.py:75-193, 522-697), one force evaluation and one dense ``eigh`` per trial step.  Here R
replicas advance together: every iteration is ONE batched force evaluation on the GPU (``ReplicaBatch.compute``: all
replicas in one kernel sequence) followed by the per-replica trust-radius algebra in batched array form - SR1 update,
spectrum, ridge search, accept / shrink - with each replica keeping its own Hessian model, trust radius and
convergence state, exactly the state machine of the reference.  A replica that has converged (or given up) is frozen;
its slot is still evaluated but no longer moves.  Replicas never communicate: on several GPUs each rank optimises its
own share (``shard``).
"""
import numpy as np

__all__ = ["ReplicaQNOptimizer", "DeviceReplicaQNOptimizer", "solve_trust_radius_batch", "shard"]

_STRAIN_SLOTS = [(0, 0), (1, 1), (2, 2), (1, 2), (2, 0), (0, 1)]  # dof.py:522-531


def shard(items, rank, world):
    """The contiguous share of ``items`` rank ``rank`` of ``world`` works on (replicas only: no communication)."""
    n = len(items)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    return items[lo:hi]


def _norm(a):
    return np.sqrt((a * a).sum(axis=-1))


def _matvec(mats, vecs, transpose=False):
    """Batched matrix-vector products [R, n, n] x [R, n] -> [R, n] through BLAS (about 1.5x faster than einsum)."""
    if transpose:
        return np.matmul(vecs[:, None, :], mats)[:, 0, :]
    return np.matmul(mats, vecs[:, :, None])[:, :, 0]


def _eigh(h, device=0, backend="device"):
    """Batched symmetric eigen-decomposition [R, n, n] -> (evals [R, n] ascending, evecs [R, n, n], columns).

    ``backend="device"`` (the product path): ``mm_batched_eigh`` on the B200 - one thread block per matrix, cyclic Jacobi
    in shared memory.  The 10 240 81 x 81 Hessian models of config 5 per sweep are LAPACK-bound on the host (~0.7 ms
    each) and ``torch.linalg.eigh`` loops cuSOLVER's syevd over the batch for n > 32, which is no faster.  It fails
    loudly without a usable device.  ``backend="lapack"`` must be asked for explicitly; the CPU tests of the host logic
    do."""
    if backend == "lapack":
        return np.linalg.eigh(h)
    if backend != "device":
        raise ValueError("eigh backend must be 'device' or 'lapack'")
    import ctypes

    from .. import _lib

    lib = _lib.load()
    h = np.ascontiguousarray(h, dtype=float)
    evals = np.empty(h.shape[:2])
    evecs = np.empty(h.shape)
    worst = ctypes.c_int32()
    _lib.check(lib.mm_batched_eigh(int(device), h.shape[0], h.shape[1], _lib.ptr(h), _lib.MM_HOST, _lib.ptr(evals),
                                   _lib.ptr(evecs), ctypes.byref(worst)))
    if worst.value >= 30:
        raise RuntimeError("mm_batched_eigh: a Hessian model did not converge in 30 Jacobi sweeps")
    return evals, evecs


def solve_trust_radius_batch(grad, evals, radius, threshold=1e-5, maxiter=20000):
    """``solve_trust_radius`` (opt.py:396-443) for R problems at once: grad, evals [R, n], radius [R] -> step [R, n].
    The Newton step where it fits inside the radius, otherwise -grad / (evals + ridge) with the ridge from the same
    bracketing and false-position iteration as the reference, run under masks until every problem has met the
    threshold (the reference needs several hundred iterations for some spectra; ``maxiter`` only guards against a
    zero gradient on an indefinite model, where the reference would loop forever)."""
    nrep = len(radius)
    out = np.zeros_like(grad)
    done = np.zeros(nrep, bool)
    posdef = np.flatnonzero(evals.min(axis=1) > 0)
    if posdef.size:
        newton = -grad[posdef] / evals[posdef]
        fits = _norm(newton) <= radius[posdef]
        out[posdef[fits]] = newton[fits]
        done[posdef[fits]] = True
    todo = np.flatnonzero(~done)
    if todo.size == 0:
        return out
    g, e, r = grad[todo], evals[todo], radius[todo]
    ridge_min = -e.min(axis=1)

    def excess(ridge, sel):
        with np.errstate(divide="ignore", invalid="ignore"):
            return _norm(g[sel] / (e[sel] + ridge[:, None])) - r[sel]

    def bracket(alpha, sign):
        ridge, err = np.zeros(len(r)), np.zeros(len(r))
        alpha = alpha.copy()
        pending = np.arange(len(r))
        for _ in range(maxiter):
            trial = ridge_min[pending] + alpha[pending]
            val = excess(trial, pending)
            ridge[pending], err[pending] = trial, val
            pending = pending[~(sign * val < 0)]
            if pending.size == 0:
                break
            alpha[pending] = alpha[pending] * 2 if sign > 0 else alpha[pending] / 2
        return ridge, err

    a0, a1 = bracket(np.minimum(1e1, abs(e.max(axis=1))), -1)
    b0, b1 = bracket(np.maximum(1e-5, abs(ridge_min)), 1)
    ridge = np.zeros(len(r))
    pending = np.arange(len(r))
    for _ in range(maxiter):
        p = pending
        trial = (a1[p] * a0[p] - b1[p] * b0[p]) / (a1[p] - b1[p])
        val = excess(trial, p)
        ridge[p] = trial
        to_a = (val > 0) & (a1[p] > 0)
        a0[p[to_a]], a1[p[to_a]] = trial[to_a], val[to_a]
        b0[p[~to_a]], b1[p[~to_a]] = trial[~to_a], val[~to_a]
        pending = p[abs(val) > r[p] * threshold]
        if pending.size == 0:
            break
    out[todo] = -g / (e + ridge[:, None])
    return out


class ReplicaQNOptimizer(object):
    """Trust-radius SR1 quasi-Newton optimisation of R replicas in lockstep.

    Parameters
    ----------
    batch : object with ``compute(pos [R,n,3], rvecs [R,3,3], gpos=True, vtens=bool) -> (energies, gpos, vtens)``
        ``micmec_b200.replicas.ReplicaBatch`` (one GPU kernel sequence for all replicas).
    pos0, rvecs0 : arrays [R, n, 3], [R, 3, 3]
        Starting geometries.
    dof : {"cartesian", "strain"}
        ``"cartesian"``: node coordinates, fixed cell (``CartesianDOF``).  ``"strain"``: six symmetric-deformation
        variables of the cell plus fractional coordinates (``StrainCellDOF``; needs the virial).
    gpos_rms, dpos_rms, grvecs_rms, drvecs_rms : convergence thresholds with the meaning (and the automatic 3x max
        thresholds) of dof.py:133-193, 382-449.
    trust_radius, small_radius, too_small_radius : as ``QNOptimizer`` (opt.py:262-300).
    device : CUDA device ordinal for the batched eigen-decompositions (the one ``batch`` lives on).
    eigh : "device" (``mm_batched_eigh``; fails without a GPU) or "lapack" (host; only for tests of the host logic).

    Attributes after / during ``run``: ``x``, ``f``, ``g`` (current accepted point), ``pos``, ``rvecs``, ``converged``,
    ``failed`` (trust radius underflow: the reference raises RuntimeError there), ``iterations`` (accepted steps per
    replica), ``trust_radius``, ``conv_val``, ``conv_count``, ``evaluations`` (batched force calls so far).
    """

    def __init__(self, batch, pos0, rvecs0, dof="cartesian", gpos_rms=1e-5, dpos_rms=1e-3, grvecs_rms=1e-5,
                 drvecs_rms=1e-3, trust_radius=1.0, small_radius=1e-5, too_small_radius=1e-10, device=0, eigh="device"):
        if dof not in ("cartesian", "strain"):
            raise ValueError("dof must be 'cartesian' or 'strain'")
        self.batch = batch
        self.device = device
        self.eigh_backend = eigh
        self.kind = dof
        pos0 = np.array(pos0, dtype=float)
        self.rvecs0 = np.array(rvecs0, dtype=float)
        self.nrep, self.nnodes = pos0.shape[0], pos0.shape[1]
        self.th = dict(gpos_rms=gpos_rms, dpos_rms=dpos_rms, grvecs_rms=grvecs_rms, drvecs_rms=drvecs_rms)
        self.initial_trust_radius = float(trust_radius)
        self.small_radius, self.too_small_radius = small_radius, too_small_radius
        self.evaluations = 0
        if dof == "cartesian":
            self.ncell = 0
            x0 = pos0.reshape(self.nrep, -1).copy()
        else:
            self.ncell = 6
            frac = np.einsum("rni,rji->rnj", pos0, self._gvecs(self.rvecs0))
            cell0 = np.tile(np.array([1.0, 1.0, 1.0, 0.0, 0.0, 0.0]), (self.nrep, 1))
            x0 = np.concatenate([cell0, frac.reshape(self.nrep, -1)], axis=1)
            # d rvecs.ravel() / d strain variable k = (basis_k . rvecs0).ravel()   (dof.py:582-697)
            basis = np.zeros((6, 3, 3))
            for k, (i, j) in enumerate(_STRAIN_SLOTS):
                if i == j:
                    basis[k, i, i] = 1.0
                else:
                    basis[k, i, j] = basis[k, j, i] = 0.5
            self._basis = basis
            self._jac = np.einsum("kij,rjl->rkil", basis, self.rvecs0).reshape(self.nrep, 6, 9).transpose(0, 2, 1)
            u = np.linalg.svd(self._jac, full_matrices=False)[0]
            self._proj = u @ u.transpose(0, 2, 1)
        self.ndof = x0.shape[1]
        n = self.nrep
        self.hessian = np.tile(np.identity(self.ndof), (n, 1, 1))
        self.trust_radius = np.full(n, self.initial_trust_radius)
        self.converged = np.zeros(n, bool)
        self.failed = np.zeros(n, bool)
        self.iterations = np.zeros(n, dtype=np.int64)
        self.conv_val = np.full(n, 2.0)
        self.conv_count = np.full(n, -1, dtype=np.int64)
        # QNOptimizer.initialize: evaluate x0, then the first trial step (opt.py:317-320)
        self.x_old = x0
        self.f_old, self.g_old, aux = self._fun(x0)
        self.x, self.f, self.g = x0.copy(), self.f_old.copy(), self.g_old.copy()
        self._aux = aux
        self._last = None
        self._fresh = np.ones(n, bool)       # needs a new spectrum before its next trial
        self._accepted = np.zeros(n, bool)   # accepted a step in the current sweep
        self._evals = np.zeros((n, self.ndof))
        self._evecs = np.zeros((n, self.ndof, self.ndof))
        self._grad_eigen = np.zeros((n, self.ndof))
        self._started = np.zeros(n, bool)    # first check_convergence call done
        self._rank1 = None                   # scratch for the rank-one terms of the SR1 update

    # ---- DOF mappings ----------------------------------------------------------------------------------------------
    @staticmethod
    def _gvecs(rvecs):
        return np.linalg.inv(rvecs).transpose(0, 2, 1)

    def _geometry(self, x):
        if self.kind == "cartesian":
            return x.reshape(self.nrep, self.nnodes, 3), self.rvecs0
        s = x[:, :6]
        deform = np.einsum("rk,kij->rij", s, self._basis)  # A_ii = s_i, A_ij = s_k / 2 (dof.py:553-566)
        rvecs = deform @ self.rvecs0
        frac = x[:, 6:].reshape(self.nrep, self.nnodes, 3)
        return frac @ rvecs, rvecs

    def _fun(self, x):
        """Energies [R], gradients with respect to x [R, ndof] and (pos, rvecs, gpos, grvecs) of this evaluation."""
        pos, rvecs = self._geometry(x)
        strain = self.kind == "strain"
        energies, gpos, vtens = self.batch.compute(pos, rvecs, gpos=True, vtens=strain)
        self.evaluations += 1
        if not strain:
            return energies, gpos.reshape(self.nrep, -1).copy(), (pos.copy(), rvecs, gpos, None)
        grvecs = self._gvecs(rvecs) @ vtens
        gx = np.empty((self.nrep, self.ndof))
        gx[:, :6] = np.einsum("ri,rik->rk", grvecs.reshape(self.nrep, 9), self._jac)
        gx[:, 6:] = (gpos @ rvecs.transpose(0, 2, 1)).reshape(self.nrep, -1)
        grvecs = np.einsum("rij,rj->ri", self._proj, grvecs.reshape(self.nrep, 9)).reshape(self.nrep, 3, 3)
        return energies, gx, (pos.copy(), rvecs.copy(), gpos, grvecs)

    # ---- convergence (dof.py:155-193, 382-449) ------------------------------------------------------------------
    def _check_convergence(self, sel):
        pos, rvecs, gpos, grvecs = self._aux
        if self._last is None:
            self._last = [pos.copy(), np.array(rvecs, dtype=float).copy()]
        first = sel & ~self._started
        self._last[0][first], self._last[1][first] = pos[first], rvecs[first]
        self._started |= first
        idx = np.flatnonzero(sel & ~first)
        if idx.size == 0:
            return

        def stats(rows):
            sq = (rows ** 2).sum(axis=-1)
            return np.sqrt(sq.max(axis=-1)), np.sqrt(sq.mean(axis=-1))

        ratios = []
        gmax, grms = stats(gpos[idx])
        dmax, drms = stats(pos[idx] - self._last[0][idx])
        th = self.th
        if th["gpos_rms"] is not None:
            ratios += [grms / th["gpos_rms"], gmax / (th["gpos_rms"] * 3)]
        if th["dpos_rms"] is not None:
            ratios += [drms / th["dpos_rms"], dmax / (th["dpos_rms"] * 3)]
        if self.kind == "strain":
            cmax, crms = stats(grvecs[idx])
            emax, erms = stats(rvecs[idx] - self._last[1][idx])
            if th["grvecs_rms"] is not None:
                ratios += [crms / th["grvecs_rms"], cmax / (th["grvecs_rms"] * 3)]
            if th["drvecs_rms"] is not None:
                ratios += [erms / th["drvecs_rms"], emax / (th["drvecs_rms"] * 3)]
        if not ratios:
            raise RuntimeError("At least one convergence criterion must be present.")
        ratios = np.array(ratios)
        self.conv_val[idx] = ratios.max(axis=0)
        self.conv_count[idx] = (ratios >= 1).sum(axis=0)
        self.converged[idx] = self.conv_count[idx] == 0
        self._last[0][idx], self._last[1][idx] = pos[idx], rvecs[idx]

    # ---- the state machine of QNOptimizer.propagate / make_step (opt.py:322-393) ---------------------------------------
    def _refresh_models(self):
        """SR1 update and new spectrum for the replicas that accepted a step last sweep.  Written to touch the
        [R, ndof, ndof] arrays as few times as possible (they are 0.5 GB each for config 5): the rank-one term is
        formed for all replicas with a zero coefficient where no update applies, and subsets are only gathered when
        they are proper subsets."""
        upd_mask = self._fresh & (self.iterations > 0)
        if upd_mask.any():
            dx = np.where(upd_mask[:, None], self.x - self.x_old, 0.0)
            dg = np.where(upd_mask[:, None], self.g - self.g_old, 0.0)
            resid = dg - _matvec(self.hessian, dx)
            denom = (resid * dx).sum(axis=1)
            safe = upd_mask & (abs(denom) > 1e-5 * _norm(dx) * _norm(resid))
            with np.errstate(divide="ignore", invalid="ignore"):
                coef = np.where(safe, 1.0 / np.where(safe, denom, 1.0), 0.0)
            if self._rank1 is None:
                self._rank1 = np.empty_like(self.hessian)
            np.multiply(resid[:, :, None], (coef[:, None] * resid)[:, None, :], out=self._rank1)
            self.hessian += self._rank1
            bad = np.flatnonzero(upd_mask & ~safe)
            if bad.size:  # a failed update poisons the model: identity and the initial radius again (opt.py:325-329)
                self.hessian[bad] = np.identity(self.ndof)
                self.trust_radius[bad] = self.initial_trust_radius
            upd = np.flatnonzero(upd_mask)
            self.x_old[upd], self.f_old[upd], self.g_old[upd] = self.x[upd], self.f[upd], self.g[upd]
        new_mask = self._fresh & ~(self.converged | self.failed)
        if new_mask.all():
            self._evals, self._evecs = _eigh(self.hessian, self.device, self.eigh_backend)
            self._grad_eigen = _matvec(self._evecs, self.g_old, transpose=True)
        elif new_mask.any():
            new = np.flatnonzero(new_mask)
            evals, evecs = _eigh(self.hessian[new], self.device, self.eigh_backend)
            self._evals[new], self._evecs[new] = evals, evecs
            self._grad_eigen[new] = _matvec(evecs, self.g_old[new], transpose=True)
        self._fresh[new_mask] = False

    def sweep(self):
        """One batched force evaluation: every live replica tries a step and either accepts it or shrinks its radius."""
        live = ~(self.converged | self.failed)
        self._refresh_models()
        delta = np.zeros_like(self.x)
        act = np.flatnonzero(live)
        if act.size:
            delta[act] = solve_trust_radius_batch(self._grad_eigen[act], self._evals[act], self.trust_radius[act])
        radius = _norm(delta)
        trial = self.x_old + _matvec(self._evecs, delta)
        trial[~live] = self.x[~live]
        f, g, aux = self._fun(trial)
        shrink = f - self.f_old > 0
        shrink |= (self.trust_radius < self.small_radius) & (_norm(g) - _norm(self.g_old) > 0)
        shrink &= live
        accept = live & ~shrink
        # rejected: halve until strictly inside the step that was just tried (opt.py:376-382)
        idx = np.flatnonzero(shrink)
        if idx.size:
            tr = self.trust_radius[idx] * 0.5
            while True:
                more = tr >= radius[idx]
                if not more.any():
                    break
                tr[more] *= 0.5
            self.trust_radius[idx] = tr
            self.failed[idx[tr < self.too_small_radius]] = True
        grow = accept & (self.trust_radius < self.initial_trust_radius)
        self.trust_radius[grow] *= 2.0
        self.x[accept], self.f[accept], self.g[accept] = trial[accept], f[accept], g[accept]
        if self._aux[3] is None:
            merged = (np.where(accept[:, None, None], aux[0], self._aux[0]), aux[1],
                      np.where(accept[:, None, None], aux[2], self._aux[2]), None)
        else:
            merged = tuple(np.where(accept[:, None, None], new, old) for new, old in zip(aux, self._aux))
        self._aux = merged
        self._check_convergence(accept)
        self.iterations[accept] += 1
        self._fresh |= accept
        self._accepted = accept
        return int(accept.sum())

    def run(self, max_sweeps=1000):
        """Sweep until every replica has converged or failed; returns the number of sweeps used."""
        for sweep in range(1, max_sweeps + 1):
            self.sweep()
            if (self.converged | self.failed).all():
                return sweep
        return max_sweeps

    @property
    def pos(self):
        return self._geometry(self.x)[0]

    @property
    def rvecs(self):
        return np.array(self._geometry(self.x)[1])


class DeviceReplicaQNOptimizer(object):
    """The same lockstep optimiser with EVERYTHING on the device (``csrc/mm_qn.cu``): Hessian models, spectra, ridge
    search, DOF mappings, accept / shrink and convergence state stay in HBM; the host reads one counter per call of
    ``mm_qn_sweep``.  ``ReplicaQNOptimizer`` (above) moves two [R, n, n] arrays over PCIe per sweep and runs the
    trust-radius algebra in NumPy.

    Same state machine, thresholds and attributes as ``ReplicaQNOptimizer`` for ``dof="cartesian"`` and ``dof="strain"``.
    """

    def __init__(self, batch, pos0, rvecs0, dof="cartesian", gpos_rms=1e-5, dpos_rms=1e-3, grvecs_rms=1e-5, drvecs_rms=1e-3,
                 trust_radius=1.0, small_radius=1e-5, too_small_radius=1e-10):
        import ctypes

        from .. import _lib

        if dof not in ("cartesian", "strain"):
            raise ValueError("dof must be 'cartesian' or 'strain'")
        self._lib_mod, self._ctypes = _lib, ctypes
        self._lib = _lib.load()
        self.batch, self.kind = batch, dof
        pos0 = np.ascontiguousarray(pos0, dtype=float)
        self.nrep, self.nnodes = pos0.shape[0], pos0.shape[1]
        self.rvecs0 = np.ascontiguousarray(rvecs0, dtype=float).reshape(self.nrep, 3, 3)
        _lib.check(self._lib.mm_set_rvecs_batch(batch._handle, _lib.ptr(self.rvecs0.reshape(self.nrep, 9).copy())))
        rv0 = jac = proj = None
        if dof == "cartesian":
            x0 = pos0.reshape(self.nrep, -1).copy()
        else:  # StrainCellDOF (dof.py:522-697), as ReplicaQNOptimizer sets it up
            frac = np.einsum("rni,rji->rnj", pos0, np.linalg.inv(self.rvecs0).transpose(0, 2, 1))
            cell0 = np.tile(np.array([1.0, 1.0, 1.0, 0.0, 0.0, 0.0]), (self.nrep, 1))
            x0 = np.ascontiguousarray(np.concatenate([cell0, frac.reshape(self.nrep, -1)], axis=1))
            basis = np.zeros((6, 3, 3))
            for k, (i, j) in enumerate(_STRAIN_SLOTS):
                if i == j:
                    basis[k, i, i] = 1.0
                else:
                    basis[k, i, j] = basis[k, j, i] = 0.5
            self._basis = basis
            jac = np.ascontiguousarray(np.einsum("kij,rjl->rkil", basis, self.rvecs0).reshape(self.nrep, 6, 9).transpose(0, 2, 1))
            u = np.linalg.svd(jac, full_matrices=False)[0]
            proj = np.ascontiguousarray(u @ u.transpose(0, 2, 1))
            rv0 = np.ascontiguousarray(self.rvecs0.reshape(self.nrep, 9))
        self.ndof = x0.shape[1]
        self._q = ctypes.c_void_p()
        _lib.check(self._lib.mm_qn_create(batch._handle, 0 if dof == "cartesian" else 1, _lib.ptr(x0), _lib.ptr(rv0), _lib.ptr(jac),
                                          _lib.ptr(proj), gpos_rms, dpos_rms, grvecs_rms, drvecs_rms, trust_radius, small_radius,
                                          too_small_radius, ctypes.byref(self._q)))
        self.nlive = self.nrep

    def __del__(self):
        q = getattr(self, "_q", None)
        if q is not None and q.value:
            self._lib.mm_qn_destroy(q)
            self._q = None

    def run(self, max_sweeps=1000, check_every=4):
        """Sweep until every replica has converged or failed; returns the number of sweeps used."""
        done = 0
        live = self._ctypes.c_int32(self.nrep)
        while done < max_sweeps and self.nlive > 0:
            k = min(check_every, max_sweeps - done)
            self._lib_mod.check(self._lib.mm_qn_sweep(self._q, k, self._ctypes.byref(live)))
            done += k
            self.nlive = int(live.value)
        return done

    def _fetch(self):
        lib, ptr = self._lib_mod, self._lib_mod.ptr
        x, g = np.zeros((self.nrep, self.ndof)), np.zeros((self.nrep, self.ndof))
        f, radius, conv_val = np.zeros(self.nrep), np.zeros(self.nrep), np.zeros(self.nrep)
        ints = [np.zeros(self.nrep, dtype=np.int32) for _ in range(4)]
        evals = self._ctypes.c_int64()
        lib.check(self._lib.mm_qn_get(self._q, ptr(x), ptr(f), ptr(g), ptr(radius), ptr(conv_val), ptr(ints[0]), ptr(ints[1]),
                                      ptr(ints[2]), ptr(ints[3]), self._ctypes.byref(evals)))
        return dict(x=x, f=f, g=g, trust_radius=radius, conv_val=conv_val, iterations=ints[0].astype(np.int64),
                    converged=ints[1].astype(bool), failed=ints[2].astype(bool), conv_count=ints[3].astype(np.int64),
                    evaluations=int(evals.value))

    def __getattr__(self, name):
        if name in ("x", "f", "g", "trust_radius", "conv_val", "iterations", "converged", "failed", "conv_count", "evaluations"):
            return self._fetch()[name]
        raise AttributeError(name)

    def _geometry(self, x):
        if self.kind == "cartesian":
            return x.reshape(self.nrep, self.nnodes, 3), self.rvecs0
        rvecs = np.einsum("rk,kij->rij", x[:, :6], self._basis) @ self.rvecs0
        return x[:, 6:].reshape(self.nrep, self.nnodes, 3) @ rvecs, rvecs

    @property
    def pos(self):
        return self._geometry(self.x)[0]

    @property
    def rvecs(self):
        return np.array(self._geometry(self.x)[1])
