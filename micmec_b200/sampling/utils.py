"""Init-time helpers of the integrators; they stay on the host so that the legacy global NumPy RNG is consumed in
the same order as in the reference (micmec/sampling/utils.py).
"""
import numpy as np

from ..units import boltzmann

__all__ = [
    "get_random_vel", "remove_com_moment", "clean_momenta", "get_ndof_internal_md", "domain_symmetrize",
    "get_random_vel_press", "get_ndof_baro", "stabilized_cholesky_decomp",
]


def get_random_vel(temp0, scalevel0, masses, select=None):
    """Maxwell-Boltzmann velocities, optionally rescaled to ``temp0`` exactly (sampling/utils.py:28-65)."""
    if select is not None:
        masses = masses[select]
    sigma = np.sqrt(boltzmann * temp0 / masses).reshape(-1, 1)
    vel0 = np.random.normal(0, 1, (len(masses), 3)) * sigma
    if scalevel0 and temp0 > 0:
        temp = np.mean(vel0 ** 2 * masses.reshape(-1, 1)) / boltzmann
        vel0 *= np.sqrt(temp0 / temp)
    return vel0


def remove_com_moment(vel, masses):
    """Subtract the centre-of-mass velocity in place (sampling/utils.py:68-92)."""
    vel[:] -= np.dot(masses, vel) / np.sum(masses)


def clean_momenta(pos, vel, masses, domain):
    """Remove the external momenta that are conserved for this periodicity (sampling/utils.py:142-174)."""
    remove_com_moment(vel, masses)
    if domain.nvec == 0:
        # isolated system: also remove the rigid-body rotation about the centre of mass
        rel = pos - np.dot(masses, pos) / np.sum(masses)
        r2 = (rel ** 2).sum(axis=1)
        itens = np.einsum("n,n,ij->ij", masses, r2, np.eye(3)) - np.einsum("n,ni,nj->ij", masses, rel, rel)
        amom = np.einsum("n,ni->i", masses, np.cross(rel, vel))
        evals, evecs = np.linalg.eigh(itens)
        keep = evals > 1e-10  # pseudo-inverse of the inertia tensor (sampling/utils.py:281-287)
        avel = evecs[:, keep] @ ((evecs[:, keep].T @ amom) / evals[keep])
        vel[:] -= np.cross(avel, rel)
    elif domain.nvec == 1:
        raise NotImplementedError


def get_ndof_internal_md(nnodes, nper):
    """Internal degrees of freedom of an MD run (sampling/utils.py:322-343)."""
    if nper == 0:
        return 3 * nnodes - 6
    if nper == 1:
        return 3 * nnodes - 4
    return 3 * nnodes - 3


def domain_symmetrize(mmf, vector_lst=None, tensor_lst=None):
    """Rotate the domain so that ``rvecs`` is symmetric; positions follow (sampling/utils.py:346-383)."""
    domain = np.array(mmf.system.domain.rvecs)
    U, _, Vt = np.linalg.svd(domain)
    rot = Vt.T @ U.T
    mmf.update_rvecs(np.ascontiguousarray(domain @ rot))
    mmf.update_pos(mmf.system.pos @ rot)
    vectors = [v @ rot for v in (vector_lst or [])]
    tensors = [rot.T @ t @ rot for t in (tensor_lst or [])]
    return vectors, tensors


def get_random_vel_press(mass, temp):
    """Random symmetric barostat velocity tensor (sampling/utils.py:450-475)."""
    rand = np.random.normal(0, np.sqrt(mass * boltzmann * temp), (3, 3)) / mass
    low = np.tril(rand)
    vel_press = low + np.tril(rand, -1).T
    vel_press[~np.eye(3, dtype=bool)] /= np.sqrt(2)
    return vel_press


def get_ndof_baro(dim, anisotropic, vol_constraint):
    """Degrees of freedom of the fluctuating domain (sampling/utils.py:478-501)."""
    ndof = dim * (dim + 1) // 2 if anisotropic else 1
    if vol_constraint:
        ndof -= 1
    if ndof == 0:
        raise AssertionError("Isotropic barostat called with a volume constraint.")
    return ndof


def stabilized_cholesky_decomp(mat):
    """Factor M with M M^T = mat for a symmetric matrix that may be slightly indefinite: plain Cholesky when positive
    definite, else L D L^T with the negative pivots of D clipped to zero (sampling/utils.py:504-528)."""
    if np.all(np.linalg.eigvals(mat) > 0):
        return np.linalg.cholesky(mat)
    n = mat.shape[0]
    diag, low = np.zeros(n), np.eye(n)
    for i in range(n):
        for j in range(i):
            acc = mat[i, j] - np.sum(low[i, :j] * low[j, :j] * diag[:j])
            low[i, j] = acc / diag[j] if abs(diag[j]) > 1e-12 else 0.0
        diag[i] = mat[i, i] - np.sum(low[i, :i] ** 2 * diag[:i])
    return low * np.sqrt(diag.clip(min=0))
