"""The lockstep replica optimiser on ReplicaBatch (one CUDA kernel sequence per sweep for all replicas) against the
same optimiser driven by the CPU oracle, for defect configurations x perturbations (BASELINE.json config 5, scaled
down)."""
import numpy as np
import pytest

from test_batchopt_cpu import OracleReplicaEvaluator, replicas

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["cartesian", "strain"])
def test_replica_optimiser_on_gpu_matches_oracle_driven_run(kind):
    from micmec_b200.replicas import ReplicaBatch
    from micmec_b200.sampling.batchopt import ReplicaQNOptimizer

    if kind == "cartesian":
        kwargs, args = dict(gpos_rms=1e-7, dpos_rms=1e-5), (["3x3x3_conf0", "3x3x3_conf3", "3x3x3_conf9"], 8, 0.5)
    else:
        kwargs = dict(gpos_rms=1e-8, dpos_rms=1e-6, grvecs_rms=1e-8, drvecs_rms=1e-6)
        args = (["3x3x3_conf0", "3x3x3_conf3", "3x3x3_conf9"], 4, 0.3, 0.02)
    systems = replicas(*args)
    pos0 = np.stack([s.pos for s in systems])
    rvecs0 = np.stack([np.array(s.domain.rvecs) for s in systems])
    batch = ReplicaBatch(systems)
    gpu = ReplicaQNOptimizer(batch, pos0, rvecs0, dof=kind, **kwargs)
    sweeps = gpu.run(400)
    assert gpu.converged.all() and not gpu.failed.any()
    assert gpu.evaluations == sweeps + 1 and batch.launches > 0
    cpu = ReplicaQNOptimizer(OracleReplicaEvaluator(replicas(*args)), pos0, rvecs0, dof=kind, eigh="lapack", **kwargs)
    cpu.run(400)
    assert cpu.converged.all()
    for r in range(len(systems)):
        scale = np.sqrt(np.mean((cpu.pos[r] - pos0[r]) ** 2))
        assert np.max(np.abs(gpu.pos[r] - cpu.pos[r])) <= 1e-4 * scale, r
        assert np.max(np.abs(gpu.rvecs[r] - cpu.rvecs[r])) <= 1e-6 * np.sqrt(np.mean(cpu.rvecs[r] ** 2)), r
        assert abs(gpu.f[r] - cpu.f[r]) <= 1e-8 * float(cpu.f_old.max()) + 1e-12, r
        assert abs(int(gpu.iterations[r]) - int(cpu.iterations[r])) <= 6, r


@pytest.mark.parametrize("n", [1, 2, 27, 81, 87, 96])
def test_batched_eigh_on_gpu_matches_lapack(n):
    """mm_batched_eigh (cyclic Jacobi, one block per matrix) against numpy.linalg.eigh: eigenvalues, reconstruction,
    orthogonality, ascending order - random matrices, the identity, SR1-like low-rank updates of it."""
    from micmec_b200.sampling.batchopt import _eigh

    rng = np.random.default_rng(2)
    a = rng.normal(size=(150, n, n))
    h = a + a.transpose(0, 2, 1)
    h[0] = np.eye(n)
    u = rng.normal(size=(n,))
    h[1] = np.eye(n) + 0.37 * np.outer(u, u)
    h[2] = 0.0
    h[3] *= 1e-12
    evals, evecs = _eigh(h, 0, "device")
    scale = np.max(np.abs(h), axis=(1, 2)) + 1e-300
    assert evals.shape == (150, n) and np.all(np.diff(evals, axis=1) >= 0)
    back = np.einsum("rij,rj,rkj->rik", evecs, evals, evecs)
    assert np.max(np.max(np.abs(back - h), axis=(1, 2)) / scale) <= 1e-13 * n
    gram = np.einsum("rki,rkj->rij", evecs, evecs)
    assert np.max(np.abs(gram - np.eye(n))) <= 1e-13 * n
    assert np.max(np.max(np.abs(evals - np.linalg.eigvalsh(h)), axis=1) / scale) <= 1e-13 * n


def test_batched_eigh_rejects_bad_sizes():
    from micmec_b200.sampling.batchopt import _eigh

    with pytest.raises(ValueError):
        _eigh(np.zeros((2, 97, 97)), 0, "device")


@pytest.mark.parametrize("kind", ["cartesian", "strain"])
def test_device_resident_optimiser_matches_host_driven_lockstep_run(kind):
    """mm_qn (SR1 update, spectra, ridge search, DOF mappings, accept / shrink, convergence all on the device) against the
    host-driven lockstep optimiser on the same replica batch: same minima, same discrete outcome, iteration counts within
    the tolerance the host-driven optimiser has against the oracle-driven one."""
    from micmec_b200.replicas import ReplicaBatch
    from micmec_b200.sampling.batchopt import ReplicaQNOptimizer, DeviceReplicaQNOptimizer

    if kind == "cartesian":
        kwargs, args = dict(gpos_rms=1e-7, dpos_rms=1e-5), (["3x3x3_conf0", "3x3x3_conf3", "3x3x3_conf9"], 16, 0.5)
    else:
        kwargs = dict(gpos_rms=1e-8, dpos_rms=1e-6, grvecs_rms=1e-8, drvecs_rms=1e-6)
        args = (["3x3x3_conf0", "3x3x3_conf3", "3x3x3_conf9"], 8, 0.3, 0.02)
    systems = replicas(*args)
    pos0 = np.stack([s.pos for s in systems])
    rvecs0 = np.stack([np.array(s.domain.rvecs) for s in systems])
    dev = DeviceReplicaQNOptimizer(ReplicaBatch(systems), pos0, rvecs0, dof=kind, **kwargs)
    sweeps = dev.run(400)
    assert dev.converged.all() and not dev.failed.any() and sweeps < 400
    host = ReplicaQNOptimizer(ReplicaBatch(replicas(*args)), pos0, rvecs0, dof=kind, **kwargs)
    host.run(400)
    assert host.converged.all()
    f0 = float(host.f_old.max())
    for r in range(len(systems)):
        scale = np.sqrt(np.mean((host.pos[r] - pos0[r]) ** 2))
        assert np.max(np.abs(dev.pos[r] - host.pos[r])) <= 1e-4 * scale, r
        assert np.max(np.abs(dev.rvecs[r] - host.rvecs[r])) <= 1e-6 * np.sqrt(np.mean(host.rvecs[r] ** 2)), r
        assert abs(dev.f[r] - host.f[r]) <= 1e-8 * f0 + 1e-12, r
        assert abs(int(dev.iterations[r]) - int(host.iterations[r])) <= 6, r
    assert np.all(dev.conv_count == 0) and np.all(dev.conv_val < 1.0)
