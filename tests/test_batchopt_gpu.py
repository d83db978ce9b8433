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
    cpu = ReplicaQNOptimizer(OracleReplicaEvaluator(replicas(*args)), pos0, rvecs0, dof=kind, **kwargs)
    cpu.run(400)
    assert cpu.converged.all()
    for r in range(len(systems)):
        scale = np.sqrt(np.mean((cpu.pos[r] - pos0[r]) ** 2))
        assert np.max(np.abs(gpu.pos[r] - cpu.pos[r])) <= 1e-4 * scale, r
        assert np.max(np.abs(gpu.rvecs[r] - cpu.rvecs[r])) <= 1e-6 * np.sqrt(np.mean(cpu.rvecs[r] ** 2)), r
        assert abs(gpu.f[r] - cpu.f[r]) <= 1e-8 * float(cpu.f_old.max()) + 1e-12, r
        assert abs(int(gpu.iterations[r]) - int(cpu.iterations[r])) <= 6, r


def test_batched_eigh_on_gpu_reconstructs_the_models():
    """Batches of 64 or more Hessian models are diagonalised on the GPU (cuSOLVER through torch)."""
    from micmec_b200.sampling.batchopt import _eigh

    rng = np.random.default_rng(2)
    a = rng.normal(size=(96, 87, 87))
    h = a + a.transpose(0, 2, 1)
    evals, evecs = _eigh(h)
    assert evals.shape == (96, 87) and np.all(np.diff(evals, axis=1) >= 0)
    back = np.einsum("rij,rj,rkj->rik", evecs, evals, evecs)
    assert np.max(np.abs(back - h)) <= 1e-10 * np.max(np.abs(h))
    assert np.allclose(evals, np.linalg.eigvalsh(h), rtol=0, atol=1e-10 * np.max(np.abs(h)))
