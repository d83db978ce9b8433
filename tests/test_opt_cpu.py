"""Host logic of the optimisers and DOF mappings (micmec_b200/sampling/opt.py, dof.py) against QNOptimizer runs
recorded from the UNMODIFIED reference.  Forces come from the CPU oracle here (test infrastructure);
tests/test_opt_gpu.py repeats the cases on the CUDA force part."""
import numpy as np
import pytest

import optcases
from oraclepart import OracleForcePart


@pytest.mark.parametrize("tag", optcases.CASES)
def test_qn_optimizer_matches_reference(tag):
    optcases.run_case(tag, lambda system: OracleForcePart(system))


@pytest.mark.parametrize("tag", ["cartesian_3x3x3_conf0", "strain_3x3x3_test", "fullcell_2x2x2_reo"])
def test_dof_gradient_is_consistent_with_energy(tag):
    d, mmf, dof = optcases.build(tag, lambda system: OracleForcePart(system))
    np.random.seed(3)
    assert dof.check_delta(eps=1e-4) <= 1e-4


def test_solve_trust_radius_properties():
    from micmec_b200.sampling.opt import solve_trust_radius

    rng = np.random.default_rng(0)
    for trial in range(20):
        evals = rng.normal(0.5, 1.0, 12)
        grad = rng.normal(0.0, 1.0, 12)
        radius = 10.0 ** rng.uniform(-3, 0.5)
        step = solve_trust_radius(grad, evals, radius)
        norm = np.linalg.norm(step)
        if evals.min() > 0 and np.linalg.norm(grad / evals) <= radius:
            assert np.allclose(step, -grad / evals)
        else:
            assert abs(norm - radius) <= 1e-5 * radius * 1.0001
            ridge = -(grad / step)[0] - evals[0]
            assert ridge > -evals.min()  # the shifted model is positive definite


def test_hessian_models():
    from micmec_b200.sampling.opt import SR1HessianModel, BFGSHessianModel

    rng = np.random.default_rng(1)
    a = rng.normal(size=(6, 6))
    true = a @ a.T + np.eye(6)
    for cls in (SR1HessianModel, BFGSHessianModel):
        model = cls(6)
        for _ in range(40):
            dx = rng.normal(size=6)
            model.update(dx, true @ dx)
            assert np.allclose(model.hessian @ dx, true @ dx)  # secant condition of the last update
        if cls is SR1HessianModel:  # SR1 recovers a quadratic's Hessian exactly after ndof independent updates
            assert np.allclose(model.hessian, true, atol=1e-9)
        assert np.all(np.linalg.eigvalsh(model.hessian) > 0) or cls is SR1HessianModel
        assert not model.update(np.zeros(6), np.zeros(6))
    with pytest.raises(TypeError):
        SR1HessianModel(3, np.eye(4))


@pytest.mark.parametrize("tag", ["cartesian_3x3x3_conf0", "strain_3x3x3_test"])
def test_cg_optimizer_converges_to_the_reference_minimum(tag):
    """CGOptimizer (opt.py:147-180) on the self-contained conjugate-gradient minimiser: molmod.minimizer is not in the
    reference tree, so the iterates are not pinned; the minimum is - it must be the one the reference's QNOptimizer
    run ended in."""
    import goldenio as gio
    from micmec_b200.sampling.opt import CGOptimizer

    d, mmf, dof = optcases.build(tag, lambda system: OracleForcePart(system))
    opt = CGOptimizer(dof)
    f0 = float(d["iter0:f"])
    opt.run(400)
    assert dof.converged
    last = "iter%d:" % int(d["meta:iterations"])
    assert opt.epot <= float(d[last + "f"]) + 1e-7 * abs(f0)
    assert gio.rel_rms(mmf.system.pos, d[last + "pos"]) <= 1e-4
    assert gio.rel_rms(np.array(mmf.system.domain.rvecs), d[last + "rvecs"]) <= 1e-5
