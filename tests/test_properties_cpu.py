"""Property tests (hypothesis) of host-side pieces whose inputs are easy to randomise: the .chk codec, the 3x3
completion of lower-dimensional domains, the batched trust-radius solver against the scalar one."""
import numpy as np
from hypothesis import given, settings, strategies as st
from hypothesis.extra import numpy as hnp

finite = st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, allow_infinity=False, width=64)


@settings(max_examples=40, deadline=None)
@given(
    arr=hnp.arrays(np.float64, hnp.array_shapes(min_dims=1, max_dims=4, min_side=1, max_side=5), elements=finite),
    ints=hnp.arrays(np.int64, hnp.array_shapes(min_dims=1, max_dims=2, min_side=1, max_side=7),
                    elements=st.integers(min_value=-10**9, max_value=10**9)),
    scalar=finite,
    flag=st.booleans(),
)
def test_chk_codec_round_trips_exactly(tmp_path_factory, arr, ints, scalar, flag):
    from micmec_b200.chk import dump_chk, load_chk

    fn = str(tmp_path_factory.mktemp("chk") / "x.chk")
    data = {"type1/cell": arr, "surrounding_nodes": ints, "type1/effective_temp": scalar, "flag": flag}
    dump_chk(fn, data)
    back = load_chk(fn)
    # the format keeps 16 significant digits ("% 22.15e", as molmod writes it): values come back as their 16-digit
    # roundings, which is exact for everything that was read from a .chk file in the first place
    as_written = np.array([float("%.15e" % v) for v in arr.ravel()]).reshape(arr.shape)
    assert back["type1/cell"].shape == arr.shape and np.array_equal(back["type1/cell"], as_written)
    dump_chk(fn, back)
    assert np.array_equal(load_chk(fn)["type1/cell"], as_written)  # idempotent from then on
    assert np.array_equal(back["surrounding_nodes"], ints) and back["flag"] == flag
    assert back["type1/effective_temp"] == float("%.15e" % scalar)


@settings(max_examples=40, deadline=None)
@given(rv=hnp.arrays(np.float64, st.sampled_from([(1, 3), (2, 3), (3, 3)]),
                     elements=st.floats(min_value=-5.0, max_value=5.0, allow_nan=False, width=64)))
def test_domain_completion_is_a_dual_basis(rv):
    from micmec_b200.system import Domain

    if np.linalg.svd(rv, compute_uv=False).min() < 1e-2:
        return  # degenerate cell
    dom = Domain(rv)
    full_r, full_g = dom._get_rvecs(full=True), dom._get_gvecs(full=True)
    assert full_r.shape == (3, 3) and np.array_equal(full_r[: len(rv)], rv)
    assert np.allclose(full_r @ full_g.T, np.eye(3), atol=1e-9)  # reciprocal basis of the completed cell
    assert np.allclose(np.asarray(dom.gvecs), full_g[: len(rv)], atol=1e-9)
    if len(rv) < 3:  # the complement is orthonormal and orthogonal to the periodic vectors
        comp = full_r[len(rv):]
        assert np.allclose(comp @ comp.T, np.eye(3 - len(rv)), atol=1e-9) and np.allclose(comp @ rv.T, 0.0, atol=1e-9)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(min_value=0, max_value=2**31 - 1), n=st.integers(min_value=2, max_value=20),
       shift=st.floats(min_value=-2.0, max_value=2.0), logr=st.floats(min_value=-4.0, max_value=1.0))
def test_batched_trust_radius_solver_agrees_with_the_scalar_one(seed, n, shift, logr):
    from micmec_b200.sampling.batchopt import solve_trust_radius_batch
    from micmec_b200.sampling.opt import solve_trust_radius

    rng = np.random.default_rng(seed)
    evals = rng.normal(shift, 1.0, (6, n))
    grad = rng.normal(0.0, 1.0, (6, n))
    radius = np.full(6, 10.0 ** logr) * rng.uniform(0.5, 2.0, 6)
    steps = solve_trust_radius_batch(grad, evals, radius)
    for r in range(6):
        ref = solve_trust_radius(grad[r], evals[r], radius[r])
        assert np.allclose(steps[r], ref, rtol=1e-8, atol=1e-12 * radius[r])
        assert np.linalg.norm(steps[r]) <= radius[r] * (1 + 2e-5)
