"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every declared symbol."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from micmec_b200 import build, _lib

    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "micmec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mm_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from micmec_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "libmicmec_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == names


def test_version_and_error_string(lib):
    assert lib.mm_version() >= 100
    assert isinstance(lib.mm_last_error(), bytes)


def test_domain_native_matches_reference_formulas(lib):
    """mm_domain vs micmec/pes/domain.c:23-48 (volume) and ext.pyx:64-71 (gvecs = pseudo-inverse transpose)."""
    from micmec_b200.system import Domain

    rng = np.random.default_rng(3)
    for nvec in (1, 2, 3):
        rvecs = rng.normal(0.0, 10.0, (nvec, 3)) + 20.0 * np.eye(3)[:nvec]
        dom = Domain(rvecs)
        assert dom.nvec == nvec
        assert np.allclose(dom.gvecs, np.linalg.pinv(rvecs).T, rtol=1e-12, atol=1e-14)
        gram = rvecs @ rvecs.T
        assert abs(dom.volume - np.sqrt(abs(np.linalg.det(gram)))) <= 1e-10 * dom.volume
        assert not dom.rvecs.flags.writeable
    assert Domain(np.zeros((0, 3))).nvec == 0 and Domain(np.zeros((0, 3))).volume == 0.0
    with pytest.raises(TypeError):
        Domain(np.zeros((4, 3)))


def test_domain_matches_reference_ext():
    """Against the reference's own compiled Domain (oracle/_ref, build container only)."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import refenv

    if not refenv.available() or not os.path.isdir(os.path.join(ROOT, "oracle", "_ref")):
        pytest.skip("reference tree / oracle/_ref not present")
    refenv.setup()
    from micmec.pes.ext import Domain as RefDomain
    from micmec_b200.system import Domain

    rng = np.random.default_rng(4)
    for nvec in (1, 2, 3):
        rvecs = np.ascontiguousarray(rng.normal(0.0, 10.0, (nvec, 3)) + 20.0 * np.eye(3)[:nvec])
        ours, ref = Domain(rvecs), RefDomain(rvecs)
        assert np.allclose(ours.gvecs, ref.gvecs, rtol=1e-12, atol=1e-15)
        assert abs(ours.volume - ref.volume) <= 1e-12 * ref.volume
        assert np.array_equal(ours.rvecs, ref.rvecs)


def test_compute_entry_points_fail_loudly_without_gpu(lib):
    """No CPU fallback: without a CUDA device mm_create must fail with a CUDA error, not compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from micmec_b200.system import System
    from micmec_b200.pes.mmff import ForcePartMechanical
    from micmec_b200.celltypes import TYPE_TEST

    system = System.periodic_grid((3, 3, 3), TYPE_TEST)
    with pytest.raises(RuntimeError):
        ForcePartMechanical(system)


def test_topology_compiler_matches_reference_mic():
    """cell_shifts (closed form) vs the dense mic table of the reference, on every golden fixture."""
    import goldenio as gio
    from micmec_b200.topology import cell_shifts, periodic_grid_arrays

    for name in gio.force_fixtures():
        d = gio.load("force_" + name)
        shift = cell_shifts(d["grid"], d["surrounding_nodes"].shape[0], True)
        assert np.array_equal(shift, d["shift_ref"]), name
    # the closed-form grid generator reproduces the reference's index arrays on full grids
    for name, shape in (("2x2x2_test", (2, 2, 2)), ("3x3x3_test", (3, 3, 3)), ("4x4x4_fcu", (4, 4, 4))):
        d = gio.load("force_" + name)
        sn, sc, bn = periodic_grid_arrays(shape)
        assert np.array_equal(sn, d["surrounding_nodes"])
        assert np.array_equal(sc, d["surrounding_cells"])
        assert np.array_equal(bn, d["boundary_nodes"])
