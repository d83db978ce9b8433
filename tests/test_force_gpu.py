"""GPU parity of ForcePartMechanical.compute (CUDA, through the C ABI) with the reference.

Checked against (a) the golden vectors of the UNMODIFIED reference and (b) the CPU oracle on the same inputs.
Tolerances are the north_star's: energy 1e-10 relative; gpos and vtens 1e-9 relative to RMS (the virial
additionally gets the reference's own cancellation-noise floor, see goldenio.virial_noise).
"""
import numpy as np
import pytest

import goldenio as gio
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

MODELS = ["original", "default"]
ETOL, GTOL = 1e-10, 1e-9


def make_system(d, prefix=""):
    from micmec_b200.system import System

    rec = gio.system_from(d, prefix)
    return System(rec.pos, rec.masses, rec.rvecs, rec.surrounding_cells, rec.surrounding_nodes,
                  boundary_nodes=rec.boundary_nodes, grid=rec.grid, types=rec.types, params=rec.params)


def make_mmf(system, model):
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical

    part = ForcePartMechanical(system, model=model)
    return MicMecForceField(system, [part]), part


def check_against(e, g, v, eref, gref, vref, noise, rest=False):
    assert abs(e - eref) <= ETOL * max(abs(eref), 1e-8), (e, eref)
    if rest:
        # at rest the strain is ~1e-7 (h0 is not exactly the node spacing): 0.5 (G G^T - I) cancels 7 digits in the
        # reference as well, so only absolute agreement is meaningful
        assert np.max(np.abs(g - gref)) <= 1e-14
        assert np.max(np.abs(v - vref)) <= 1e-11
        return
    if np.sqrt(np.mean(gref ** 2)) < 1e-12:
        assert np.max(np.abs(g - gref)) <= 1e-14
    else:
        assert gio.rel_rms(g, gref) <= GTOL
    assert gio.virial_close(v, vref, GTOL, noise + 1e-13)


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("name", gio.force_fixtures())
def test_compute_matches_reference_golden(name, model):
    d = gio.load("force_" + name)
    system = make_system(d)
    mmf, part = make_mmf(system, model)
    oracle = orc.Oracle(gio.system_from(d), model=model)
    for case in ("rest", "rng0", "rng1", "shear"):
        pos, rvecs = d[case + ":pos"], d[case + ":rvecs"]
        mmf.update_rvecs(np.ascontiguousarray(rvecs))
        mmf.update_pos(pos)
        gpos, vtens = np.zeros(pos.shape), np.zeros((3, 3))
        e = mmf.compute(gpos, vtens)
        key = "%s:%s:" % (case, model)
        _, gc, vc = oracle.deformation(pos, rvecs)
        noise = gio.virial_noise(gc, vc)
        rest = case == "rest"
        check_against(e, gpos, vtens, float(d[key + "energy"]), d[key + "gpos"], d[key + "vtens"], noise, rest)
        eo, go, vo = oracle.compute(pos, rvecs, gpos=True, vtens=True)
        check_against(e, gpos, vtens, eo, go, vo, noise, rest)
        if case == "rng0":  # per-cell caches, mmff.py:290-292
            assert gio.rel_rms(part.epot_cells, d[key + "epot_cells"]) <= GTOL
            assert gio.rel_rms(part.gpos_cells, d[key + "gpos_cells"]) <= GTOL


@pytest.mark.parametrize("model", MODELS)
def test_multistate_mixing_matches_reference_golden(model):
    d = gio.load("multistate")
    for tag in ("a", "b"):
        system = make_system(d, tag + ":")
        mmf, _ = make_mmf(system, model)
        oracle = orc.Oracle(gio.system_from(d, tag + ":"), model=model)
        for case in ("small", "large"):
            base = "%s:%s:" % (tag, case)
            pos, rvecs = d[base + "pos"], d[base + "rvecs"]
            mmf.update_rvecs(np.ascontiguousarray(rvecs))
            mmf.update_pos(pos)
            gpos, vtens = np.zeros(pos.shape), np.zeros((3, 3))
            e = mmf.compute(gpos, vtens)
            key = base + model + ":"
            _, gc, vc = oracle.deformation(pos, rvecs)
            check_against(e, gpos, vtens, float(d[key + "energy"]), d[key + "gpos"], d[key + "vtens"],
                          gio.virial_noise(gc, vc))


def test_plugin_contract():
    """ForcePart.compute semantics (mmff.py:87-149): accumulate into caller arrays, optional outputs, caches, NaN."""
    d = gio.load("force_3x3x3_conf0")
    system = make_system(d)
    mmf, part = make_mmf(system, "original")
    assert part.name == "micmec" and mmf.part_micmec is part
    pos = d["rng0:pos"]
    mmf.update_pos(pos)
    assert np.isnan(mmf.energy) and np.isnan(part.gpos).all()  # clear() after update_pos
    gpos, vtens = np.full(pos.shape, 2.0), np.full((3, 3), -1.0)
    e = mmf.compute(gpos, vtens)
    assert gio.rel_rms(gpos - 2.0, d["rng0:original:gpos"]) <= GTOL  # ADDED to the caller's content
    assert np.allclose(vtens + 1.0, d["rng0:original:vtens"], rtol=1e-8, atol=1e-12)
    assert mmf.energy == e and part.energy == e
    # any combination of outputs (stress_strain.py:85 asks for vtens only)
    v2 = np.zeros((3, 3))
    assert mmf.compute(vtens=v2) == e and np.array_equal(v2, part.vtens)
    g2 = np.zeros(pos.shape)
    assert mmf.compute(gpos=g2) == e and np.array_equal(g2, part.gpos) and np.allclose(g2, gpos - 2.0, atol=1e-15)
    assert mmf.compute() == e
    # NaN positions raise the reference's ValueError
    bad = pos.copy()
    bad[3, 1] = np.nan
    mmf.update_pos(bad)
    with pytest.raises(ValueError):
        mmf.compute(np.zeros(pos.shape))
    with pytest.raises(ValueError):
        from micmec_b200.pes.mmff import MicMecForceField

        MicMecForceField(system, [part, part])  # duplicate part name, mmff.py:186
    assert part.launches > 0


@pytest.mark.parametrize("model", MODELS)
def test_structured_grid_equals_indexed_and_oracle(model):
    """Implicit-topology periodic grid == explicit index arrays == oracle, on a perturbed 12x10x8 fcu grid."""
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU

    shape = (12, 10, 8)
    sys_a = System.periodic_grid(shape, TYPE_FCU, explicit=True)
    sys_b = System.periodic_grid(shape, TYPE_FCU, explicit=False)
    rng = np.random.default_rng(0)
    pos = sys_a.pos + 0.5 * rng.standard_normal(sys_a.pos.shape)
    strain = np.eye(3) + 0.01 * rng.standard_normal((3, 3))
    pos, rvecs = pos @ strain, np.ascontiguousarray(sys_a.domain.rvecs @ strain)
    res = []
    for system in (sys_a, sys_b):
        mmf, _ = make_mmf(system, model)
        mmf.update_rvecs(rvecs)
        mmf.update_pos(pos)
        g, v = np.zeros(pos.shape), np.zeros((3, 3))
        res.append((mmf.compute(g, v), g, v))
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    oracle = orc.Oracle(sys_a, model=model, nthreads=4)
    eo, go, vo = oracle.compute(pos, rvecs, gpos=True, vtens=True)
    _, gc, vc = oracle.deformation(pos, rvecs)
    check_against(res[0][0], res[0][1], res[0][2], eo, go, vo, gio.virial_noise(gc, vc))


def test_size_independent_properties_64cubed():
    """At a BASELINE size (64^3 cells) the oracle is too slow for every check; use invariants instead:
    sum of gradients = 0, translation invariance, symmetric virial, virial = dE/d(strain) by finite differences,
    and energy/gradient agreement with the oracle on the same input (oracle: OpenMP, a few seconds)."""
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_TEST

    shape = (64, 64, 64)
    system = System.periodic_grid(shape, TYPE_TEST, explicit=False)
    mmf, _ = make_mmf(system, "original")
    rng = np.random.default_rng(1)
    pos0 = system.pos + 0.3 * rng.standard_normal(system.pos.shape)
    rvecs0 = np.array(system.domain.rvecs)

    def evaluate(pos, rvecs):
        mmf.update_rvecs(np.ascontiguousarray(rvecs))
        mmf.update_pos(pos)
        g, v = np.zeros(pos.shape), np.zeros((3, 3))
        return mmf.compute(g, v), g, v

    e, g, v = evaluate(pos0, rvecs0)
    rms = np.sqrt(np.mean(g ** 2))
    assert np.max(np.abs(g.sum(axis=0))) <= 1e-9 * rms * np.sqrt(len(g))
    assert np.max(np.abs(v - v.T)) <= 1e-12 * np.max(np.abs(v))
    e2, g2, _ = evaluate(pos0 + np.array([3.1, -7.7, 11.3]), rvecs0)
    assert abs(e2 - e) <= 1e-10 * abs(e) and gio.rel_rms(g2, g) <= 1e-8
    # virial = dE/dD under pos -> pos.D, rvecs -> rvecs.D (TYPE_TEST has a major-symmetric C, so gpos is exact)
    h = 1e-5
    for (a, b) in ((0, 0), (1, 2)):
        D = np.eye(3)
        D[a, b] += h
        ep = evaluate(pos0 @ D, rvecs0 @ D)[0]
        D[a, b] -= 2 * h
        em = evaluate(pos0 @ D, rvecs0 @ D)[0]
        assert abs((ep - em) / (2 * h) - v[a, b]) <= 2e-6 * np.max(np.abs(v))
    # same input through the oracle (implicit topology on the GPU side, explicit arrays on the CPU side)
    from micmec_b200.topology import periodic_grid_arrays

    sn, sc, _ = periodic_grid_arrays(shape)
    oracle = orc.Oracle(model="original", nthreads=16, surrounding_nodes=sn, surrounding_cells=sc, grid=system.grid,
                        types=system.types, params=system.params, pbc=True)
    eo, go, vo = oracle.compute(pos0, rvecs0, gpos=True, vtens=True)
    assert abs(e - eo) <= ETOL * abs(eo)
    assert gio.rel_rms(g, go) <= GTOL
    assert gio.rel_rms(v, vo) <= 1e-7  # the oracle's own sum over absolute positions is the noisy side here


def test_errors_match_reference():
    from micmec_b200.system import System
    from micmec_b200.pes.mmff import ForcePartMechanical
    from micmec_b200.celltypes import TYPE_TEST

    system = System.periodic_grid((3, 3, 3), TYPE_TEST)
    system.domain.update_rvecs(np.array(system.domain.rvecs)[:2])
    with pytest.raises(ValueError):  # mmff.py:249-254: only 0-D or 3-D periodic
        ForcePartMechanical(system)
    system = System.periodic_grid((3, 3, 3), TYPE_TEST)
    with pytest.raises(ValueError):
        ForcePartMechanical(system, model="nonsense")


def two_type_params():
    """Two cell types; type 2 has two metastable states close enough in energy to mix."""
    from micmec_b200.celltypes import TYPE_FCU, TYPE_TEST

    params = {}
    for key in ("cell", "elasticity", "free_energy", "effective_temp", "mass"):
        params["type1/" + key] = TYPE_FCU[key]
    h0 = TYPE_FCU["cell"][0]
    params["type2/cell"] = np.array([h0 * 1.01, h0 * 1.03])
    params["type2/elasticity"] = np.array([TYPE_FCU["elasticity"][0] * 0.8, TYPE_FCU["elasticity"][0] * 0.6])
    params["type2/free_energy"] = np.array([0.0, 3.0e-4])
    params["type2/effective_temp"] = 350.0
    params["type2/mass"] = TYPE_FCU["mass"]
    return params


@pytest.mark.parametrize("shape,mixed,model", [((12, 10, 8), False, "original"), ((33, 7, 5), False, "original"),
                                               ((40, 20, 70), False, "original"), ((16, 12, 10), True, "original"),
                                               ((2, 2, 3), True, "original"), ((12, 10, 8), False, "default"),
                                               ((33, 7, 5), False, "default"), ((40, 20, 70), False, "default"),
                                               ((2, 3, 2), False, "default")])
@pytest.mark.parametrize("images", [0, 1])
def test_structured_kernels_match_generic_and_oracle(shape, mixed, model, images, monkeypatch):
    """The marching structured-grid kernels (k_march2 for one-type grids, k_march otherwise; periodic images from the
    ghost nodes or taken on load) vs the indexed kernels vs the oracle, both per-cell models."""
    monkeypatch.setenv("MICMEC_B200_WRAP_ON_LOAD", str(images))
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU

    system = System.periodic_grid(shape, TYPE_FCU, explicit=True)
    rng = np.random.default_rng(3)
    if mixed:
        system.types = rng.integers(1, 3, size=system.ncells)
        system.grid = system.types.reshape(shape)
        system.params = two_type_params()
    pos = system.pos + 0.4 * rng.standard_normal(system.pos.shape)
    strain = np.eye(3) + 0.01 * rng.standard_normal((3, 3))
    pos, rvecs = pos @ strain, np.ascontiguousarray(system.domain.rvecs @ strain)
    res = {}
    for structured in (False, True):
        sysx = System(system.pos.copy(), system.masses, np.array(system.domain.rvecs), None if structured else system.surrounding_cells,
                      None if structured else system.surrounding_nodes, grid=system.grid, types=system.types,
                      params=system.params, structured_shape=shape)
        from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical

        part = ForcePartMechanical(sysx, model=model, structured=structured)
        assert part.structured == structured
        mmf = MicMecForceField(sysx, [part])
        mmf.update_rvecs(rvecs)
        mmf.update_pos(pos)
        g, v = np.zeros(pos.shape), np.zeros((3, 3))
        res[structured] = (mmf.compute(g, v), g, v)
        v2 = np.zeros((3, 3))
        assert mmf.compute(vtens=v2) == res[structured][0] and np.array_equal(v2, v)  # vtens without gpos
    ea, ga, va = res[False]
    eb, gb, vb = res[True]
    assert abs(ea - eb) <= 1e-12 * abs(ea)
    assert gio.rel_rms(gb, ga) <= 1e-11
    assert gio.rel_rms(vb, va) <= 1e-11
    oracle = orc.Oracle(system, model=model, nthreads=8)
    eo, go, vo = oracle.compute(pos, rvecs, gpos=True, vtens=True)
    _, gc, vc = oracle.deformation(pos, rvecs)
    check_against(eb, gb, vb, eo, go, vo, gio.virial_noise(gc, vc))


@pytest.mark.parametrize("model", MODELS)
def test_atomic_scatter_matches_gather(model):
    """The cell-centric warp-aggregated atomic scatter gives the gather's gradient up to summation order."""
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU

    cases = [make_system(gio.load("force_5x5x5_fcu_hollow")), make_system(gio.load("force_3x3x3_conf0")),
             System.periodic_grid((9, 7, 40), TYPE_FCU, explicit=True)]
    rng = np.random.default_rng(5)
    for system in cases:
        pos = system.pos + 0.3 * rng.standard_normal(system.pos.shape)
        out = []
        for scatter in (False, True):
            part = ForcePartMechanical(system, model=model, structured=False, scatter=scatter)
            mmf = MicMecForceField(system, [part])
            mmf.update_pos(pos)
            g, v = np.zeros(pos.shape), np.zeros((3, 3))
            out.append((mmf.compute(g, v), g, v))
        assert out[0][0] == out[1][0] and np.array_equal(out[0][2], out[1][2])  # energy / virial: same kernel code
        assert gio.rel_rms(out[1][1], out[0][1]) <= 1e-13


@pytest.mark.parametrize("model", MODELS)
def test_replica_batch_matches_per_system_oracle(model):
    """Config 5 in small: defect configurations x perturbations x strained cells evaluated as ONE batch."""
    from micmec_b200.replicas import ReplicaBatch

    names = ["3x3x3_conf0", "3x3x3_conf3", "3x3x3_conf9"]
    base = [make_system(gio.load("force_" + n)) for n in names]
    rng = np.random.default_rng(21)
    systems, pos, rvecs = [], [], []
    for rep in range(12):
        s = base[rep % 3]
        strain = np.eye(3) + 0.01 * rng.standard_normal((3, 3))
        systems.append(s)
        pos.append((s.pos + 0.3 * rng.standard_normal(s.pos.shape)) @ strain)
        rvecs.append(np.array(s.domain.rvecs) @ strain)
    batch = ReplicaBatch(systems, model=model)
    e, g, v = batch.compute(np.array(pos), np.array(rvecs))
    assert e.shape == (12,) and g.shape == (12, 27, 3) and v.shape == (12, 3, 3)
    for rep in range(12):
        o = orc.Oracle(base[rep % 3], model=model)
        eo, go, vo = o.compute(pos[rep], rvecs[rep], gpos=True, vtens=True)
        _, gc, vc = o.deformation(pos[rep], rvecs[rep])
        check_against(e[rep], g[rep], v[rep], eo, go, vo, gio.virial_noise(gc, vc))
    e2, g2, v2 = batch.compute(np.array(pos), np.array(rvecs), gpos=False)
    assert g2 is None and np.array_equal(e2, e) and np.array_equal(v2, v)


def open_grid_system(shape, type_params, hollow=()):
    """Finite (non-periodic) grid of one cell type, conventions of micmec/utils.py:164-263 with pbc = [False] * 3: nodes
    (nx+1)(ny+1)(nz+1) in C order (only those touched by a cell), cells in C order, -1 for a missing neighbour cell."""
    from micmec_b200.system import System

    nx, ny, nz = shape
    grid = np.ones(shape, dtype=np.int64)
    for idx in hollow:
        grid[idx] = 0
    offsets = np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)])
    node_id, nodes = {}, []
    for k in range(nx + 1):
        for l in range(ny + 1):
            for m in range(nz + 1):
                touched = any(0 <= k - d[0] < nx and 0 <= l - d[1] < ny and 0 <= m - d[2] < nz and grid[k - d[0], l - d[1], m - d[2]]
                              for d in offsets)
                if touched:
                    node_id[(k, l, m)] = len(nodes)
                    nodes.append((k, l, m))
    cells = [tuple(c) for c in np.argwhere(grid != 0)]
    cell_id = {c: n for n, c in enumerate(cells)}
    sn = np.array([[node_id[(c[0] + d[0], c[1] + d[1], c[2] + d[2])] for d in offsets] for c in cells], dtype=np.int64)
    sc = np.array([[cell_id.get((n[0] - d[0], n[1] - d[1], n[2] - d[2]), -1) for d in offsets] for n in nodes], dtype=np.int64)
    h0 = np.asarray(type_params["cell"], dtype=float).reshape(-1, 3, 3)[0]
    pos = np.array(nodes, dtype=float) * np.diag(h0)
    masses = np.array([0.125 * float(type_params["mass"]) * np.sum(row >= 0) for row in sc])
    params = {"type1/" + key: type_params[key] for key in ("cell", "elasticity", "free_energy", "effective_temp", "mass")}
    return System(pos, masses, np.zeros((0, 3)), sc, sn, grid=grid, types=np.ones(len(cells), dtype=np.int64), params=params)


@pytest.mark.parametrize("model", MODELS)
def test_finite_system_without_periodic_images(model):
    """pbc = False (mmff.py:262-263: no minimum-image shifts; Domain with nvec = 0, domain.c:23-25).  The reference's own
    `deformation` cannot evaluate such a system (it indexes rvecs[0] of an empty array, mmff.py:349), so the parity target
    is the sum of the reference's PER-CELL functions (pinned through the oracle's cell_state) over the cells, gathered in
    the reference's order - which is exactly what `deformation` would compute with zero shifts."""
    from micmec_b200.celltypes import TYPE_FCU

    system = open_grid_system((3, 2, 2), TYPE_FCU, hollow=[(1, 1, 1)])
    assert system.domain.nvec == 0 and system.domain.volume == 0.0
    rng = np.random.default_rng(12)
    system.pos[:] = system.pos + 0.4 * rng.standard_normal(system.pos.shape)
    mmf, part = make_mmf(system, model)
    assert not part.pbc
    g, v = np.zeros(system.pos.shape), np.zeros((3, 3))
    e = mmf.compute(g, v)
    h0 = np.asarray(TYPE_FCU["cell"]).reshape(-1, 3, 3)[0]
    C = np.asarray(TYPE_FCU["elasticity"]).reshape(-1, 3, 3, 3, 3)[0]
    eref, gref, vref = 0.0, np.zeros_like(g), np.zeros((3, 3))
    for c, verts in enumerate(system.surrounding_nodes):
        ec, gc = orc.Oracle.cell_state(model, system.pos[verts], h0, C)
        eref += ec
        gref[verts] += gc
        vref += gc.T @ system.pos[verts]
    assert abs(e - eref) <= ETOL * abs(eref)
    assert gio.rel_rms(g, gref) <= GTOL
    assert gio.rel_rms(v, vref) <= 1e-8
    assert abs(g.sum(axis=0)).max() <= 1e-12 * np.abs(g).max()  # a free cluster feels no net force
    # and the oracle's whole-system path agrees with zero shifts
    oracle = orc.Oracle(system, model=model, nthreads=2)
    eo, go, _ = oracle.compute(system.pos, system.domain.rvecs, gpos=True, vtens=False)
    assert abs(e - eo) <= ETOL * abs(eo) and gio.rel_rms(g, go) <= GTOL
