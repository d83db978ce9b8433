"""Shared by test_opt_cpu.py / test_opt_gpu.py: QNOptimizer runs of tests/golden/make_golden.py:OPT_CASES repeated
with this package's optimiser and DOF classes on top of a given force part."""
import numpy as np

import goldenio as gio

CASES = ["cartesian_3x3x3_conf0", "cartesian_5x5x5_fcu_hollow", "strain_3x3x3_test", "strain_frozen_3x3x3_conf3",
         "fullcell_2x2x2_reo"]


class Recorder(object):
    def __init__(self):
        self.rows = []

    def expects_call(self, counter):
        return True

    def __call__(self, it):
        self.rows.append(dict(x=it.x.copy(), f=float(it.f), g=it.g.copy(), trust_radius=float(it.trust_radius),
                              conv_val=float(it.dof.conv_val), conv_count=int(it.dof.conv_count)))


def build(tag, make_part):
    from test_force_gpu import make_system
    from micmec_b200.pes.mmff import MicMecForceField
    from micmec_b200.sampling.dof import CartesianDOF, StrainCellDOF, FullCellDOF

    d = gio.load("opt_" + tag)
    system = make_system(d)
    mmf = MicMecForceField(system, [make_part(system)])
    kwargs = {key[3:]: (bool(val) if val.dtype == bool else float(val)) for key, val in d.items() if key.startswith("kw:")}
    dof = {"cartesian": CartesianDOF, "strain": StrainCellDOF, "full": FullCellDOF}[str(d["meta:kind"])](mmf, **kwargs)
    return d, mmf, dof


def run_case(tag, make_part):
    """The optimiser path is compared iteration by iteration while the energy is well above the rounding floor of
    the reference's own evaluation, and by its end point afterwards.  Tolerances along the path are 1e-4: the secant
    iteration of solve_trust_radius stops at |error| < 1e-5 radius and its last update divides by a difference of
    two such errors, so the reference's own step carries ~1e-11 noise which the following iterations amplify to
    ~1e-6 before the run contracts again (measured: oracle-backed run vs recorded run).  The discrete decisions
    (trust radius, number of unmet criteria) must agree exactly."""
    from micmec_b200.sampling.opt import QNOptimizer

    d, mmf, dof = build(tag, make_part)
    rec = Recorder()
    opt = QNOptimizer(dof, hooks=[rec])
    niter = int(d["meta:iterations"])
    opt.run(niter + 10)
    assert dof.converged == bool(d["meta:converged"])
    f0 = float(d["iter0:f"])
    checked = 0
    for it in range(min(niter, len(rec.rows) - 1) + 1):
        p = "iter%d:" % it
        fref = float(d[p + "f"])
        if abs(fref) < 1e-7 * abs(f0):
            break  # below this the accept/shrink decisions depend on rounding in the reference itself
        row = rec.rows[it]
        xscale = np.sqrt(np.mean((d[p + "x"] - d["iter0:x"]) ** 2)) + np.sqrt(np.mean(d["iter0:g"] ** 2))
        assert np.max(np.abs(row["x"] - d[p + "x"])) <= 1e-4 * max(xscale, 1e-3), (tag, it)
        assert abs(row["f"] - fref) <= 1e-3 * abs(fref) + 1e-9 * abs(f0), (tag, it, row["f"], fref)
        assert gio.rel_rms(row["g"], d[p + "g"]) <= 2e-2, (tag, it)  # cell DOFs: d g / d x ~ 1e3 (dimensionless x)
        assert row["trust_radius"] == float(d[p + "trust_radius"]), (tag, it)
        assert row["conv_count"] == int(d[p + "conv_count"]), (tag, it)
        checked += 1
    assert checked >= 4, (tag, checked)
    last = "iter%d:" % niter
    assert abs(opt.counter - niter) <= 3, (tag, opt.counter, niter)
    assert gio.rel_rms(mmf.system.pos, d[last + "pos"]) <= 1e-5, tag
    assert gio.rel_rms(np.array(mmf.system.domain.rvecs), d[last + "rvecs"]) <= 1e-6, tag
    assert abs(opt.f - float(d[last + "f"])) <= 1e-8 * abs(f0), tag
    return opt
