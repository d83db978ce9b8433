"""GPU parity of the device-resident integrator (Verlet + NHC + MTK) with the reference's VerletIntegrator.

Golden trajectories come from the UNMODIFIED reference (tests/golden/make_golden.py); the CUDA integrator is
started from the recorded initial state (positions after domain symmetrisation, velocities, chain and barostat
velocities) and compared at steps 0, 1, 2, 10, 50 and 100.  north_star tolerance: 1e-8 over 100 steps.
"""
import numpy as np
import pytest

import goldenio as gio
from test_force_gpu import make_system

pytestmark = pytest.mark.gpu


def build(d, hooks_extra=(), model=None):
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import NHCThermostat
    from micmec_b200.sampling.npt import MTKBarostat, TBCombination

    system = make_system(d)
    system.pos[:] = d["init:pos"]
    system.domain.update_rvecs(np.ascontiguousarray(d["init:rvecs"]))
    mmf = MicMecForceField(system, [ForcePartMechanical(system, model=model or str(d["meta:model"]))])
    ens = str(d["meta:ensemble"])
    thermo = baro = None
    if ens in ("nvt", "npt"):
        thermo = NHCThermostat(float(d["meta:temp"]), timecon=float(d["meta:timecon_thermo"]),
                               chainlength=int(d["meta:chainlength"]), chain_pos0=d["step0:chain_pos"],
                               chain_vel0=d["step0:chain_vel"], restart=True)
    if ens in ("npt", "nph"):
        baro = MTKBarostat(mmf, float(d["meta:temp"]), float(d["meta:press"]), timecon=float(d["meta:timecon_baro"]),
                           anisotropic=bool(d["meta:anisotropic"]), vol_constraint=bool(d["meta:vol_constraint"]),
                           vel_press0=np.array(d["step0:vel_press"]) if bool(d["meta:anisotropic"])
                           else float(d["step0:vel_press"]), restart=True)
    hooks = list(hooks_extra)
    if thermo is not None and baro is not None:
        hooks.append(TBCombination(thermo, baro))
    elif thermo is not None:
        hooks.append(thermo)
    elif baro is not None:
        hooks.append(baro)
    verlet = VerletIntegrator(mmf, timestep=float(d["meta:timestep"]), hooks=hooks, vel0=d["init:vel"])
    return verlet, thermo, baro


@pytest.mark.parametrize("name", gio.traj_fixtures())
def test_trajectory_matches_reference_golden(name):
    d = gio.load("traj_" + name)
    verlet, thermo, baro = build(d)
    assert verlet.device_mode
    assert float(verlet.ndof) == float(d["meta:ndof"])
    done = 0
    for counter in [int(c) for c in d["meta:counters"]]:
        tol = 1e-10 if counter <= 2 else 1e-8
        xtol = tol if counter <= 10 else 1e-7  # chain / barostat variables of tiny systems, see test_oracle_golden
        verlet.run(counter - done)
        done = counter
        assert verlet.counter == counter
        p = "step%d:" % counter
        assert gio.rel_rms(verlet.pos, d[p + "pos"]) <= tol, counter
        assert gio.rel_rms(verlet.vel, d[p + "vel"]) <= tol, counter
        assert gio.rel_rms(verlet.gpos, d[p + "gpos"]) <= max(tol, 1e-9), counter
        assert gio.rel_rms(verlet.mmf.system.pos, d[p + "pos"]) <= tol
        assert gio.rel_rms(np.array(verlet.mmf.system.domain.rvecs), d[p + "rvecs"]) <= tol
        for key in ("epot", "ekin", "etot", "econs", "temp", "rmsd_gpos", "rmsd_delta", "time"):
            ref = float(d[p + key])
            ktol = max(tol, 1e-9) if key == "rmsd_delta" else tol
            assert abs(getattr(verlet, key) - ref) <= ktol * max(abs(ref), 1e-3), (counter, key, getattr(verlet, key), ref)
        if counter > 0:
            vtol = max(tol, 1e-9)
            assert gio.rel_rms(verlet.vtens, d[p + "vtens"]) <= vtol, counter
            assert gio.rel_rms(verlet.ptens, d[p + "ptens"]) <= vtol, counter
            assert abs(verlet.press - float(d[p + "press"])) <= vtol * np.sqrt(np.mean(d[p + "ptens"] ** 2))
        if counter >= 2:
            ref = float(d[p + "cons_err"])
            assert abs(verlet.cons_err - ref) <= 1e-5 * max(abs(ref), 1.0), (counter, verlet.cons_err, ref)
        if thermo is not None:
            assert gio.rel_rms(thermo.chain.vel, d[p + "chain_vel"]) <= xtol
            assert np.max(np.abs(thermo.chain.pos - d[p + "chain_pos"])) <= xtol
        if baro is not None:
            assert gio.rel_rms(np.asarray(baro.vel_press), d[p + "vel_press"]) <= max(xtol, 1e-9)


@pytest.mark.parametrize("structured", [False, True])
def test_matches_oracle_md_on_larger_grid(structured):
    """NVE + NVT + NPT on a perturbed 6x5x4 fcu grid: CUDA integrator (indexed kernels / structured-grid fused
    kernels) vs the CPU oracle's MD, 40 steps."""
    from oracle import oracle as orc
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import NHCThermostat
    from micmec_b200.sampling.npt import MTKBarostat, TBCombination
    from micmec_b200.units import femtosecond, pascal

    rng = np.random.default_rng(8)
    for ens in ("nve", "nvt", "npt"):
        system = System.periodic_grid((6, 5, 4), TYPE_FCU, explicit=True)
        system.pos[:] = system.pos + 0.2 * rng.standard_normal(system.pos.shape)
        sysx = system
        if structured:
            sysx = System.periodic_grid((6, 5, 4), TYPE_FCU, explicit=False)
            sysx.pos[:] = system.pos
        mmf = MicMecForceField(sysx, [ForcePartMechanical(sysx, structured=structured)])
        vel0 = 1e-5 * rng.standard_normal(system.pos.shape)
        vel0 -= vel0.mean(axis=0)
        cvel = np.array([1e-4, -2e-4, 5e-5])
        vp0 = 1e-6 * np.array([[1.0, 0.2, -0.1], [0.2, -0.5, 0.3], [-0.1, 0.3, 0.8]])
        thermo = NHCThermostat(300.0, timecon=100 * femtosecond, chain_vel0=cvel, chain_pos0=np.zeros(3), restart=True)
        hooks = [thermo] if ens != "nve" else []
        baro_kw = None
        if ens == "npt":
            baro = MTKBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e5 * femtosecond, vel_press0=vp0, restart=True)
            hooks = [TBCombination(thermo, baro)]
            baro_kw = dict(temp=300.0, press=1e6 * pascal, timecon=1e5 * femtosecond, vel_press0=vp0)
        verlet = VerletIntegrator(mmf, timestep=10 * femtosecond, hooks=hooks, vel0=vel0)
        o = orc.Oracle(system, nthreads=2)
        md = o.md(system.pos, vel0, system.masses, np.array(system.domain.rvecs), 10 * femtosecond,
                  thermo=dict(temp=300.0, timecon=100 * femtosecond, chain_vel0=cvel, chain_pos0=np.zeros(3))
                  if ens != "nve" else None, baro=baro_kw)
        for _ in range(4):
            verlet.run(10)
            md.run(10)
            assert gio.rel_rms(verlet.pos, md.pos) <= 1e-9
            assert gio.rel_rms(verlet.vel, md.vel) <= 1e-8
            assert abs(verlet.econs - md.econs) <= 1e-8 * abs(md.econs)
            assert abs(verlet.temp - md.temp) <= 1e-8 * md.temp
            assert gio.rel_rms(verlet.gpos, md.gpos) <= 1e-8
            assert abs(verlet.rmsd_delta - md.rmsd_delta) <= 1e-8 * md.rmsd_delta
            assert abs(verlet.rmsd_gpos - md.rmsd_gpos) <= 1e-8 * md.rmsd_gpos
            assert abs(verlet.press - md.press) <= 1e-7 * abs(md.press)
            if ens != "nve":
                assert gio.rel_rms(thermo.chain.vel, md.chain_vel) <= 1e-7


@pytest.mark.parametrize("path,model", [("march2", "original"), ("march2_images", "original"), ("march2_tail_in_kernel", "original"),
                                        ("general", "original"), ("march2", "default"), ("march2_images", "default")])
@pytest.mark.parametrize("ens", ["nve", "npt"])
def test_structured_md_multitile_matches_indexed(ens, path, model, monkeypatch):
    """Fused marching kernels vs the indexed kernels on a grid that needs several tiles in x (partial last tile, odd
    pitch), partial tiles in y and more than one chunk along z: 12 steps.  Paths: k_march2 with ghost nodes (the
    single-GPU default), with the periodic images taken on load + tail launch (the slab default), with the tail inside
    the marching kernel, and the general kernel k_march; both per-cell models."""
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import NHCThermostat
    from micmec_b200.sampling.npt import MTKBarostat, TBCombination
    from micmec_b200.units import femtosecond, pascal

    env = {"march2": {"MICMEC_B200_WRAP_ON_LOAD": "0"}, "march2_images": {"MICMEC_B200_WRAP_ON_LOAD": "1"},
           "march2_tail_in_kernel": {"MICMEC_B200_WRAP_ON_LOAD": "1", "MICMEC_B200_TAIL_IN_KERNEL": "1"},
           "general": {"MICMEC_B200_MARCH2": "0"}}[path]
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    shape = (37, 11, 41)
    rng = np.random.default_rng(21)
    base = System.periodic_grid(shape, TYPE_FCU, explicit=True)
    pos0 = base.pos + 0.2 * rng.standard_normal(base.pos.shape)
    vel0 = 1e-5 * rng.standard_normal(base.pos.shape)
    vel0 -= vel0.mean(axis=0)
    runs = {}
    for structured in (False, True):
        system = base if not structured else System.periodic_grid(shape, TYPE_FCU, explicit=False)
        system.pos[:] = pos0
        part = ForcePartMechanical(system, model=model, structured=structured)
        assert part.structured == structured
        mmf = MicMecForceField(system, [part])
        hooks = []
        if ens == "npt":
            thermo = NHCThermostat(300.0, timecon=100 * femtosecond, chain_vel0=np.array([1e-4, -2e-4, 5e-5]),
                                   chain_pos0=np.zeros(3), restart=True)
            baro = MTKBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e5 * femtosecond,
                               vel_press0=1e-6 * np.array([[1.0, 0.2, -0.1], [0.2, -0.5, 0.3], [-0.1, 0.3, 0.8]]), restart=True)
            hooks = [TBCombination(thermo, baro)]
        verlet = VerletIntegrator(mmf, timestep=10 * femtosecond, hooks=hooks, vel0=vel0.copy())
        verlet.run(12)
        runs[structured] = (verlet.pos.copy(), verlet.vel.copy(), verlet.gpos.copy(), verlet.econs, verlet.temp, verlet.press,
                            np.array(verlet.mmf.system.domain.rvecs))
    a, b = runs[False], runs[True]
    assert gio.rel_rms(b[0], a[0]) <= 1e-11
    assert gio.rel_rms(b[1], a[1]) <= 1e-9
    assert gio.rel_rms(b[2], a[2]) <= 1e-9
    assert abs(b[3] - a[3]) <= 1e-10 * abs(a[3])
    assert abs(b[4] - a[4]) <= 1e-10 * a[4]
    assert abs(b[5] - a[5]) <= 1e-8 * abs(a[5])
    assert gio.rel_rms(b[6], a[6]) <= 1e-12


class CountingHook(object):
    """A conventional hook as a user of the reference would write it."""

    def __init__(self, start=0, step=1):
        self.start, self.step, self.seen = start, step, []

    def expects_call(self, counter):
        return counter >= self.start and (counter - self.start) % self.step == 0

    def __call__(self, iterative):
        self.seen.append((iterative.counter, iterative.state["pos"].value.copy(), iterative.state["vel"].value.copy(),
                          float(iterative.state["epot"].value), float(iterative.state["volume"].value)))


def test_hook_protocol_and_lazy_sync():
    from micmec_b200.sampling.verlet import VerletScreenLog

    d = gio.load("traj_nvt_5x5x5_fcu_hollow")
    every10, log = CountingHook(step=10), VerletScreenLog(step=5)
    verlet, thermo, _ = build(d, hooks_extra=[every10, log])
    verlet.run(50)
    assert [c for c, *_ in every10.seen] == [0, 10, 20, 30, 40, 50]
    assert log.lines == 11
    c, pos, vel, epot, vol = every10.seen[-1]
    assert gio.rel_rms(pos, d["step50:pos"]) <= 1e-8 and gio.rel_rms(vel, d["step50:vel"]) <= 1e-8
    assert abs(epot - float(d["step50:epot"])) <= 1e-8 * abs(float(d["step50:epot"]))
    c, pos, vel, epot, vol = every10.seen[1]
    assert gio.rel_rms(pos, d["step10:pos"]) <= 1e-9
    assert verlet.state["cons_err"].value == verlet.cons_err
    # single-step API
    verlet.propagate()
    assert verlet.counter == 51


def test_host_driven_mode_for_foreign_verlet_hooks():
    """A VerletHook the library does not know forces the reference-style host loop; forces still come from the GPU."""
    from micmec_b200.sampling.verlet import VerletHook

    class Passive(VerletHook):
        calls = 0

        def init(self, iterative):
            pass

        def pre(self, iterative):
            Passive.calls += 1

        def post(self, iterative):
            pass

    d = gio.load("traj_nve_3x3x3_test")
    verlet, _, _ = build(d, hooks_extra=[Passive()])
    assert not verlet.device_mode
    verlet.run(10)
    assert Passive.calls == 10 and verlet.counter == 10
    assert gio.rel_rms(verlet.pos, d["step10:pos"]) <= 1e-9
    assert gio.rel_rms(verlet.vel, d["step10:vel"]) <= 1e-9
    assert abs(verlet.econs - float(d["step10:econs"])) <= 1e-9 * abs(float(d["step10:econs"]))


def test_restart_continues_time_and_counter_in_device_mode():
    """VerletIntegrator(time0=..., counter0=...) (verlet.py:96, iterative.py): a restarted device-resident run continues
    the clock and the step counter instead of starting from zero, and its tracker / part energies are mirrored."""
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.iterative import ConsErrStateItem, EPotContribStateItem

    d = gio.load("traj_" + gio.traj_fixtures()[0])
    system = make_system(d)
    system.pos[:] = d["init:pos"]
    mmf = MicMecForceField(system, [ForcePartMechanical(system)])
    dt = float(d["meta:timestep"])
    verlet = VerletIntegrator(mmf, timestep=dt, vel0=d["init:vel"], time0=123.0 * dt, counter0=123)
    assert verlet.device_mode
    assert verlet.time == 123.0 * dt and verlet.counter == 123
    verlet.run(7)
    assert verlet.counter == 130
    assert abs(verlet.time - 130.0 * dt) <= 1e-12 * 130.0 * dt
    assert ConsErrStateItem("counter").get_value(verlet) == 8  # initial sample + 7 steps
    assert ConsErrStateItem("ekin_m").get_value(verlet) > 0.0
    contribs = EPotContribStateItem().get_value(verlet)
    assert contribs.shape == (1,) and abs(contribs[0] - verlet.epot) <= 1e-12 * abs(verlet.epot)


def test_device_raw_writer_streams_frames_without_host_mirrors(tmp_path):
    """DeviceRawWriter: frames exported from the device layout into pinned buffers and written by a
    background thread equal what the integrator's host mirrors would have held; the mirrors are not refreshed for it."""
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import NHCThermostat
    from micmec_b200.sampling.trajectory import DeviceRawWriter, RawWriter, load_raw
    from micmec_b200.units import femtosecond

    def run(writer_cls, directory, **kw):
        rng = np.random.default_rng(2)
        system = System.periodic_grid((40, 24, 32), TYPE_FCU, explicit=False)
        system.pos[:] = system.pos + 0.2 * rng.standard_normal(system.pos.shape)
        mmf = MicMecForceField(system, [ForcePartMechanical(system, structured=True)])
        vel0 = 1e-5 * rng.standard_normal(system.pos.shape)
        writer = writer_cls(str(directory), start=0, step=3, **kw)
        thermo = NHCThermostat(300.0, timecon=100 * femtosecond, chain_vel0=np.array([1e-4, -2e-4, 5e-5]), chain_pos0=np.zeros(3), restart=True)
        verlet = VerletIntegrator(mmf, timestep=10 * femtosecond, hooks=[thermo, writer], vel0=vel0)
        verlet.run(10)
        if hasattr(writer, "close"):
            writer.close()
        return verlet

    a = run(DeviceRawWriter, tmp_path / "dev", fields=("pos", "vel"))
    b = run(RawWriter, tmp_path / "host", keys=("pos", "vel", "epot", "time"))
    da, db = load_raw(str(tmp_path / "dev")), load_raw(str(tmp_path / "host"))
    assert da["pos"].shape == db["pos"].shape == (4, 40 * 24 * 32, 3)  # counters 0, 3, 6, 9
    assert np.array_equal(np.asarray(da["pos"]), np.asarray(db["pos"]))
    assert np.array_equal(np.asarray(da["vel"]), np.asarray(db["vel"]))
    names = da["attrs"]["scalars"]
    assert np.array_equal(da["scalars"][:, names.index("epot")], np.asarray(db["epot"]))
    assert np.array_equal(da["scalars"][:, names.index("counter")], [0.0, 3.0, 6.0, 9.0])
    assert np.array_equal(a.pos, b.pos)  # the final state is synchronised as always
