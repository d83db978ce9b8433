"""Shared by test_hooks_cpu.py / test_hooks_gpu.py: the Verlet-hook cases of tests/golden/make_golden.py:HOOK_CASES
rebuilt from this package's classes, and the comparison with the reference's recorded trajectory."""
import numpy as np

import goldenio as gio

CASES = ["nvt_langevin", "npt_langevin", "npt_langevin_iso", "npt_berendsen", "nvt_berendsen", "nvt_csvr", "nvt_andersen",
         "nvt_gle", "npt_langevin_mtk", "nph_mtk_own_chain"]


def build_hooks(tag, mmf, dt):
    from micmec_b200.sampling import nvt, npt
    from micmec_b200.units import femtosecond, pascal

    if tag == "nvt_langevin":
        return [nvt.LangevinThermostat(300.0, timecon=100 * dt)]
    if tag == "npt_langevin":
        return [npt.TBCombination(nvt.LangevinThermostat(300.0, timecon=100 * dt),
                                  npt.LangevinBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e4 * dt))]
    if tag == "npt_langevin_iso":
        return [npt.TBCombination(nvt.LangevinThermostat(300.0, timecon=100 * dt),
                                  npt.LangevinBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e4 * dt, anisotropic=False))]
    if tag == "npt_berendsen":
        return [npt.TBCombination(nvt.BerendsenThermostat(300.0, timecon=100 * femtosecond),
                                  npt.BerendsenBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e5 * femtosecond))]
    if tag == "nvt_berendsen":
        return [nvt.BerendsenThermostat(250.0, timecon=200 * femtosecond)]
    if tag == "nvt_csvr":
        return [nvt.CSVRThermostat(300.0, timecon=100 * femtosecond)]
    if tag == "nvt_andersen":
        return [nvt.AndersenThermostat(300.0, annealing=0.99)]
    if tag == "nvt_gle":
        return [nvt.GLEThermostat(300.0, np.array([[2e-3, 1e-3], [-1e-3, 4e-3]]))]
    if tag == "npt_langevin_mtk":
        return [npt.TBCombination(nvt.LangevinThermostat(300.0, timecon=100 * dt),
                                  npt.MTKBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e5 * femtosecond))]
    if tag == "nph_mtk_own_chain":
        return [npt.MTKBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e5 * femtosecond,
                                baro_thermo=nvt.NHCThermostat(300.0, timecon=100 * femtosecond))]
    raise KeyError(tag)


def run_case(tag, make_part):
    """Seed the global RNG like the recording did, build the integrator from this package's classes on top of the
    force part ``make_part(system)`` returns, and compare with the reference at the recorded counters."""
    from test_force_gpu import make_system
    from micmec_b200.pes.mmff import MicMecForceField
    from micmec_b200.sampling.verlet import VerletIntegrator

    d = gio.load("hooks_" + tag)
    system = make_system(d)
    mmf = MicMecForceField(system, [make_part(system)])
    dt = float(d["meta:timestep"])
    np.random.seed(42)
    hooks = build_hooks(tag, mmf, dt)
    verlet = VerletIntegrator(mmf, timestep=dt, hooks=hooks, temp0=300.0)
    # these hooks act on host arrays with forces from the part - except the Berendsen thermostat on its own, whose single
    # deterministic velocity scale per step the CUDA integrator applies itself (exact parity with the recorded trajectory)
    on_device = tag == "nvt_berendsen" and type(mmf.parts[0]).__name__ == "ForcePartMechanical"
    assert verlet.device_mode == on_device
    assert float(verlet.ndof) == float(d["meta:ndof"])
    done = 0
    for counter in [int(c) for c in d["meta:counters"]]:
        tol = 1e-10 if counter <= 2 else 1e-8
        verlet.run(counter - done)
        done = counter
        p = "step%d:" % counter
        assert gio.rel_rms(verlet.pos, d[p + "pos"]) <= tol, (tag, counter)
        assert gio.rel_rms(verlet.vel, d[p + "vel"]) <= tol, (tag, counter)
        assert gio.rel_rms(verlet.gpos, d[p + "gpos"]) <= max(tol, 1e-9), (tag, counter)
        assert gio.rel_rms(np.array(mmf.system.domain.rvecs), d[p + "rvecs"]) <= tol, (tag, counter)
        for key in ("epot", "ekin", "etot", "econs", "temp", "rmsd_gpos", "time"):
            ref = float(d[p + key])
            assert abs(getattr(verlet, key) - ref) <= max(tol, 1e-9) * max(abs(ref), 1e-3), (tag, counter, key, getattr(verlet, key), ref)
        if counter > 0:
            assert gio.rel_rms(verlet.vtens, d[p + "vtens"]) <= max(tol, 1e-9), (tag, counter)
            assert abs(verlet.press - float(d[p + "press"])) <= max(tol, 1e-9) * np.sqrt(np.mean(d[p + "ptens"] ** 2))
    return verlet
