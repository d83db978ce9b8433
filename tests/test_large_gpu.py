"""GPU parity at the sizes of BASELINE.json configs[2] and configs[3] (64^3 and 256^3 cells) against the CPU oracle.

The oracle (oracle/micmec_oracle.c, OpenMP) evaluates a 256^3 grid in seconds, so the largest configuration is checked
directly - not only through invariants: one force evaluation and three NPT (NHC + MTK) steps of the synthetic fcu grid
of the benchmark (same initial state as bench.py), and 20 NVT steps at 64^3.
Tolerances (north_star): energy 1e-10 relative, forces / virial 1e-9 of RMS, positions / velocities 1e-9 of RMS after
the steps.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel_rms(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.sqrt(np.mean(np.asarray(b) ** 2)))


def run_case(grid, ensemble, nsteps, etol=1e-10):
    import bench
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import NHCThermostat
    from micmec_b200.sampling.npt import MTKBarostat, TBCombination
    from oracle import oracle as orc

    p = bench.md_params(ensemble)
    system, vel0 = bench.make_state(grid)
    part = ForcePartMechanical(system, model="original", device=0)
    assert part.structured
    mmf = MicMecForceField(system, [part])
    arrays = orc.periodic_grid_system((grid,) * 3, TYPE_FCU)[0]
    oracle = orc.Oracle(model="original", nthreads=os.cpu_count() or 1, **arrays)

    # ---- one force evaluation through the plugin API ----------------------------------------------------------------
    gpos, vtens = np.zeros(system.pos.shape), np.zeros((3, 3))
    energy = mmf.compute(gpos, vtens)
    eo, go, vo = oracle.compute(system.pos, system.domain.rvecs, gpos=True, vtens=True)
    assert abs(energy - eo) <= etol * abs(eo), (energy, eo)
    assert rel_rms(gpos, go) <= 1e-9
    assert rel_rms(vtens, vo) <= 1e-9
    del gpos, go

    # ---- MD steps: device-resident integrator vs the oracle's ---------------------------------------------------------
    thermo = NHCThermostat(p["temp"], timecon=p["timecon_thermo"], chain_vel0=p["chain_vel0"], chain_pos0=np.zeros(3), restart=True)
    hooks = [thermo]
    baro_kw = None
    if p["baro"]:
        baro = MTKBarostat(mmf, p["temp"], p["press"], timecon=p["timecon_baro"], vel_press0=p["vel_press0"], restart=True)
        hooks = [TBCombination(thermo, baro)]
        baro_kw = dict(temp=p["temp"], press=p["press"], timecon=p["timecon_baro"], vel_press0=p["vel_press0"])
    verlet = VerletIntegrator(mmf, timestep=p["timestep"], hooks=hooks, vel0=vel0)
    assert verlet.device_mode
    md = oracle.md(system.pos, vel0, system.masses, np.array(system.domain.rvecs), p["timestep"],
                   thermo=dict(temp=p["temp"], timecon=p["timecon_thermo"], chain_vel0=p["chain_vel0"], chain_pos0=np.zeros(3)),
                   baro=baro_kw)
    verlet.run(nsteps)
    md.run(nsteps)
    assert rel_rms(verlet.pos, md.pos) <= 1e-9
    assert rel_rms(verlet.vel, md.vel) <= 1e-9
    assert rel_rms(verlet.gpos, md.gpos) <= 1e-9
    assert rel_rms(np.array(verlet.rvecs), md.rvecs) <= 1e-10
    for key in ("epot", "ekin", "econs", "temp"):
        a, b = getattr(verlet, key), getattr(md, key)
        assert abs(a - b) <= 1e-9 * abs(b), (key, a, b)
    assert rel_rms(verlet.vtens, md.vtens) <= 1e-8
    assert rel_rms(thermo.chain.vel, md.chain_vel) <= 1e-8


def test_256_cubed_npt_force_and_three_steps_match_oracle():
    run_case(256, "npt", 3)


def test_64_cubed_nvt_twenty_steps_match_oracle():
    run_case(64, "nvt", 20)


def test_64_cubed_npt_twenty_steps_match_oracle():
    run_case(64, "npt", 20)
