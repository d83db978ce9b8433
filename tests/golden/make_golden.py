#!/usr/bin/env python
"""Record golden vectors from the UNMODIFIED reference (run in the build container only).

    make -C oracle ref && python tests/golden/make_golden.py

Imports /root/reference through ``refenv`` (stand-ins for the absent molmod/h5py, reference Domain
built into oracle/_ref/) and writes small ``.npz`` fixtures next to this script.  The fixtures carry
their own inputs (topology, parameters, positions) because /root/reference does not exist on the GPU box.

Contents
  cells.npz                 per-cell energy/gradient of both per-cell models on random vertices
                            (micmec/pes/nanocell_original.py, micmec/pes/nanocell.py) + the stencil tables
  force_<fixture>.npz       ForcePartMechanical.compute on data/<fixture>_micmec.chk: rest, seeded
                            perturbations, sheared domain; models original + default; the dense ``mic`` table
                            reduced to the (v0, vk) pairs that are looked up
  multistate.npz            synthetic 2- and 3-state types injected into 3x3x3_test / 3x3x3_conf0
  traj_<ensemble>.npz       100-step NVE / NVT(NHC) / NPT(NHC+MTK, aniso + iso) / NPH(MTK) trajectories of the
                            reference VerletIntegrator with recorded initial vel0 / chain / barostat state
  hooks_<case>.npz          30-step trajectories with the reference's other Verlet hooks (Langevin / Berendsen /
                            CSVR / Andersen / GLE thermostats, Langevin / Berendsen barostats, MTK with its own chain),
                            global NumPy RNG seeded with 42 right before the integrator is constructed: a replacement
                            that consumes the RNG in the same order reproduces them
  opt_<case>.npz            QNOptimizer runs (CartesianDOF, StrainCellDOF free / frozen, FullCellDOF): x, f, g, trust
                            radius and convergence bookkeeping after every iteration
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refenv  # noqa: E402

refenv.setup()

from micmec.system import System  # noqa: E402
from micmec.pes import mmff  # noqa: E402
from micmec.pes.mmff import MicMecForceField, ForcePartMechanical  # noqa: E402
from micmec.pes import nanocell, nanocell_original, nanocell_utils  # noqa: E402
from micmec.sampling.verlet import VerletIntegrator, VerletHook  # noqa: E402
from micmec.sampling.iterative import Hook  # noqa: E402
from micmec.sampling.nvt import (  # noqa: E402
    NHCThermostat, AndersenThermostat, BerendsenThermostat, LangevinThermostat, CSVRThermostat, GLEThermostat,
)
from micmec.sampling.npt import MTKBarostat, TBCombination, BerendsenBarostat, LangevinBarostat  # noqa: E402
from molmod.units import femtosecond, pascal  # noqa: E402

DATA = os.path.join(refenv.REFERENCE, "data")
FIXTURES = [
    "2x2x2_test", "2x2x2_fcu", "2x2x2_reo", "2x2x2_fcu_hollow", "3x3x3_test", "3x3x3_fcu_hollow",
    "3x3x3_conf0", "3x3x3_conf3", "3x3x3_conf9", "4x4x4_fcu", "4x4x4_fcu_hollow", "5x5x5_fcu_hollow",
]
MODELS = ["original", "default"]


def load(name):
    return System.from_file(os.path.join(DATA, name + "_micmec.chk"))


def system_arrays(system):
    """Everything needed to rebuild the system without the reference."""
    out = dict(
        pos=system.pos.copy(),
        masses=system.masses.copy(),
        rvecs=np.array(system.domain.rvecs),
        surrounding_cells=np.asarray(system.surrounding_cells, dtype=np.int64),
        surrounding_nodes=np.asarray(system.surrounding_nodes, dtype=np.int64),
        boundary_nodes=np.asarray(system.boundary_nodes, dtype=np.int64),
        grid=np.asarray(system.grid),
        types=np.asarray(system.types),
    )
    for key, val in system.params.items():
        if key.split("/")[1] in ("cell", "elasticity", "free_energy", "effective_temp", "mass"):
            out["params:" + key] = np.asarray(val, dtype=float)
    return out


def compute_all(system, pos, rvecs):
    """(E, gpos, vtens, epot_cells, gpos_cells) for both models at the given geometry."""
    res = {}
    for model in MODELS:
        refenv.use_model(model)
        fpm = ForcePartMechanical(system)
        mmf = MicMecForceField(system, [fpm])
        mmf.update_rvecs(np.ascontiguousarray(rvecs))
        mmf.update_pos(pos)
        gpos = np.zeros(pos.shape)
        vtens = np.zeros((3, 3))
        energy = mmf.compute(gpos, vtens)
        res[model] = dict(
            energy=energy, gpos=gpos, vtens=vtens, epot_cells=fpm.epot_cells.copy(), gpos_cells=fpm.gpos_cells.copy()
        )
    refenv.use_model("original")
    return res


def make_cells():
    rng = np.random.default_rng(1234)
    ncase = 12
    verts = np.zeros((ncase, 8, 3))
    h0 = np.zeros((ncase, 3, 3))
    C = np.zeros((ncase, 3, 3, 3, 3))
    out = {}
    base = np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)], float)
    for n in range(ncase):
        h0[n] = np.diag(rng.uniform(15.0, 40.0, 3)) + rng.normal(0.0, 0.8, (3, 3))
        verts[n] = base @ h0[n] + rng.normal(0.0, 0.8, (8, 3)) + rng.normal(0.0, 30.0, 3)
        c = rng.normal(0.0, 1e-3, (3, 3, 3, 3))  # generic tensor without any symmetry
        if n % 2 == 0:  # half of the cases get a physical-looking, fully symmetric tensor
            v = rng.normal(0.0, 1e-3, (6, 6))
            v = v @ v.T
            idx = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]
            c = np.zeros((3, 3, 3, 3))
            for a, (i, j) in enumerate(idx):
                for b, (k, l) in enumerate(idx):
                    c[i, j, k, l] = c[j, i, k, l] = c[i, j, l, k] = c[j, i, l, k] = v[a, b]
        C[n] = c
    for model, mod in (("original", nanocell_original), ("default", nanocell)):
        out["energy_" + model] = np.array([mod.elastic_energy_nanocell(verts[n], h0[n], C[n]) for n in range(ncase)])
        out["grad_" + model] = np.array([mod.grad_elastic_energy_nanocell(verts[n], h0[n], C[n]) for n in range(ncase)])
    out.update(verts=verts, h0=h0, C=C)
    out["multiplicator"] = np.asarray(nanocell_utils.multiplicator, dtype=float)
    out["cell_derivs"] = np.array(
        [nanocell_utils.cell_xderivs, nanocell_utils.cell_yderivs, nanocell_utils.cell_zderivs], dtype=float
    )
    np.savez_compressed(os.path.join(HERE, "cells.npz"), **out)


def make_force(name):
    system = load(name)
    out = system_arrays(system)
    fpm = ForcePartMechanical(system)
    sn = out["surrounding_nodes"]
    out["shift_ref"] = fpm.mic[sn[:, :1], sn, :].astype(np.int8)  # mic[v0, vk, :]
    pos0 = system.pos.copy()
    rvecs0 = np.array(system.domain.rvecs)
    shear = np.eye(3) + np.array([[0.02, 0.03, -0.01], [0.0, -0.015, 0.025], [0.0, 0.0, 0.01]])
    cases = {
        "rest": (pos0, rvecs0),
        "rng0": (pos0 + 0.5 * np.random.default_rng(0).standard_normal(pos0.shape), rvecs0),
        "rng1": (pos0 + 0.3 * np.random.default_rng(1).standard_normal(pos0.shape), rvecs0),
        "shear": ((pos0 + 0.4 * np.random.default_rng(2).standard_normal(pos0.shape)) @ shear, rvecs0 @ shear),
    }
    for case, (pos, rvecs) in cases.items():
        out["%s:pos" % case] = pos
        out["%s:rvecs" % case] = rvecs
        for model, res in compute_all(system, pos, rvecs).items():
            for key, val in res.items():
                if key in ("epot_cells", "gpos_cells") and case != "rng0":
                    continue
                out["%s:%s:%s" % (case, model, key)] = val
    np.savez_compressed(os.path.join(HERE, "force_%s.npz" % name), **out)


def make_multistate():
    out = {}
    for tag, name, nstate in (("a", "3x3x3_test", 2), ("b", "3x3x3_conf0", 3)):
        system = load(name)
        rng = np.random.default_rng(7)
        for t in sorted({int(x) for x in system.types}):
            h0 = np.asarray(system.params["type%d/cell" % t])[0]
            C0 = np.asarray(system.params["type%d/elasticity" % t])[0]
            cells, elas, free = [h0], [C0], [0.0]
            for s in range(1, nstate):
                cells.append(h0 * (1.0 + 0.012 * s) + rng.normal(0.0, 0.02, (3, 3)))
                elas.append(C0 * (1.0 - 0.2 * s))
                free.append(4.0e-4 * s)  # ~ 0.4 kT_eff at 300 K: states genuinely mix
            system.params["type%d/cell" % t] = np.array(cells)
            system.params["type%d/elasticity" % t] = np.array(elas)
            system.params["type%d/free_energy" % t] = np.array(free)
            system.params["type%d/effective_temp" % t] = 300.0 + 50.0 * t
        arrays = system_arrays(system)
        for key, val in arrays.items():
            out["%s:%s" % (tag, key)] = val
        pos0 = system.pos.copy()
        rvecs0 = np.array(system.domain.rvecs)
        for case, amp in (("small", 0.02), ("large", 0.4)):
            pos = pos0 + amp * np.random.default_rng(11).standard_normal(pos0.shape)
            out["%s:%s:pos" % (tag, case)] = pos
            out["%s:%s:rvecs" % (tag, case)] = rvecs0
            for model, res in compute_all(system, pos, rvecs0).items():
                for key, val in res.items():
                    out["%s:%s:%s:%s" % (tag, case, model, key)] = val
    np.savez_compressed(os.path.join(HERE, "multistate.npz"), **out)


class Recorder(Hook):
    """Conventional hook: snapshot the integrator at chosen counters."""

    KEYS = ["epot", "ekin", "etot", "econs", "cons_err", "temp", "press", "rmsd_gpos", "rmsd_delta", "time"]

    def __init__(self, counters, store, thermo=None, baro=None):
        Hook.__init__(self, 0, 1)
        self.counters, self.store, self.thermo, self.baro = set(counters), store, thermo, baro

    def __call__(self, it):
        if it.counter not in self.counters:
            return
        p = "step%d:" % it.counter
        s = self.store
        s[p + "pos"], s[p + "vel"], s[p + "gpos"] = it.pos.copy(), it.vel.copy(), it.gpos.copy()
        s[p + "rvecs"], s[p + "vtens"] = np.array(it.mmf.system.domain.rvecs), it.vtens.copy()
        s[p + "ptens"] = np.array(getattr(it, "ptens", np.zeros((3, 3))))
        for key in self.KEYS:
            s[p + key] = float(getattr(it, key))
        if self.thermo is not None:
            s[p + "chain_pos"], s[p + "chain_vel"] = self.thermo.chain.pos.copy(), self.thermo.chain.vel.copy()
        if self.baro is not None:
            s[p + "vel_press"] = np.array(self.baro.vel_press, dtype=float)


def make_traj(tag, name, ensemble, model="original", nsteps=100, amp=0.3, **opts):
    refenv.use_model(model)
    system = load(name)
    system.pos[:] = system.pos + amp * np.random.default_rng(5).standard_normal(system.pos.shape)
    out = system_arrays(system)
    fpm = ForcePartMechanical(system)
    mmf = MicMecForceField(system, [fpm])
    temp, press = 300.0, 1e6 * pascal
    timestep = 10.0 * femtosecond
    # NOTE: the class default (1000 fs, micmec/sampling/npt.py:525) makes the barostat mass so small for these
    # stiff coarse-grained cells that the domain oscillation period drops below the 10 fs timestep: the unmodified
    # reference then collapses the cell within two steps (T -> 0, epot x1000).  1e5 fs is stable.
    timecon_baro = opts.get("timecon_baro", 1.0e5) * femtosecond
    thermo = baro = None
    hooks = []
    if ensemble in ("nvt", "npt"):
        thermo = NHCThermostat(temp, timecon=100 * femtosecond, chainlength=opts.get("chainlength", 3))
    if ensemble in ("npt", "nph"):
        baro = MTKBarostat(
            mmf, temp, press, timecon=timecon_baro, anisotropic=opts.get("anisotropic", True),
            vol_constraint=opts.get("vol_constraint", False),
        )
    if thermo is not None and baro is not None:
        hooks.append(TBCombination(thermo, baro))
    elif thermo is not None:
        hooks.append(thermo)
    elif baro is not None:
        hooks.append(baro)
    counters = [0, 1, 2, 10, 50, nsteps]
    hooks.append(Recorder(counters, out, thermo, baro))
    np.random.seed(42)
    verlet = VerletIntegrator(mmf, timestep=timestep, hooks=hooks, temp0=temp)
    # the state the integrator actually starts from (after domain symmetrisation, momentum cleaning and the
    # random draws of the hooks' init): this is what the oracle and the CUDA integrator are started from.
    out["init:pos"], out["init:vel"] = out["step0:pos"], out["step0:vel"]
    out["init:rvecs"] = out["step0:rvecs"]
    out["meta:timestep"], out["meta:temp"], out["meta:press"] = timestep, temp, press
    out["meta:ndof"] = float(verlet.ndof)
    out["meta:timecon_thermo"], out["meta:timecon_baro"] = 100 * femtosecond, timecon_baro
    out["meta:anisotropic"] = int(opts.get("anisotropic", True))
    out["meta:vol_constraint"] = int(opts.get("vol_constraint", False))
    out["meta:chainlength"] = int(opts.get("chainlength", 3))
    out["meta:nsteps"] = nsteps
    if baro is not None:
        out["meta:mass_press"] = float(baro.mass_press)
    if thermo is not None:
        out["meta:chain_masses"] = thermo.chain.masses.copy()
    verlet.run(nsteps)
    out["meta:ensemble"] = np.array(ensemble)
    out["meta:model"] = np.array(model)
    out["meta:counters"] = np.array(counters)
    np.savez_compressed(os.path.join(HERE, "traj_%s.npz" % tag), **out)
    refenv.use_model("original")


HOOK_CASES = {
    # tag: (fixture, builder(mmf, timestep) -> list of Verlet hooks); time constants as in simulations/md.py:57-66
    # unless the class default is unstable for these stiff cells (see the note in make_traj)
    "nvt_langevin": ("3x3x3_test", lambda mmf, dt: [LangevinThermostat(300.0, timecon=100 * dt)]),
    "npt_langevin": ("3x3x3_conf0", lambda mmf, dt: [TBCombination(
        LangevinThermostat(300.0, timecon=100 * dt), LangevinBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e4 * dt))]),
    "npt_langevin_iso": ("2x2x2_fcu", lambda mmf, dt: [TBCombination(
        LangevinThermostat(300.0, timecon=100 * dt),
        LangevinBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e4 * dt, anisotropic=False))]),
    "npt_berendsen": ("3x3x3_test", lambda mmf, dt: [TBCombination(
        BerendsenThermostat(300.0, timecon=100 * femtosecond),
        BerendsenBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e5 * femtosecond))]),
    "nvt_berendsen": ("5x5x5_fcu_hollow", lambda mmf, dt: [BerendsenThermostat(250.0, timecon=200 * femtosecond)]),
    "nvt_csvr": ("3x3x3_conf3", lambda mmf, dt: [CSVRThermostat(300.0, timecon=100 * femtosecond)]),
    "nvt_andersen": ("2x2x2_test", lambda mmf, dt: [AndersenThermostat(300.0, annealing=0.99)]),
    "nvt_gle": ("2x2x2_reo", lambda mmf, dt: [GLEThermostat(300.0, np.array([[2e-3, 1e-3], [-1e-3, 4e-3]]))]),
    "npt_langevin_mtk": ("3x3x3_conf9", lambda mmf, dt: [TBCombination(
        LangevinThermostat(300.0, timecon=100 * dt), MTKBarostat(mmf, 300.0, 1e6 * pascal, timecon=1e5 * femtosecond))]),
    "nph_mtk_own_chain": ("3x3x3_test", lambda mmf, dt: [MTKBarostat(
        mmf, 300.0, 1e6 * pascal, timecon=1e5 * femtosecond, baro_thermo=NHCThermostat(300.0, timecon=100 * femtosecond))]),
}


def make_hook_traj(tag, nsteps=30, amp=0.3):
    name, builder = HOOK_CASES[tag]
    refenv.use_model("original")
    system = load(name)
    system.pos[:] = system.pos + amp * np.random.default_rng(5).standard_normal(system.pos.shape)
    out = system_arrays(system)  # geometry BEFORE any hook symmetrises the domain: the replacement starts from here
    fpm = ForcePartMechanical(system)
    mmf = MicMecForceField(system, [fpm])
    timestep = 10.0 * femtosecond
    np.random.seed(42)
    hooks = builder(mmf, timestep)
    counters = [0, 1, 2, 10, nsteps]
    hooks.append(Recorder(counters, out))
    verlet = VerletIntegrator(mmf, timestep=timestep, hooks=hooks, temp0=300.0)
    verlet.run(nsteps)
    for counter in counters:
        assert np.isfinite(out["step%d:epot" % counter]) and out["step%d:temp" % counter] < 3000.0, (tag, counter)
    out["meta:timestep"], out["meta:ndof"], out["meta:nsteps"] = timestep, float(verlet.ndof), nsteps
    out["meta:counters"] = np.array(counters)
    np.savez_compressed(os.path.join(HERE, "hooks_%s.npz" % tag), **out)


class OptRecorder(Hook):
    def __init__(self, store):
        Hook.__init__(self, 0, 1)
        self.store = store

    def __call__(self, it):
        p = "iter%d:" % it.counter
        self.store[p + "x"], self.store[p + "f"], self.store[p + "g"] = it.x.copy(), float(it.f), it.g.copy()
        self.store[p + "trust_radius"] = float(it.trust_radius)
        self.store[p + "conv_val"], self.store[p + "conv_count"] = float(it.dof.conv_val), int(it.dof.conv_count)
        self.store[p + "pos"] = it.mmf.system.pos.copy()
        self.store[p + "rvecs"] = np.array(it.mmf.system.domain.rvecs)


OPT_CASES = {
    # tag: (fixture, dof kind, dof kwargs, perturbation amplitude, cell strain, max iterations)
    "cartesian_3x3x3_conf0": ("3x3x3_conf0", "cartesian", dict(gpos_rms=1e-7, dpos_rms=1e-5), 0.5, None, 60),
    "cartesian_5x5x5_fcu_hollow": ("5x5x5_fcu_hollow", "cartesian", dict(), 0.4, None, 60),
    "strain_3x3x3_test": ("3x3x3_test", "strain",
                          dict(gpos_rms=1e-8, dpos_rms=1e-6, grvecs_rms=1e-8, drvecs_rms=1e-6), 0.3,
                          np.array([[1.02, 0.01, 0.0], [0.01, 0.98, -0.015], [0.0, -0.015, 1.01]]), 80),
    "strain_frozen_3x3x3_conf3": ("3x3x3_conf3", "strain", dict(do_frozen=True), 0.0,
                                  np.array([[0.97, 0.0, 0.02], [0.0, 1.03, 0.0], [0.02, 0.0, 1.0]]), 40),
    "fullcell_2x2x2_reo": ("2x2x2_reo", "full", dict(), 0.2,
                           np.array([[1.01, 0.0, 0.0], [0.0, 0.99, 0.01], [0.0, 0.01, 1.0]]), 60),
}


def make_opt(tag):
    from micmec.sampling.opt import QNOptimizer
    from micmec.sampling.dof import CartesianDOF, StrainCellDOF, FullCellDOF

    name, kind, kwargs, amp, strain, maxiter = OPT_CASES[tag]
    refenv.use_model("original")
    system = load(name)
    if strain is not None:
        system.domain.update_rvecs(np.ascontiguousarray(np.dot(np.array(system.domain.rvecs), strain)))
        system.pos[:] = np.dot(system.pos, strain)
    system.pos[:] = system.pos + amp * np.random.default_rng(9).standard_normal(system.pos.shape)
    out = system_arrays(system)
    mmf = MicMecForceField(system, [ForcePartMechanical(system)])
    dof = {"cartesian": CartesianDOF, "strain": StrainCellDOF, "full": FullCellDOF}[kind](mmf, **kwargs)
    opt = QNOptimizer(dof, hooks=[OptRecorder(out)])
    opt.run(maxiter)
    out["meta:iterations"], out["meta:converged"] = opt.counter, int(dof.converged)
    out["meta:kind"] = np.array(kind)
    for key, val in kwargs.items():
        out["kw:" + key] = val
    np.savez_compressed(os.path.join(HERE, "opt_%s.npz" % tag), **out)
    print(tag, "iterations", opt.counter, "converged", dof.converged, "f", opt.f)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "opt":
        for tag in OPT_CASES:
            make_opt(tag)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "hooks":
        for tag in HOOK_CASES:
            make_hook_traj(tag)
        return
    make_cells()
    for name in FIXTURES:
        make_force(name)
    make_multistate()
    make_traj("nve_3x3x3_test", "3x3x3_test", "nve")
    make_traj("nve_default_3x3x3_conf0", "3x3x3_conf0", "nve", model="default")
    make_traj("nvt_5x5x5_fcu_hollow", "5x5x5_fcu_hollow", "nvt")
    make_traj("npt_4x4x4_fcu", "4x4x4_fcu", "npt")
    make_traj("npt_iso_3x3x3_test", "3x3x3_test", "npt", anisotropic=False)
    make_traj("npt_volc_2x2x2_reo", "2x2x2_reo", "npt", vol_constraint=True)
    make_traj("nph_3x3x3_conf3", "3x3x3_conf3", "nph")
    for tag in HOOK_CASES:
        make_hook_traj(tag)
    for tag in OPT_CASES:
        make_opt(tag)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
