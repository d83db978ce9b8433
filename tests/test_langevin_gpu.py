"""Device-resident Langevin thermostat (reference: micmec/sampling/nvt.py:165-218).

The device generator (Philox4x32-10) is not NumPy's, so parity with the reference is statistical: the canonical
temperature, the bookkeeping of the conserved quantity, and - because the noise is keyed by the reference node id -
exact agreement between the structured (marching) and the indexed kernels and between runs of one seed.
The host-driven mode (seeded NumPy stream, reference trajectory reproduced) is covered by tests/test_hooks_gpu.py.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def build(shape, structured, seed, device=True, amp=0.2):
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import LangevinThermostat
    from micmec_b200.units import femtosecond, boltzmann

    rng = np.random.default_rng(4)
    system = System.periodic_grid(shape, TYPE_FCU, explicit=not structured)
    system.pos[:] = system.pos + amp * rng.standard_normal(system.pos.shape)
    part = ForcePartMechanical(system, structured=structured)
    assert part.structured == structured
    mmf = MicMecForceField(system, [part])
    vel0 = rng.standard_normal(system.pos.shape) * np.sqrt(boltzmann * 300.0 / system.masses)[:, None]
    np.random.seed(seed)
    thermo = LangevinThermostat(300.0, timecon=100 * femtosecond, device=device)
    verlet = VerletIntegrator(mmf, timestep=10 * femtosecond, hooks=[thermo], vel0=vel0)
    return verlet, thermo


def test_device_langevin_is_selected_and_reproducible():
    a, ta = build((12, 10, 8), True, seed=7)
    assert a.device_mode and ta.seed is not None
    b, _ = build((12, 10, 8), True, seed=7)
    c, _ = build((12, 10, 8), True, seed=8)
    for v in (a, b, c):
        v.run(25)
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.vel, b.vel)
    assert np.max(np.abs(a.vel - c.vel)) > 1e-3 * np.sqrt(np.mean(a.vel ** 2))
    # host-driven mode stays available (and is the default for small systems)
    d, td = build((12, 10, 8), True, seed=7, device=None)
    assert not d.device_mode


@pytest.mark.parametrize("shape", [(12, 10, 8), (37, 11, 9)])
def test_structured_and_indexed_kernels_see_the_same_noise(shape):
    """The kick of a node is a function of (seed, reference node id, half-step): ghost copies stay consistent without
    any exchange, and the two kernel families integrate the same stochastic trajectory."""
    a, _ = build(shape, True, seed=3)
    b, _ = build(shape, False, seed=3)
    for n in (1, 6, 20):
        a.run(n)
        b.run(n)
        scale = np.sqrt(np.mean(b.vel ** 2))
        assert np.max(np.abs(a.vel - b.vel)) <= 1e-9 * scale
        assert np.max(np.abs(a.pos - b.pos)) <= 1e-9 * np.sqrt(np.mean(b.pos ** 2))
        assert abs(a.econs - b.econs) <= 1e-9 * abs(b.econs)
        assert abs(a.ekin - b.ekin) <= 1e-9 * b.ekin


def test_canonical_temperature_and_conserved_quantity():
    """Equilibrium of the Ornstein-Uhlenbeck update: <T> = 300 K (3840 nodes, 600 steps, time constant 10 steps: the
    window mean has ~0.3 % noise), velocity components are normal with variance kB T / m, and
    econs = etot + sum(ekin before - ekin after) drifts only by the integrator's error."""
    from micmec_b200.units import boltzmann

    verlet, thermo = build((16, 16, 15), True, seed=11)
    verlet.run(100)
    econs0 = verlet.econs
    temps = []
    for _ in range(50):
        verlet.run(10)
        temps.append(verlet.temp)
    assert abs(np.mean(temps) - 300.0) <= 0.02 * 300.0, np.mean(temps)
    z = verlet.vel * np.sqrt(verlet.masses[:, None] / (boltzmann * 300.0))
    assert abs(z.mean()) <= 0.05 and abs(z.var() - 1.0) <= 0.06
    assert abs(np.mean(z ** 4) - 3.0) <= 0.25  # normal kurtosis
    assert abs(verlet.econs - econs0) <= 2e-2 * verlet.ekin, (verlet.econs, econs0, verlet.ekin)
    assert abs(thermo.econs_correction) > 0.0


def test_device_csvr_thermostat_samples_the_canonical_kinetic_energy():
    """CSVRThermostat(device=True): the stochastic velocity scale is drawn on the device (Philox; chi-square by
    Marsaglia-Tsang).  Canonical ensemble of the kinetic energy: <T> = 300 K and var(T) / <T>^2 = 2 / ndof; the conserved
    quantity (etot + sum (1 - alpha^2) ekin) drifts only by the integrator's error; one seed -> one trajectory."""
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
    from micmec_b200.sampling.verlet import VerletIntegrator
    from micmec_b200.sampling.nvt import CSVRThermostat
    from micmec_b200.units import femtosecond, boltzmann

    def build(seed):
        rng = np.random.default_rng(4)
        system = System.periodic_grid((16, 16, 15), TYPE_FCU, explicit=False)
        system.pos[:] = system.pos + 0.2 * rng.standard_normal(system.pos.shape)
        mmf = MicMecForceField(system, [ForcePartMechanical(system, structured=True)])
        vel0 = rng.standard_normal(system.pos.shape) * np.sqrt(boltzmann * 300.0 / system.masses)[:, None]
        np.random.seed(seed)
        thermo = CSVRThermostat(300.0, timecon=50 * femtosecond, device=True)
        verlet = VerletIntegrator(mmf, timestep=10 * femtosecond, hooks=[thermo], vel0=vel0)
        assert verlet.device_mode
        return verlet, thermo

    a, ta = build(5)
    b, _ = build(5)
    a.run(20)
    b.run(20)
    assert np.array_equal(a.vel, b.vel) and ta.seed is not None
    a.run(80)
    econs0 = a.econs
    temps = []
    for _ in range(150):
        a.run(8)
        temps.append(a.temp)
    temps = np.array(temps)
    ndof = float(a.ndof)
    assert abs(temps.mean() - 300.0) <= 0.01 * 300.0, temps.mean()
    ratio = temps.std() / (300.0 * np.sqrt(2.0 / ndof))
    assert 0.6 <= ratio <= 1.6, ratio
    assert abs(a.econs - econs0) <= 2e-2 * a.ekin, (a.econs, econs0, a.ekin)
    assert abs(ta.econs_correction) > 0.0
