"""Nose-Hoover chain thermostat (micmec/sampling/nvt.py:361-532) for the device-resident integrator.

``NHChain`` holds the chain state on the host (positions, velocities, masses) exactly like the reference; the
chain *propagation* of ``NHChain.__call__`` (nvt.py:410-451) runs inside the scalar kernel of libmicmec_b200.so and
this object is refreshed from the device whenever the host needs it.  ``NHChain.__call__`` is kept as a plain
NumPy method too, because the host-driven compatibility mode of ``VerletIntegrator`` (arbitrary Python hooks)
calls thermostats the way the reference does.
"""
import numpy as np

from ..units import boltzmann, femtosecond
from .iterative import StateItem
from .utils import clean_momenta, get_ndof_internal_md
from .verlet import VerletHook

__all__ = ["NHChain", "NHCThermostat", "NHCAttributeStateItem"]


class NHChain(object):
    def __init__(self, length, timestep, temp, ndof, pos0, vel0, timecon=100 * femtosecond):
        self.length = length
        self.timestep = timestep
        self.temp = temp
        self.timecon = timecon
        self.restart_pos = pos0 is not None
        self.restart_vel = vel0 is not None
        if ndof > 0:
            self.set_ndof(ndof)
        self.pos = np.array(pos0, dtype=float) if self.restart_pos else np.zeros(length)
        self.vel = np.array(vel0, dtype=float) if self.restart_vel else np.zeros(length)

    def set_ndof(self, ndof):
        """Chain masses Q_k = kT / omega^2, Q_0 *= ndof; random chain velocities unless restarted (nvt.py:393-408)."""
        self.ndof = ndof
        afreq = 2 * np.pi / self.timecon
        self.masses = np.ones(self.length) * (boltzmann * self.temp / afreq ** 2)
        self.masses[0] *= ndof
        if not self.restart_vel:
            self.vel = self.get_random_vel_therm()

    def get_random_vel_therm(self):
        return np.random.normal(0, np.sqrt(self.masses * boltzmann * self.temp), self.length) / self.masses

    def __call__(self, ekin, vel, G1_add):
        """Host version of the half-step chain propagation (only used in host-driven compatibility mode)."""
        dt, kt = self.timestep, self.temp * boltzmann

        def bead(k, ekin):
            if k == 0:
                g = 2 * ekin - self.ndof * kt + (G1_add if G1_add is not None else 0.0)
            else:
                g = self.masses[k - 1] * self.vel[k - 1] ** 2 - kt
            g /= self.masses[k]
            if k == self.length - 1:
                self.vel[k] += g * dt / 4
            else:
                damp = np.exp(-self.vel[k + 1] * dt / 8)
                self.vel[k] = (self.vel[k] * damp + g * dt / 4) * damp

        for k in range(self.length - 1, -1, -1):
            bead(k, ekin)
        self.pos += self.vel * dt / 2
        factor = np.exp(-self.vel[0] * dt / 2)
        vel *= factor
        ekin *= factor ** 2
        for k in range(self.length):
            bead(k, ekin)
        return vel, ekin

    def get_econs_correction(self):
        kt = boltzmann * self.temp
        return 0.5 * (self.vel ** 2 * self.masses).sum() + kt * (self.ndof * self.pos[0] + self.pos[1:].sum())


class NHCThermostat(VerletHook):
    name = "NHC"
    kind = "deterministic"
    method = "thermostat"
    native = True  # propagated on the device by libmicmec_b200.so

    def __init__(self, temp, start=0, timecon=100 * femtosecond, chainlength=3, chain_pos0=None, chain_vel0=None,
                 restart=False):
        self.temp = temp
        self.restart = restart
        self.chain = NHChain(chainlength, 0.0, temp, 0, chain_pos0, chain_vel0, timecon)
        VerletHook.__init__(self, start, 1)

    def init(self, iterative):
        if not self.restart:
            clean_momenta(iterative.pos, iterative.vel, iterative.masses, iterative.mmf.system.domain)
        if iterative.ndof is None:
            iterative.ndof = get_ndof_internal_md(iterative.pos.shape[0], iterative.mmf.system.domain.nvec)
        self.chain.timestep = iterative.timestep
        self.chain.set_ndof(iterative.ndof)

    def pre(self, iterative, G1_add=None):
        velnew, iterative.ekin = self.chain(iterative.ekin, iterative.vel, G1_add)
        iterative.vel[:] = velnew

    def post(self, iterative, G1_add=None):
        velnew, iterative.ekin = self.chain(iterative.ekin, iterative.vel, G1_add)
        iterative.vel[:] = velnew
        self.econs_correction = self.chain.get_econs_correction()


class NHCAttributeStateItem(StateItem):
    def __init__(self, attr):
        StateItem.__init__(self, "thermo_" + attr)
        self.attr = attr

    def get_value(self, iterative):
        from .npt import TBCombination

        for hook in iterative.hooks:
            if isinstance(hook, NHCThermostat):
                return getattr(hook.chain, self.attr)
            if isinstance(hook, TBCombination) and isinstance(hook.thermostat, NHCThermostat):
                return getattr(hook.thermostat.chain, self.attr)
        raise TypeError("Iterative does not contain a NHCThermostat hook.")

    def copy(self):
        return self.__class__(self.attr)
