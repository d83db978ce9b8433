"""MTK barostat and the thermostat-barostat combination (micmec/sampling/npt.py:52-177, 513-757).

In the device-resident integrator the whole of ``MTKBarostat.baro`` (npt.py:653-736) - barostat velocity update,
3x3 eigen-decompositions, position / cell / velocity rotations and the extra force evaluation - runs inside
libmicmec_b200.so; these classes carry the parameters and mirror the state (``vel_press``, ``mass_press``,
``econs_correction``).  A barostat with its own thermostat (``baro_thermo``) is not propagated on the device.
"""
import numpy as np

from ..units import boltzmann, femtosecond
from .iterative import StateItem
from .nvt import NHCThermostat
from .utils import clean_momenta, domain_symmetrize, get_ndof_baro, get_ndof_internal_md, get_random_vel_press
from .verlet import VerletHook

__all__ = ["TBCombination", "MTKBarostat", "MTKAttributeStateItem"]


class MTKBarostat(VerletHook):
    name = "MTTK"
    kind = "deterministic"
    method = "barostat"
    native = True

    def __init__(self, mmf, temp, press, start=0, step=1, timecon=1000 * femtosecond, anisotropic=True,
                 vol_constraint=False, baro_thermo=None, vel_press0=None, restart=False):
        if baro_thermo is not None:
            raise NotImplementedError(
                "MTKBarostat(baro_thermo=...) is outside the device-resident path; couple the barostat to the "
                "particle thermostat through TBCombination instead."
            )
        self.temp = temp
        self.press = press
        self.timecon_press = timecon
        self.anisotropic = anisotropic
        self.vol_constraint = vol_constraint
        self.baro_thermo = None
        self.dim = mmf.system.domain.nvec
        self.restart = restart
        self.baro_ndof = get_ndof_baro(self.dim, self.anisotropic, self.vol_constraint)
        if self.anisotropic and not self.restart:
            domain_symmetrize(mmf)  # npt.py:572-575
        self.domain = np.array(mmf.system.domain.rvecs)
        self.vel_press = vel_press0
        VerletHook.__init__(self, start, step)

    def init(self, iterative):
        """npt.py:579-614 minus the force re-evaluation, which the device integrator performs itself."""
        self.timestep_press = iterative.timestep
        if not self.restart:
            clean_momenta(iterative.pos, iterative.vel, iterative.masses, iterative.mmf.system.domain)
        if iterative.ndof is None:
            iterative.ndof = get_ndof_internal_md(iterative.mmf.system.nnodes, iterative.mmf.system.domain.nvec)
        angfreq = 2 * np.pi / self.timecon_press
        self.mass_press = (iterative.ndof + self.dim ** 2) * boltzmann * self.temp / angfreq ** 2
        if self.vel_press is None:
            self.vel_press = get_random_vel_press(self.mass_press, self.temp)
            if not self.anisotropic:
                self.vel_press = self.vel_press[0][0]
        if self.vol_constraint:
            self.vel_press = self.vel_press - np.trace(self.vel_press) / 3 * np.eye(3)

    def add_press_cont(self):
        return 2 * self._compute_ekin_baro() - self.baro_ndof * self.temp * boltzmann

    def _compute_ekin_baro(self):
        if self.anisotropic:
            return 0.5 * self.mass_press * np.trace(np.dot(self.vel_press.T, self.vel_press))
        return 0.5 * self.mass_press * self.vel_press ** 2


class TBCombination(VerletHook):
    name = "TBCombination"
    native = True

    def __init__(self, thermostat, barostat, start=0):
        self.thermostat = thermostat
        self.barostat = barostat
        self.start = start
        if not self.verify():
            self.thermostat, self.barostat = barostat, thermostat
            if not self.verify():
                raise TypeError("The Thermostat or Barostat instance is not supported (yet).")
        self.step_thermo = self.thermostat.step
        self.step_baro = self.barostat.step
        if self.step_thermo != 1 or self.step_baro != 1:
            raise NotImplementedError("the device integrator calls thermostat and barostat every step")
        VerletHook.__init__(self, start, 1)

    def init(self, iterative):
        self.thermostat.init(iterative)
        self.barostat.init(iterative)
        self.chainvel0 = None
        self.G1_add = None

    def verify(self):
        return isinstance(self.thermostat, NHCThermostat) and isinstance(self.barostat, MTKBarostat)


class MTKAttributeStateItem(StateItem):
    def __init__(self, attr):
        StateItem.__init__(self, "baro_" + attr)
        self.attr = attr

    def get_value(self, iterative):
        for hook in iterative.hooks:
            if isinstance(hook, MTKBarostat):
                return getattr(hook, self.attr)
            if isinstance(hook, TBCombination) and isinstance(hook.barostat, MTKBarostat):
                return getattr(hook.barostat, self.attr)
        raise TypeError("Iterative does not contain an MTKBarostat hook.")

    def copy(self):
        return self.__class__(self.attr)
