"""Barostats and the thermostat-barostat combination (micmec/sampling/npt.py:52-757).

In the device-resident integrator the whole of ``MTKBarostat.baro`` (npt.py:653-736) - barostat velocity update,
3x3 eigen-decompositions, position / cell / velocity rotations and the extra force evaluation - runs inside
libmicmec_b200.so; ``MTKBarostat`` and ``TBCombination(NHCThermostat, MTKBarostat)`` then only carry the parameters
and mirror the state (``vel_press``, ``mass_press``, ``econs_correction``).

Every other combination (``BerendsenBarostat``, ``LangevinBarostat`` - what ``simulations/md.py -press`` uses -,
an MTK barostat next to a Berendsen / Langevin thermostat, or with its own ``baro_thermo`` chain) runs in the
host-driven mode of ``VerletIntegrator``: the hook algebra below acts on the host arrays exactly like the reference
(same NumPy calls on the legacy global RNG, hence seeded parity) and each of its force evaluations is one
``mmf.compute`` on the GPU.
"""
import numpy as np

from ..units import bar, boltzmann, femtosecond
from .iterative import StateItem
from .nvt import BerendsenThermostat, LangevinThermostat, NHCThermostat
from .utils import clean_momenta, domain_symmetrize, get_ndof_baro, get_ndof_internal_md, get_random_vel_press
from .verlet import VerletHook

__all__ = ["TBCombination", "BerendsenBarostat", "LangevinBarostat", "MTKBarostat", "MTKAttributeStateItem"]


def _force(iterative):
    """One force evaluation with virial at the integrator's current geometry (npt.py:254-256 and alike)."""
    iterative.gpos[:] = 0.0
    iterative.vtens[:] = 0.0
    iterative.epot = iterative.mmf.compute(iterative.gpos, iterative.vtens)


def _move_geometry(iterative, pos_new, rvecs_new):
    iterative.mmf.update_pos(pos_new)
    iterative.pos[:] = pos_new
    iterative.mmf.update_rvecs(rvecs_new)
    iterative.rvecs[:] = rvecs_new


def _sym_exp(mat, dt, sign):
    """exp(sign * mat * dt / 2) of a symmetric 3x3 through its eigen-decomposition, with the operation order of
    npt.py:433-436, 470-472 (also 683-686, 716-718)."""
    evals, evecs = np.linalg.eigh(mat)
    return np.dot(np.dot(evecs, np.diagflat(np.exp(sign * evals * dt / 2))), evecs.T)


def _kinetic_stress(iterative):
    """Volume times the symmetrised instantaneous pressure tensor (npt.py:397-402, 659-664)."""
    pv = np.dot(iterative.vel.T * iterative.masses, iterative.vel) - iterative.vtens
    return 0.5 * (pv.T + pv)


def _rotate_and_evaluate(hook, iterative):
    """Shared middle part of ``LangevinBarostat.baro`` / ``MTKBarostat.baro``: move positions and cell with
    exp(v_g dt/2), evaluate the forces there, transform the velocities (npt.py:432-484, 683-733)."""
    dt = hook.timestep_press
    if hook.anisotropic:
        rot = _sym_exp(hook.vel_press, dt, 1)
        pos_new, rvecs_new = np.dot(iterative.pos, rot), np.dot(iterative.rvecs, rot)
    else:
        scale = np.exp(hook.vel_press * dt / 2)
        pos_new, rvecs_new = scale * iterative.pos, scale * iterative.rvecs
    _move_geometry(iterative, pos_new, rvecs_new)
    _force(iterative)
    if hook.anisotropic:
        vp = hook.vel_press
        if not hook.vol_constraint:
            vp = vp + (np.trace(vp) / iterative.ndof) * np.eye(3)
        vel_new = np.dot(iterative.vel, _sym_exp(vp, dt, -1))
    else:
        vel_new = np.exp(-((1.0 + 3.0 / iterative.ndof) * hook.vel_press) * dt / 2) * iterative.vel
    iterative.vel[:] = vel_new
    iterative.ekin = iterative._compute_ekin()


class BerendsenBarostat(VerletHook):
    """Weak-coupling barostat: one symmetric scaling of positions and cell per step (npt.py:180-292)."""

    name = "Berendsen"
    kind = "deterministic"
    method = "barostat"

    def __init__(self, mmf, temp, press, start=0, step=1, timecon=1000 * femtosecond, beta=4.57e-5 / bar,
                 anisotropic=True, vol_constraint=False, restart=False):
        self.temp = temp
        self.press = press
        self.timecon_press = timecon
        self.beta = beta
        self.mass_press = 3.0 * timecon / beta
        self.anisotropic = anisotropic
        self.vol_constraint = vol_constraint
        self.dim = mmf.system.domain.nvec
        self.baro_ndof = get_ndof_baro(self.dim, anisotropic, vol_constraint)
        if anisotropic:
            domain_symmetrize(mmf)
        self.domain = np.array(mmf.system.domain.rvecs)
        self.restart = restart
        VerletHook.__init__(self, start, step)

    def init(self, iterative):
        self.timestep_press = iterative.timestep
        if not self.restart:
            clean_momenta(iterative.pos, iterative.vel, iterative.masses, iterative.mmf.system.domain)
        _force(iterative)
        if iterative.ndof is None:
            iterative.ndof = get_ndof_internal_md(iterative.pos.shape[0], iterative.mmf.system.domain.nvec)
        self.mass_press *= np.sqrt(iterative.ndof)

    def pre(self, iterative, chainvel0=None):
        pass

    def post(self, iterative, chainvel0=None):
        before = iterative.epot
        ptens = (np.dot(iterative.vel.T * iterative.masses, iterative.vel) - iterative.vtens) / iterative.mmf.system.domain.volume
        dmu = self.timestep_press / self.mass_press * (self.press * np.eye(3) - ptens)
        if self.vol_constraint:
            dmu -= np.trace(dmu) / self.dim * np.eye(self.dim)
        mu = np.eye(3) - dmu
        mu = 0.5 * (mu + mu.T)
        if not self.anisotropic:
            mu = ((np.trace(mu) / 3.0) ** (1.0 / 3.0)) * np.eye(3)
        _move_geometry(iterative, np.dot(iterative.pos, mu), np.dot(iterative.rvecs, mu))
        _force(iterative)
        self.econs_correction += before - iterative.epot


class LangevinBarostat(VerletHook):
    """Stochastic piston (Feller et al.) acting on a 3x3 barostat velocity (npt.py:295-510)."""

    name = "Langevin"
    kind = "stochastic"
    method = "barostat"

    def __init__(self, mmf, temp, press, start=0, step=1, timecon=1000 * femtosecond, anisotropic=True,
                 vol_constraint=False):
        self.temp = temp
        self.press = press
        self.timecon = timecon
        self.anisotropic = anisotropic
        self.vol_constraint = vol_constraint
        self.dim = mmf.system.domain.nvec
        self.baro_ndof = get_ndof_baro(self.dim, anisotropic, vol_constraint)
        if anisotropic:
            domain_symmetrize(mmf)
        self.domain = np.array(mmf.system.domain.rvecs)
        VerletHook.__init__(self, start, step)

    def init(self, iterative):
        self.timestep_press = iterative.timestep
        clean_momenta(iterative.pos, iterative.vel, iterative.masses, iterative.mmf.system.domain)
        if iterative.ndof is None:
            iterative.ndof = iterative.pos.size
        self.mass_press = (iterative.ndof + 3) / 3 * boltzmann * self.temp * (self.timecon / (2 * np.pi)) ** 2
        self.vel_press = get_random_vel_press(self.mass_press, self.temp)
        if self.vol_constraint:
            self.vel_press -= np.trace(self.vel_press) / 3 * np.eye(3)
        if not self.anisotropic:
            self.vel_press = self.vel_press[0][0]
        _force(iterative)

    def _tracked_baro(self, iterative, chainvel0):
        epot0, ekin0 = iterative.epot, iterative.ekin
        self.baro(iterative, chainvel0)
        self.econs_correction += epot0 - iterative.epot + ekin0 - iterative.ekin

    def pre(self, iterative, chainvel0=None):
        self._tracked_baro(iterative, chainvel0)

    def post(self, iterative, chainvel0=None):
        self._tracked_baro(iterative, chainvel0)

    def _update_vel_press(self, iterative, chainvel0):
        dt = self.timestep_press
        friction = np.exp(-dt / (8 * self.timecon))
        self.vel_press *= friction
        if chainvel0 is not None:
            self.vel_press *= np.exp(-dt * chainvel0 / 8)
        drive = (_kinetic_stress(iterative) + (2.0 * iterative.ekin / iterative.ndof
                                               - self.press * iterative.mmf.system.domain.volume) * np.eye(3)) / self.mass_press
        noise = self.getR()
        if self.vol_constraint:
            drive -= np.trace(drive) / self.dim * np.eye(self.dim)
            noise -= np.trace(noise) / self.dim * np.eye(self.dim)
        if not self.anisotropic:
            drive = np.trace(drive)
            noise = noise[0][0]
        self.vel_press += (drive - noise / self.mass_press) * dt / 4
        self.vel_press *= friction
        if chainvel0 is not None:
            self.vel_press *= np.exp(-dt * chainvel0 / 8)

    def baro(self, iterative, chainvel0):
        self._update_vel_press(iterative, chainvel0)
        _rotate_and_evaluate(self, iterative)
        self._update_vel_press(iterative, chainvel0)

    def getR(self):
        """Symmetric 3x3 Gaussian noise: the lower triangle of one (3, 3) draw mirrored upwards (npt.py:492-510)."""
        sigma = np.sqrt(2 * self.mass_press * boltzmann * self.temp / (self.timestep_press * self.timecon))
        rand = np.random.normal(0, 1, (3, 3)) * sigma
        return np.tril(rand) + np.tril(rand, -1).T


class MTKBarostat(VerletHook):
    name = "MTTK"
    kind = "deterministic"
    method = "barostat"
    native = True

    def __init__(self, mmf, temp, press, start=0, step=1, timecon=1000 * femtosecond, anisotropic=True,
                 vol_constraint=False, baro_thermo=None, vel_press0=None, restart=False):
        self.temp = temp
        self.press = press
        self.timecon_press = timecon
        self.anisotropic = anisotropic
        self.vol_constraint = vol_constraint
        self.baro_thermo = baro_thermo  # a barostat with its own chain runs in host-driven mode (native = False)
        if baro_thermo is not None:
            self.native = False
        self.dim = mmf.system.domain.nvec
        self.restart = restart
        self.baro_ndof = get_ndof_baro(self.dim, self.anisotropic, self.vol_constraint)
        if self.anisotropic and not self.restart:
            domain_symmetrize(mmf)  # npt.py:572-575
        self.domain = np.array(mmf.system.domain.rvecs)
        self.vel_press = vel_press0
        VerletHook.__init__(self, start, step)

    def init(self, iterative):
        """npt.py:579-614; in device mode the force re-evaluation is left to the device integrator."""
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
        if self.baro_thermo is not None:
            self.baro_thermo.chain.timestep = iterative.timestep
            self.baro_thermo.chain.set_ndof(self.baro_ndof)
        if self.vol_constraint:
            self.vel_press = self.vel_press - np.trace(self.vel_press) / 3 * np.eye(3)
        if not getattr(iterative, "device_mode", False):
            _force(iterative)

    # ---- host-driven mode (npt.py:616-736); the device integrator runs the same algebra in its scalar kernel ----
    def _thermostat_barostat(self):
        vp = np.array(self.vel_press, dtype=float)  # the chain scales a copy: only its own state advances
        self.baro_thermo.chain(self._compute_ekin_baro(), vp, 0)

    def pre(self, iterative, chainvel0=None):
        if self.baro_thermo is not None:
            chainvel0 = self.baro_thermo.chain.vel[0]
        self.baro(iterative, chainvel0)
        if self.baro_thermo is not None:
            self._thermostat_barostat()

    def post(self, iterative, chainvel0=None):
        if self.baro_thermo is not None:
            self._thermostat_barostat()
            chainvel0 = self.baro_thermo.chain.vel[0]
        self.baro(iterative, chainvel0)
        self.econs_correction = self._compute_ekin_baro()
        if not self.vol_constraint:
            self.econs_correction += self.press * iterative.mmf.system.domain.volume
        if self.baro_thermo is not None:
            self.econs_correction += self.baro_thermo.chain.get_econs_correction()

    def _update_vel_press(self, iterative, chainvel0):
        dt = self.timestep_press
        if chainvel0 is not None:
            self.vel_press *= np.exp(-dt * chainvel0 / 8)
        drive = (_kinetic_stress(iterative) + (2.0 * iterative.ekin / iterative.ndof
                                               - self.press * iterative.mmf.system.domain.volume) * np.eye(3)) / self.mass_press
        if not self.anisotropic:
            drive = np.trace(drive)
        if self.vol_constraint:
            drive -= np.trace(drive) / self.dim * np.eye(self.dim)
        self.vel_press += drive * dt / 4
        if chainvel0 is not None:
            self.vel_press *= np.exp(-dt * chainvel0 / 8)

    def baro(self, iterative, chainvel0):
        self._update_vel_press(iterative, chainvel0)
        _rotate_and_evaluate(self, iterative)
        self._update_vel_press(iterative, chainvel0)

    def add_press_cont(self):
        if self.baro_thermo is not None:
            return 0
        return 2 * self._compute_ekin_baro() - self.baro_ndof * self.temp * boltzmann

    def _compute_ekin_baro(self):
        if self.anisotropic:
            return 0.5 * self.mass_press * np.trace(np.dot(self.vel_press.T, self.vel_press))
        return 0.5 * self.mass_press * self.vel_press ** 2


class TBCombination(VerletHook):
    """Calls a thermostat and a barostat in the order, and with the couplings, of npt.py:52-177.  The pair
    (NHCThermostat, MTKBarostat) called every step is propagated on the device; any other supported pair runs in the
    host-driven mode."""

    name = "TBCombination"

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
        VerletHook.__init__(self, start, min(self.step_thermo, self.step_baro))

    @property
    def native(self):
        return (isinstance(self.thermostat, NHCThermostat) and isinstance(self.barostat, MTKBarostat)
                and self.barostat.baro_thermo is None and self.step_thermo == 1 and self.step_baro == 1 and self.start == 0)

    def init(self, iterative):
        ndof_given = iterative.ndof is not None
        self.thermostat.init(iterative)
        self.barostat.init(iterative)
        # the centre of mass fluctuates under Langevin dynamics: all 3N degrees of freedom count (npt.py:89-95)
        if not ndof_given and (isinstance(self.thermostat, LangevinThermostat) or isinstance(self.barostat, LangevinBarostat)):
            iterative.ndof = iterative.pos.size
        self.chainvel0 = None
        self.G1_add = None

    def _call_baro(self, iterative, kind):
        if self.expectscall(iterative, "baro"):
            if isinstance(self.thermostat, NHCThermostat):
                self.chainvel0 = self.thermostat.chain.vel[0]  # v_xi,1 damps the barostat velocity
            getattr(self.barostat, kind)(iterative, self.chainvel0)

    def _call_thermo(self, iterative, kind):
        if self.expectscall(iterative, "thermo"):
            if isinstance(self.barostat, MTKBarostat):
                self.G1_add = self.barostat.add_press_cont()  # the barostat's kinetic energy drives the first bead
            getattr(self.thermostat, kind)(iterative, self.G1_add)

    def pre(self, iterative):
        self._call_baro(iterative, "pre")
        self._call_thermo(iterative, "pre")

    def post(self, iterative):
        self._call_thermo(iterative, "post")
        self._call_baro(iterative, "post")
        self.econs_correction = self.thermostat.econs_correction + self.barostat.econs_correction
        if isinstance(self.thermostat, NHCThermostat) and not (
                isinstance(self.barostat, MTKBarostat) and self.barostat.baro_thermo is not None):
            # the particle chain also thermostats the barostat (npt.py:134-148)
            self.econs_correction += self.barostat.baro_ndof * boltzmann * self.thermostat.temp * self.thermostat.chain.pos[0]

    def expectscall(self, iterative, kind):
        step = self.step_thermo if kind == "thermo" else self.step_baro
        return iterative.counter >= self.start and (iterative.counter - self.start) % step == 0

    def verify(self):
        return (isinstance(self.thermostat, (NHCThermostat, LangevinThermostat, BerendsenThermostat))
                and isinstance(self.barostat, (BerendsenBarostat, LangevinBarostat, MTKBarostat)))


class MTKAttributeStateItem(StateItem):
    def __init__(self, attr):
        StateItem.__init__(self, "baro_" + attr)
        self.attr = attr

    def get_value(self, iterative):
        """npt.py:765-782: the first MTK barostat (bare or inside the first ``TBCombination``); ``chain_pos`` /
        ``chain_vel`` come from the barostat's own thermostat chain (0 without one)."""
        found = None
        for hook in iterative.hooks:
            if isinstance(hook, MTKBarostat):
                found = hook
                break
            if isinstance(hook, TBCombination):
                found = hook.barostat if isinstance(hook.barostat, MTKBarostat) else None
                break
        if found is None:
            raise TypeError("Iterative does not contain an MTKBarostat hook.")
        if self.key.startswith("baro_chain_"):
            chain_attr = self.key.split("_")[2]
            return getattr(found.baro_thermo.chain, chain_attr) if found.baro_thermo is not None else 0
        return getattr(found, self.attr)

    def copy(self):
        return self.__class__(self.attr)
