"""Thermostats: the Nose-Hoover chain (micmec/sampling/nvt.py:361-532) of the device-resident integrator, and
host-driven Andersen / Berendsen / Langevin / CSVR / GLE hooks (nvt.py:48-358).

``NHChain`` holds the chain state on the host (positions, velocities, masses) exactly like the reference; the
chain *propagation* of ``NHChain.__call__`` (nvt.py:410-451) runs inside the scalar kernel of libmicmec_b200.so and
this object is refreshed from the device whenever the host needs it.  ``NHChain.__call__`` is kept as a plain
NumPy method too, because the host-driven compatibility mode of ``VerletIntegrator`` (arbitrary Python hooks)
calls thermostats the way the reference does.

The other thermostats act on the host arrays of ``VerletIntegrator`` in its host-driven mode (every force
evaluation still runs on the GPU through ``mmf.compute``).  They draw from the legacy global NumPy RNG with the
same calls in the same order as the reference, so a seeded run reproduces the reference's trajectory exactly
(tests/test_hooks_cpu.py, tests/test_hooks_gpu.py).
"""
import numpy as np

from ..units import boltzmann, femtosecond
from .iterative import StateItem
from .utils import clean_momenta, get_ndof_internal_md, get_random_vel, stabilized_cholesky_decomp
from .verlet import VerletHook

__all__ = [
    "AndersenThermostat", "BerendsenThermostat", "LangevinThermostat", "CSVRThermostat", "GLEThermostat",
    "NHChain", "NHCThermostat", "NHCAttributeStateItem",
]


class _HostThermostat(VerletHook):
    """Common part of the thermostats that act on the host arrays: momentum cleaning and the default number of
    degrees of freedom at ``init``; ``pre`` does the work unless a subclass also defines ``post``.  ``kind`` says whether
    random numbers are involved; ``TBCombination`` and ``VerletIntegrator._verify_hooks`` look at ``method``."""

    method = "thermostat"
    sets_ndof = False      # fills in iterative.ndof (3N minus the conserved momenta) when the user gave none
    always_cleans = True   # False: skip the momentum cleaning on a restart

    def __init__(self, temp, start=0, step=1, timecon=None):
        self.temp = temp
        if timecon is not None:
            self.timecon = timecon
        VerletHook.__init__(self, start, step)

    def init(self, iterative):
        if self.always_cleans or not getattr(self, "restart", False):
            clean_momenta(iterative.pos, iterative.vel, iterative.masses, iterative.mmf.system.domain)
        if self.sets_ndof and iterative.ndof is None:
            iterative.ndof = get_ndof_internal_md(iterative.pos.shape[0], iterative.mmf.system.domain.nvec)

    def post(self, iterative, G1_add=None):
        pass


class AndersenThermostat(_HostThermostat):
    """Velocities redrawn from the Maxwell-Boltzmann distribution every call (nvt.py:48-107)."""

    name = "Andersen"
    kind = "stochastic"

    def __init__(self, temp, start=0, step=1, select=None, annealing=1.0):
        _HostThermostat.__init__(self, temp, start, step)
        self.select, self.annealing = select, annealing

    def pre(self, iterative, G1_add=None):
        before = iterative._compute_ekin()
        if self.select is None:
            iterative.vel[:] = get_random_vel(self.temp, False, iterative.masses)
        else:
            iterative.vel[self.select] = get_random_vel(self.temp, False, iterative.masses, self.select)
        clean_momenta(iterative.pos, iterative.vel, iterative.masses, iterative.mmf.system.domain)
        self.econs_correction += before - iterative._compute_ekin()
        self.temp *= self.annealing


class BerendsenThermostat(_HostThermostat):
    """Weak-coupling velocity rescaling (nvt.py:110-162)."""

    name = "Berendsen"
    kind = "deterministic"
    sets_ndof = True
    always_cleans = False

    def __init__(self, temp, start=0, timecon=100 * femtosecond, restart=False):
        _HostThermostat.__init__(self, temp, start, 1, timecon)
        self.restart = restart

    def pre(self, iterative, G1_add=None):
        ekin = iterative.ekin
        temp_now = 2.0 * ekin / (boltzmann * iterative.ndof)
        scale = np.sqrt(1 + iterative.timestep / self.timecon * (self.temp / temp_now - 1))
        iterative.vel[:] = scale * iterative.vel
        iterative.ekin = iterative._compute_ekin()
        self.econs_correction += (1 - scale ** 2) * ekin


class LangevinThermostat(_HostThermostat):
    """Half-step Ornstein-Uhlenbeck velocity update before and after the Verlet step (nvt.py:165-218); this is the
    thermostat ``simulations/md.py`` uses.

    Two implementations behind the reference's constructor:

    * host-driven (the reference's arithmetic on NumPy arrays, noise from the legacy global NumPy generator in the
      reference's call order): a seeded run reproduces the reference's trajectory; every force evaluation crosses PCIe;
    * device-resident (``device=True``; the default from 4096 nodes on, when this is the only Verlet hook): the update
      runs in ``k_langevin`` inside ``mm_md_run`` with a counter-based generator (Philox4x32-10) seeded from the NumPy
      generator at initialisation.  Same distribution, different stream: parity with the reference is statistical.
    """

    name = "Langevin"
    kind = "stochastic"
    device_threshold = 4096  # nodes

    def __init__(self, temp, start=0, timecon=100 * femtosecond, device=None):
        _HostThermostat.__init__(self, temp, start, 1, timecon)
        self.device = device
        self.seed = None

    def wants_device(self, nnodes):
        return bool(self.device) if self.device is not None else nnodes >= self.device_threshold

    def thermo(self, iterative):
        damp = np.exp(-iterative.timestep / self.timecon / 2)
        kick = np.sqrt((1.0 - damp ** 2) * self.temp * boltzmann / iterative.masses).reshape(-1, 1)
        iterative.vel[:] = damp * iterative.vel + kick * np.random.normal(0, 1, iterative.vel.shape)
        iterative.ekin = iterative._compute_ekin()

    def _tracked(self, iterative, G1_add=None):
        before = iterative.ekin
        self.thermo(iterative)
        self.econs_correction += before - iterative.ekin

    pre = post = _tracked


class CSVRThermostat(_HostThermostat):
    """Canonical sampling through stochastic velocity rescaling (nvt.py:221-274).  Host-driven by default for small systems
    (NumPy stream, the reference's trajectory reproduced); with ``device=True`` (default from 4096 nodes on, when it is the
    only Verlet hook) the single velocity scale per step is drawn and applied on the device (Philox stream, chi-square by
    Marsaglia-Tsang instead of ``ndof - 1`` explicit normal deviates): same distribution, statistical parity."""

    name = "CSVR"
    kind = "stochastic"
    sets_ndof = True
    device_threshold = 4096

    def __init__(self, temp, start=0, timecon=100 * femtosecond, device=None):
        _HostThermostat.__init__(self, temp, start, 1, timecon)
        self.device = device
        self.seed = None

    def wants_device(self, nnodes):
        return bool(self.device) if self.device is not None else nnodes >= self.device_threshold

    def init(self, iterative):
        _HostThermostat.init(self, iterative)
        self.kin = 0.5 * iterative.ndof * boltzmann * self.temp

    def pre(self, iterative, G1_add=None):
        decay = np.exp(-iterative.timestep / self.timecon)
        first = np.random.normal(0, 1)
        rest = (np.random.normal(0, 1, iterative.ndof - 1) ** 2).sum()
        iterative.ekin = iterative._compute_ekin()
        fact = (1 - decay) * self.kin / iterative.ndof / iterative.ekin
        alpha = np.sign(first + np.sqrt(decay / fact)) * np.sqrt(
            decay + (rest + first ** 2) * fact + 2 * first * np.sqrt(decay * fact))
        iterative.vel[:] = alpha * iterative.vel
        iterative.ekin_new = alpha ** 2 * iterative.ekin
        self.econs_correction += (1 - alpha ** 2) * iterative.ekin
        iterative.ekin = iterative.ekin_new


class GLEThermostat(_HostThermostat):
    """Coloured-noise (generalised Langevin) thermostat with ``ns`` auxiliary momenta per coordinate (nvt.py:277-358)."""

    name = "GLE"
    kind = "stochastic"

    def __init__(self, temp, a_p, c_p=None, start=0):
        _HostThermostat.__init__(self, temp, start, 1)
        self.ns = int(a_p.shape[0] - 1)
        self.a_p = a_p
        self.c_p = boltzmann * temp * np.eye(self.ns + 1) if c_p is None else c_p

    def init(self, iterative):
        _HostThermostat.init(self, iterative)
        self.s = 0.5 * boltzmann * self.temp * np.random.normal(size=(self.ns, iterative.pos.size))
        evals, evecs = np.linalg.eig(-self.a_p * iterative.timestep / 2)
        self.t = np.dot(evecs * np.exp(evals), np.linalg.inv(evecs)).real
        self.S = stabilized_cholesky_decomp(self.c_p - self.t @ self.c_p @ self.t.T).real
        self.n_atoms = iterative.pos.shape[0]

    def thermo(self, iterative, G1_add=None):
        before = iterative.ekin
        root_m = np.sqrt(iterative.masses).reshape(-1, 1)
        old = np.vstack([(root_m * iterative.vel).reshape(-1), self.s])
        new = self.t @ old + self.S @ np.random.normal(size=(self.ns + 1, 3 * self.n_atoms))
        iterative.vel[:] = np.sqrt(1.0 / iterative.masses).reshape(-1, 1) * new[0].reshape(self.n_atoms, 3)
        self.s[:] = new[1:]
        iterative.ekin = iterative._compute_ekin()
        self.econs_correction += before - iterative.ekin

    pre = post = thermo


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
