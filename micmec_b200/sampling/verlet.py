"""Velocity Verlet on the device: drop-in for ``micmec.sampling.verlet.VerletIntegrator``.

Constructor, attributes (``pos vel gpos vtens ekin epot etot econs cons_err temp press ptens rmsd_* time counter
ndof timestep masses rvecs``), the state-item list and the hook protocol are those of the reference
(micmec/sampling/verlet.py:49-272).  What differs is *where* a step runs:

* **device mode** (no Verlet hook, or the native ones: ``NHCThermostat``, ``MTKBarostat``, ``TBCombination`` of the
  two).  ``run(n)`` advances up to the next iteration at which a conventional hook (screen log, trajectory
  writer, user hook) expects a call with ONE C call (``mm_md_run``): kick, drift, force, kick, thermostat and
  barostat all execute as CUDA kernels on one stream, no host round trip.  Host mirrors of the state are refreshed
  only then - scalars always, the big arrays only if a hook other than ``VerletScreenLog`` fires.
* **host-driven mode** (any other ``VerletHook`` with ``pre``/``post`` written against NumPy arrays): the step is
  the reference's ``propagate`` on host arrays and every force evaluation goes to the GPU through
  ``mmf.compute``.  Correct but PCIe-bound; it exists so that arbitrary Python hooks keep working.
"""
import ctypes
import time

import numpy as np

from .. import _lib
from ..units import boltzmann, kelvin
from .iterative import (
    Iterative, Hook, AttributeStateItem, PosStateItem, TemperatureStateItem, VolumeStateItem, DomainStateItem,
)
from .utils import get_random_vel, clean_momenta

__all__ = ["VerletIntegrator", "VerletHook", "VerletScreenLog", "ConsErrTracker"]


class ConsErrTracker(object):
    """Running ratio of the fluctuations of the conserved quantity and the kinetic energy (verlet.py:275-307)."""

    def __init__(self):
        self.counter = 0
        self.ekin_m = self.ekin_s = self.econs_m = self.econs_s = 0.0

    def update(self, ekin, econs):
        if self.counter == 0:
            self.ekin_m, self.econs_m = ekin, econs
        else:
            d = ekin - self.ekin_m
            self.ekin_m += d / (self.counter + 1)
            self.ekin_s += d * (ekin - self.ekin_m)
            d = econs - self.econs_m
            self.econs_m += d / (self.counter + 1)
            self.econs_s += d * (econs - self.econs_m)
        self.counter += 1

    def get(self):
        return np.sqrt(self.econs_s / self.ekin_s) if self.counter > 1 else 0.0


class VerletHook(Hook):
    """Hook with ``init`` / ``pre`` / ``post`` entry points around the Verlet step (verlet.py:310-339).

    As in the reference, ``start`` and ``step`` are ignored: Verlet hooks fire every step (verlet.py:327).
    """

    native = False  # True for hooks libmicmec_b200.so propagates itself

    def __init__(self, start=0, step=1):
        self.econs_correction = 0.0
        Hook.__init__(self, start=0, step=1)

    def __call__(self, iterative):
        pass

    def init(self, iterative):
        raise NotImplementedError

    def pre(self, iterative):
        raise NotImplementedError

    def post(self, iterative):
        raise NotImplementedError


class VerletScreenLog(Hook):
    """Screen logger (verlet.py:342-374).  Prints when ``verbose`` is True, or - with ``verbose=None``, the default -
    when the package log level is at least medium, as the reference does (``micmec_b200.log.log.set_level``; the level
    is quiet by default and raised to medium by ``python -m micmec_b200.dropin``)."""

    def __init__(self, start=0, step=1, verbose=None):
        Hook.__init__(self, start, step)
        self.time0 = None
        self._verbose = verbose
        self.lines = 0

    @property
    def verbose(self):
        if self._verbose is None:
            from ..log import log

            return log.do_medium
        return self._verbose

    def __call__(self, iterative):
        if self.time0 is None:
            self.time0 = time.time()
            if self.verbose:
                print("counter  Cons.Err.       Temp     d-RMSD     g-RMSD   Walltime")
        self.lines += 1
        if self.verbose:
            print("%7i %10.5f %10.2f %10.3e %10.3e %10.1f" % (
                iterative.counter, iterative.cons_err, iterative.temp, iterative.rmsd_delta, iterative.rmsd_gpos,
                time.time() - self.time0))


class VerletIntegrator(Iterative):
    default_state = [
        AttributeStateItem("counter"), AttributeStateItem("time"), AttributeStateItem("epot"), PosStateItem(),
        AttributeStateItem("vel"), AttributeStateItem("rmsd_delta"), AttributeStateItem("rmsd_gpos"),
        AttributeStateItem("ekin"), TemperatureStateItem(), AttributeStateItem("etot"), AttributeStateItem("econs"),
        AttributeStateItem("cons_err"), AttributeStateItem("ptens"), AttributeStateItem("vtens"),
        AttributeStateItem("press"), VolumeStateItem(), DomainStateItem(),
    ]
    log_name = "VERLET"

    def __init__(self, mmf, timestep=None, state=None, hooks=None, vel0=None, temp0=300 * kelvin, scalevel0=True,
                 time0=None, ndof=None, counter0=None):
        self.ndof = ndof
        self.hooks = hooks
        self.pos = mmf.system.pos.copy()
        self.rvecs = np.array(mmf.system.domain.rvecs)
        self.masses = mmf.system.masses.copy()
        self.timestep = timestep
        self.time = 0.0 if time0 is None else time0
        self._verify_hooks()
        if vel0 is None:
            self.vel = get_random_vel(temp0, scalevel0, self.masses)
            clean_momenta(self.pos, self.vel, self.masses, mmf.system.domain)
        else:
            self.vel = vel0.copy()
        self.gpos = np.zeros(self.pos.shape, float)
        self.delta = np.zeros(self.pos.shape, float)
        self.vtens = np.zeros((3, 3), float)
        self.ptens = np.zeros((3, 3), float)
        self.press = 0.0
        self._cons_err_tracker = ConsErrTracker()
        self._md = ctypes.c_void_p()
        self._lib = None
        Iterative.__init__(self, mmf, state, self.hooks, 0 if counter0 is None else counter0)

    # ---- hook bookkeeping ----------------------------------------------------------------------------------
    def _add_default_hooks(self):
        if not any(isinstance(hook, VerletScreenLog) for hook in self.hooks):
            self.hooks.append(VerletScreenLog())

    def _verify_hooks(self):
        """Merge a separate thermostat and barostat into one TBCombination (verlet.py:216-248)."""
        if self.hooks is None or not hasattr(self.hooks, "__len__"):
            return
        thermo = [h for h in self.hooks if getattr(h, "method", None) == "thermostat"]
        baro = [h for h in self.hooks if getattr(h, "method", None) == "barostat"]
        if thermo and baro:
            from .npt import TBCombination

            self.hooks.remove(thermo[-1])
            self.hooks.remove(baro[-1])
            self.hooks.append(TBCombination(thermo[-1], baro[-1]))

    def _verlet_hooks(self):
        return [hook for hook in self.hooks if isinstance(hook, VerletHook)]

    def call_verlet_hooks(self, kind):
        for hook in self._verlet_hooks():
            if hook.expects_call(self.counter):
                getattr(hook, kind)(self)

    def _native_setup(self):
        """Return (part, thermostat, barostat) when the step can stay on the device, else None."""
        from ..pes.mmff import ForcePartMechanical
        from .nvt import NHCThermostat, LangevinThermostat, BerendsenThermostat, CSVRThermostat
        from .npt import MTKBarostat, TBCombination

        parts = getattr(self.mmf, "parts", None)
        if not parts or len(parts) != 1 or not isinstance(parts[0], ForcePartMechanical):
            return None
        thermo = baro = None
        vhooks = self._verlet_hooks()
        if len(vhooks) > 1:
            return None
        for hook in vhooks:
            if isinstance(hook, (LangevinThermostat, CSVRThermostat)) and hook.start == 0 and getattr(parts[0], "slab", None) is None \
                    and hook.wants_device(self.pos.shape[0]):
                thermo = hook  # device-resident Langevin thermostat (k_langevin)
                continue
            if isinstance(hook, BerendsenThermostat) and hook.start == 0 and getattr(parts[0], "slab", None) is None:
                thermo = hook  # weak coupling: one deterministic velocity scale per step, computed by the scalar kernel
                continue
            if not hook.native:  # Berendsen / CSVR / ... hooks, an MTK barostat with its own chain, user hooks
                return None
            if isinstance(hook, TBCombination):
                thermo, baro = hook.thermostat, hook.barostat
            elif isinstance(hook, NHCThermostat):
                thermo = hook
            elif isinstance(hook, MTKBarostat):
                baro = hook
            else:
                return None
        if thermo is not None and hasattr(thermo, "chain") and thermo.chain.length > _lib.MM_MAX_CHAIN:
            return None
        return parts[0], thermo, baro

    # ---- initialisation (verlet.py:119-137) ---------------------------------------------------------------
    def initialize(self):
        self.delta[:] = 0.0
        setup = self._native_setup()
        self.device_mode = setup is not None
        if not self.device_mode:
            return self._initialize_host()
        part, thermo, baro = setup
        self._part, self._thermo, self._baro = part, thermo, baro
        self.call_verlet_hooks("init")  # host side of the hooks' init: RNG draws, momentum cleaning, ndof
        if self.ndof is None:
            self.ndof = np.size(self.pos)
        self.posold = self.pos.copy()
        self._lib = _lib.load()
        desc = _lib.MDDesc()
        desc.timestep = self.timestep
        desc.ndof = float(self.ndof)
        desc.time0, desc.counter0 = float(self.time), int(self.counter)
        self._langevin = self._berendsen = None
        if thermo is not None and thermo.name in ("Berendsen", "CSVR"):  # one velocity scale per step from the scalar kernel
            self._berendsen, thermo = thermo, None
            self._thermo = None
            desc.has_thermo, desc.thermo_kind, desc.chain_length = 1, (1 if self._berendsen.name == "Berendsen" else 2), 0
            desc.thermo_temp, desc.thermo_timecon = self._berendsen.temp, self._berendsen.timecon
            if self._berendsen.name == "CSVR":
                if self._berendsen.seed is None:
                    self._berendsen.seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + 12345
                desc.langevin_seed = self._berendsen.seed & (2 ** 64 - 1)
        if thermo is not None and not hasattr(thermo, "chain"):  # LangevinThermostat in device mode
            self._langevin, thermo = thermo, None
            self._thermo = None
            if self._langevin.seed is None:  # from the legacy global generator: np.random.seed(k) makes the run reproducible
                self._langevin.seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + 12345
            desc.has_langevin = 1
            desc.langevin_temp, desc.langevin_timecon = self._langevin.temp, self._langevin.timecon
            desc.langevin_seed = self._langevin.seed & (2 ** 64 - 1)
        if thermo is not None:
            desc.has_thermo, desc.chain_length = 1, thermo.chain.length
            desc.thermo_temp, desc.thermo_timecon = thermo.chain.temp, thermo.chain.timecon
        if baro is not None:
            desc.has_baro, desc.anisotropic, desc.vol_constraint = 1, int(baro.anisotropic), int(baro.vol_constraint)
            desc.baro_temp, desc.baro_press, desc.baro_timecon = baro.temp, baro.press, baro.timecon_press
        _lib.check(self._lib.mm_md_create(part.handle, ctypes.byref(desc), ctypes.byref(self._md)))
        rv = np.zeros((3, 3))
        rv[: self.rvecs.shape[0]] = self.rvecs
        cpos = np.ascontiguousarray(thermo.chain.pos, dtype=float) if thermo is not None else None
        cvel = np.ascontiguousarray(thermo.chain.vel, dtype=float) if thermo is not None else None
        vp = None
        if baro is not None:
            vp = np.zeros((3, 3))
            if baro.anisotropic:
                vp[:] = baro.vel_press
            else:
                vp[0, 0] = baro.vel_press
        pos = np.ascontiguousarray(self.pos, dtype=float)
        vel = np.ascontiguousarray(self.vel, dtype=float)
        masses = np.ascontiguousarray(self.masses, dtype=float)
        _lib.check(self._lib.mm_md_init(self._md, _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(masses), _lib.MM_HOST,
                                        _lib.ptr(rv), _lib.ptr(cpos), _lib.ptr(cvel), _lib.ptr(vp)))
        self._sync(arrays=True)
        Iterative.initialize(self)

    def __del__(self):
        md = getattr(self, "_md", None)
        if md is not None and md.value and self._lib is not None:
            self._lib.mm_md_destroy(md)
            self._md = ctypes.c_void_p()

    # ---- device mode ------------------------------------------------------------------------------------------
    def _sync(self, arrays):
        """Refresh the host mirrors from the device (scalars always; pos/vel/gpos when ``arrays``)."""
        out = np.zeros(_lib.S_COUNT)
        _lib.check(self._lib.mm_md_scalars(self._md, _lib.ptr(out)))
        self.epot, self.ekin, self.temp = out[_lib.S_EPOT], out[_lib.S_EKIN], out[_lib.S_TEMP]
        self.etot, self.econs, self.cons_err = out[_lib.S_ETOT], out[_lib.S_ECONS], out[_lib.S_CONS_ERR]
        self.press, self.rmsd_gpos, self.rmsd_delta = out[_lib.S_PRESS], out[_lib.S_RMSD_GPOS], out[_lib.S_RMSD_DELTA]
        self.time = out[_lib.S_TIME]
        self.vtens = out[_lib.S_VTENS:_lib.S_VTENS + 9].reshape(3, 3).copy()
        self.ptens = out[_lib.S_PTENS:_lib.S_PTENS + 9].reshape(3, 3).copy()
        self.nforce = int(out[_lib.S_NFORCE])
        # the tracker and the per-part energies live on the device in this mode: mirror them for the state items
        tracker = self._cons_err_tracker
        tracker.counter = int(out[_lib.S_CE_N])
        tracker.ekin_m, tracker.ekin_s, tracker.econs_m, tracker.econs_s = out[_lib.S_CE_N + 1:_lib.S_CE_N + 5]
        econs_corr = out[_lib.S_ECONS_CORR]
        rv, vp = np.zeros((3, 3)), np.zeros((3, 3))
        thermo, baro = self._thermo, self._baro
        cpos = np.zeros(thermo.chain.length) if thermo is not None else None
        cvel = np.zeros(thermo.chain.length) if thermo is not None else None
        p = v = g = None
        if arrays:
            p, v, g = self.pos, self.vel, self.gpos
        _lib.check(self._lib.mm_md_get_state(self._md, _lib.ptr(p), _lib.ptr(v), _lib.ptr(g), _lib.MM_HOST, _lib.ptr(rv),
                                             _lib.ptr(cpos), _lib.ptr(cvel), _lib.ptr(vp)))
        if arrays:
            self.mmf.update_pos(self.pos)
        part_energy = self.epot  # device mode runs a single force part (_native_setup)
        if baro is not None:
            self.rvecs = rv[: self.rvecs.shape[0]].copy()
            self.mmf.update_rvecs(np.ascontiguousarray(self.rvecs))
            baro.vel_press = vp.copy() if baro.anisotropic else vp[0, 0]
            baro.econs_correction = baro._compute_ekin_baro()
            if not baro.vol_constraint:
                baro.econs_correction += baro.press * self.mmf.system.domain.volume
        if thermo is not None:
            thermo.chain.pos, thermo.chain.vel = cpos, cvel
            thermo.econs_correction = thermo.chain.get_econs_correction()
        for hook in self._verlet_hooks():
            if hook.name == "TBCombination":
                hook.econs_correction = econs_corr
        for hook in (getattr(self, "_langevin", None), getattr(self, "_berendsen", None)):
            if hook is not None:
                hook.econs_correction = econs_corr
        self._part.energy = self.mmf.energy = part_energy  # after update_pos / update_rvecs cleared the caches
        self._arrays_fresh = arrays

    def _steps_to_next_call(self, limit):
        """Steps until a conventional hook next expects a call, at most ``limit``; also whether arrays are needed."""
        conventional = [hook for hook in self.hooks if not isinstance(hook, VerletHook)]
        for k in range(1, limit + 1):
            firing = [hook for hook in conventional if hook.expects_call(self.counter + k)]
            if firing:
                return k, any(not isinstance(hook, VerletScreenLog) and getattr(hook, "wants_arrays", True) for hook in firing)
        return limit, False

    def run(self, nsteps=None):
        if not self.device_mode:
            return Iterative.run(self, nsteps)
        if nsteps is None:
            raise ValueError("The device-resident integrator needs a finite number of steps.")
        done = 0
        while done < nsteps:
            k, need_arrays = self._steps_to_next_call(nsteps - done)
            _lib.check(self._lib.mm_md_run(self._md, k))
            done += k
            self.counter += k
            self._sync(arrays=need_arrays or done == nsteps)
            self.call_hooks()
        self.finalize()

    def propagate(self):
        if not self.device_mode:
            return self._propagate_host()
        _lib.check(self._lib.mm_md_run(self._md, 1))
        self.counter += 1
        self._sync(arrays=True)
        self.call_hooks()

    def call_hooks(self):
        if not getattr(self, "device_mode", False):
            return Iterative.call_hooks(self)
        state_updated = False
        for hook in self.hooks:
            if isinstance(hook, VerletHook) or not hook.expects_call(self.counter):
                continue
            if not state_updated and getattr(hook, "wants_arrays", True):  # hooks that read iterative.state
                for item in self.state_list:
                    item.update(self)
                state_updated = True
            hook(self)

    # ---- host-driven compatibility mode: the reference's propagate on NumPy arrays ---------------------------
    def _initialize_host(self):
        self.gpos[:] = 0.0
        self.mmf.update_pos(self.pos)
        self.epot = self.mmf.compute(self.gpos)
        self.acc = -self.gpos / self.masses.reshape(-1, 1)
        self.posold = self.pos.copy()
        self.call_verlet_hooks("init")
        if self.ndof is None:
            self.ndof = np.size(self.pos)
        self.compute_properties()
        Iterative.initialize(self)

    def _propagate_host(self):
        self.call_verlet_hooks("pre")
        self.acc = -self.gpos / self.masses.reshape(-1, 1)
        self.vel += 0.5 * self.acc * self.timestep
        self.pos += self.timestep * self.vel
        self.mmf.update_pos(self.pos)
        self.gpos[:] = 0.0
        self.vtens[:] = 0.0
        self.epot = self.mmf.compute(self.gpos, self.vtens)
        self.acc = -self.gpos / self.masses.reshape(-1, 1)
        self.vel += 0.5 * self.acc * self.timestep
        self.ekin = self._compute_ekin()
        self.call_verlet_hooks("post")
        self.delta[:] = self.pos - self.posold
        self.posold[:] = self.pos
        self.time += self.timestep
        self.compute_properties()
        Iterative.propagate(self)

    def _compute_ekin(self):
        return np.sum(0.5 * (self.vel ** 2.0 * self.masses.reshape(-1, 1)))

    def compute_properties(self):
        """verlet.py:171-190 (host-driven mode; in device mode the scalar kernel does this)."""
        self.rmsd_gpos = np.sqrt(np.mean(self.gpos ** 2))
        self.rmsd_delta = np.sqrt(np.mean(self.delta ** 2))
        self.ekin = self._compute_ekin()
        self.temp = (self.ekin / self.ndof) * (2.0 / boltzmann)
        self.etot = self.ekin + self.epot
        self.econs = self.etot + sum(hook.econs_correction for hook in self._verlet_hooks())
        self._cons_err_tracker.update(self.ekin, self.econs)
        self.cons_err = self._cons_err_tracker.get()
        if self.mmf.system.domain.nvec > 0:
            self.ptens = (np.dot(self.vel.T * self.masses, self.vel) - self.vtens) / self.mmf.system.domain.volume
            self.press = np.trace(self.ptens) / 3.0
