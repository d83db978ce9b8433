"""Scaffolding of the iterative algorithms (integrator, optimisers): hooks and state items.

The PROTOCOL is the one of micmec/sampling/iterative.py:41-222, so that hooks and state items written for the reference
keep working: a hook is any object with ``expects_call(counter)`` and ``__call__(iterative)``; a state item has ``key``,
``update(iterative)``, ``value``, ``shape``, ``dtype``, ``iter_attrs(iterative)`` and ``copy()``; an ``Iterative`` exposes
``state_list`` / ``state``, ``hooks``, ``counter``, ``initialize``, ``call_hooks``, ``run``, ``propagate``, ``finalize``.
The implementation is this package's own: state items are declared by a getter function instead of one hand-written class
each, and the state snapshot is refreshed lazily - only at iterations where some hook actually fires, which matters for
the device-resident integrator (a refresh of ``pos`` / ``vel`` is a device-to-host copy).
"""
import numpy as np

__all__ = [
    "Iterative", "StateItem", "AttributeStateItem", "PosStateItem", "EPotContribStateItem", "ConsErrStateItem",
    "TemperatureStateItem", "VolumeStateItem", "DomainStateItem", "Hook",
]


class Hook(object):
    """Called at iterations ``start, start + step, start + 2 step, ...``."""

    name = kind = method = None

    def __init__(self, start=0, step=1):
        self.start, self.step = start, step

    def expects_call(self, counter):
        since = counter - self.start
        return since >= 0 and since % self.step == 0

    def __call__(self, iterative):
        raise NotImplementedError


class StateItem(object):
    """One named quantity of the running algorithm.  ``update`` stores the current ``value``; ``shape`` and ``dtype``
    are fixed by the first update (trajectory writers size their datasets from them).  Subclasses either override
    ``get_value`` (the reference's way) or pass ``getter`` / ``attrs`` callables."""

    def __init__(self, key, getter=None, attrs=None):
        self.key = key
        self.shape = self.dtype = None
        self._getter, self._attrs = getter, attrs

    def get_value(self, iterative):
        if self._getter is None:
            raise NotImplementedError
        return self._getter(iterative)

    def update(self, iterative):
        value = self.value = self.get_value(iterative)
        if self.shape is not None:
            return
        if isinstance(value, np.ndarray):
            self.shape, self.dtype = value.shape, value.dtype
        else:
            self.shape, self.dtype = (), type(value)

    def iter_attrs(self, iterative):
        return [] if self._attrs is None else list(self._attrs(iterative))

    def copy(self):
        return type(self)()


class _KeyedItem(StateItem):
    """State item whose constructor takes the key (``AttributeStateItem("epot")``)."""

    def copy(self):
        return type(self)(self.key)


class AttributeStateItem(_KeyedItem):
    """``getattr(iterative, key)`` (``None`` while the attribute does not exist yet)."""

    def __init__(self, key):
        StateItem.__init__(self, key, getter=lambda it: getattr(it, key, None))


class ConsErrStateItem(_KeyedItem):
    """An attribute of the integrator's conserved-quantity tracker."""

    def __init__(self, key):
        StateItem.__init__(self, key, getter=lambda it: getattr(it._cons_err_tracker, key, None))


def _fixed_item(name, key, getter, attrs=None, doc=""):
    """Class of a state item with a fixed key: ``PosStateItem()`` and friends."""

    def __init__(self):
        StateItem.__init__(self, key, getter=getter, attrs=attrs)

    return type(name, (StateItem,), {"__init__": __init__, "__doc__": doc})


PosStateItem = _fixed_item("PosStateItem", "pos", lambda it: it.mmf.system.pos, doc="Node positions of the system.")
VolumeStateItem = _fixed_item("VolumeStateItem", "volume", lambda it: it.mmf.system.domain.volume, doc="Domain volume.")
DomainStateItem = _fixed_item("DomainStateItem", "domain", lambda it: it.mmf.system.domain.rvecs, doc="Domain vectors.")
TemperatureStateItem = _fixed_item(
    "TemperatureStateItem", "temp", lambda it: getattr(it, "temp", None), attrs=lambda it: [("ndof", it.ndof)],
    doc="Instantaneous temperature; the number of degrees of freedom travels as an attribute.")
EPotContribStateItem = _fixed_item(
    "EPotContribStateItem", "epot_contribs", lambda it: np.array([part.energy for part in it.mmf.parts]),
    attrs=lambda it: [("epot_contrib_names", np.array([part.name for part in it.mmf.parts], dtype="S"))],
    doc="Energy of every force part; the part names travel as an attribute.")


class Iterative(object):
    """Counter, hooks and state of an iterative algorithm; subclasses implement ``propagate`` (returning True to stop)."""

    default_state = []
    log_name = "ITER"

    def __init__(self, mmf, state=None, hooks=None, counter0=0):
        self.mmf = mmf
        self.state_list = [item.copy() for item in self.default_state] + list(state or [])
        self.state = {item.key: item for item in self.state_list}
        if hooks is None:
            hooks = []
        elif not hasattr(hooks, "__len__"):
            hooks = [hooks]
        self.hooks = hooks
        self._add_default_hooks()
        self.counter0 = self.counter = counter0
        self.initialize()

    def _add_default_hooks(self):
        pass

    def initialize(self):
        self.call_hooks()

    def _refresh_state(self):
        for item in self.state_list:
            item.update(self)

    def call_hooks(self):
        firing = [hook for hook in self.hooks if hook.expects_call(self.counter)]
        if firing:
            self._refresh_state()  # once per iteration, and only when somebody looks
        for hook in firing:
            hook(self)

    def run(self, nsteps=None):
        done = 0
        while nsteps is None or done < nsteps:
            done += 1
            if self.propagate():
                break
        self.finalize()

    def propagate(self):
        self.counter += 1
        self.call_hooks()

    def finalize(self):
        pass
