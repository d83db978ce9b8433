"""Minimal screen logger with the call surface the ``simulations/*.py`` scripts use from ``micmec.log``
(``log(...)``, ``log.set_level``, ``log.hline``, unit formatters); quiet by default."""
import sys

from . import units

__all__ = ["log", "timer"]


class _Section(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class ScreenLog(object):
    silent, warning, low, medium, high, debug = range(6)

    def __init__(self, name="MICMEC-B200", stream=None):
        self.name = name
        self.stream = stream or sys.stdout
        self._level = self.warning

    do_warning = property(lambda self: self._level >= self.warning)
    do_low = property(lambda self: self._level >= self.low)
    do_medium = property(lambda self: self._level >= self.medium)
    do_high = property(lambda self: self._level >= self.high)
    do_debug = property(lambda self: self._level >= self.debug)

    def set_level(self, level):
        self._level = int(level)

    def __call__(self, *words):
        self.stream.write(" ".join(str(w) for w in words).replace("&", " ") + "\n")

    def warn(self, *words):
        self("WARNING:", *words)

    def hline(self, char="~"):
        self(char * 80)

    def blank(self):
        self("")

    def section(self, name):
        return _Section()

    def print_footer(self):
        pass

    # unit formatters (atomic units in, the reference's default display units out)
    def length(self, value):
        return "%10.4f" % (value / units.angstrom)

    def energy(self, value):
        return "%10.1f" % (value / units.kjmol)

    def force(self, value):
        return "%10.1f" % (value / (units.kjmol / units.angstrom))

    def temperature(self, value):
        return "%10.1f" % value

    def angle(self, value):
        return "%10.4f" % (value * 57.29577951308232)

    def volume(self, value):
        return "%10.4f" % (value / units.angstrom ** 3)


class _Timer(object):
    def section(self, name):
        return _Section()


log = ScreenLog()
timer = _Timer()
