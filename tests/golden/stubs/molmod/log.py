"""Silent ScreenLog / TimerGroup stand-ins (surface used by micmec/log.py and the hooks)."""
import contextlib


class TimerGroup(object):
    @contextlib.contextmanager
    def section(self, label):
        yield


class ScreenLog(object):
    silent, warning, low, medium, high, debug = 0, 1, 2, 3, 4, 5

    def __init__(self, name, version, head_banner, foot_banner, timer):
        self._level = self.silent

    do_warning = property(lambda self: self._level >= self.warning)
    do_low = property(lambda self: self._level >= self.low)
    do_medium = property(lambda self: self._level >= self.medium)
    do_high = property(lambda self: self._level >= self.high)
    do_debug = property(lambda self: self._level >= self.debug)

    def set_level(self, level):
        self._level = level

    def __call__(self, *words):
        pass

    warn = hline = blank = center = print_header = print_footer = enter = leave = __call__

    @contextlib.contextmanager
    def section(self, key):
        yield

    def _unit(self, value):
        return "%10.5f" % value

    length = energy = force = temperature = angle = volume = pressure = time = mass = _unit
