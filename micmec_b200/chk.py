"""Reader/writer for the text checkpoint (``.chk``) files MicMec systems are stored in.

The reference delegates this to ``molmod.io.chk.load_chk/dump_chk`` (micmec/system.py:30,238-252,
298-335), a dependency that is not vendored under /root/reference.  The format is restated here from
the files in ``data/*.chk``:

* one record per key; header line = key left-justified in 40 columns, ``kind=`` + 5-character kind,
  then the value (scalars) or the comma-separated shape (arrays);
* kinds ``str int flt bln none`` and arrays ``strar intar fltar blnar``;
* array data follow row-major, four values per line, each right-justified in 22 columns.
"""

import numpy as np

__all__ = ["load_chk", "dump_chk"]

_SCALAR = {
    "str": str,
    "int": int,
    "flt": float,
    "bln": lambda s: s.strip().lower() in ("true", "1", "t", "yes"),
}
_ARRAY = {"intar": np.int64, "fltar": np.float64, "blnar": np.bool_, "strar": None}


def load_chk(filename):
    """Return a dict with every record of a ``.chk`` file."""
    result = {}
    with open(filename, "r") as handle:
        lines = handle.read().split("\n")
    pos = 0
    nline = len(lines)
    while pos < nline:
        line = lines[pos]
        pos += 1
        if len(line.strip()) == 0:
            continue
        if len(line) < 52 or line[42:47] != "kind=":
            raise IOError("Malformed chk header at line %d: %r" % (pos, line))
        key = line[:40].strip()
        kind = line[47:52].strip()
        value = line[53:]
        if kind == "none":
            result[key] = None
        elif kind in _SCALAR:
            result[key] = _SCALAR[kind](value.strip() if kind != "str" else value.rstrip("\n"))
            if kind == "str":
                result[key] = value.strip()
        elif kind in _ARRAY:
            shape = tuple(int(w) for w in value.split(",") if w.strip() != "")
            count = int(np.prod(shape)) if len(shape) > 0 else 1
            words = []
            while len(words) < count:
                if pos >= nline:
                    raise IOError("Unexpected end of file while reading %s" % key)
                words.extend(lines[pos].split())
                pos += 1
            if len(words) != count:
                raise IOError("Wrong number of values for %s" % key)
            if kind == "strar":
                arr = np.array(words).reshape(shape)
            elif kind == "blnar":
                arr = np.array([w.lower() in ("true", "1", "t") for w in words], dtype=bool).reshape(shape)
            else:
                arr = np.array(words, dtype=_ARRAY[kind]).reshape(shape)
            result[key] = arr
        else:
            raise IOError("Unknown chk kind %r for key %s" % (kind, key))
    return result


def _fmt(value):
    if isinstance(value, (bool, np.bool_)):
        return "%22s" % bool(value)
    if isinstance(value, (int, np.integer)):
        return "%22d" % int(value)
    if isinstance(value, (float, np.floating)):
        return "% 22.15e" % float(value)
    return "%22s" % value


def dump_chk(filename, data):
    """Write a dict of scalars / arrays as a ``.chk`` file (keys sorted)."""
    with open(filename, "w") as handle:
        for key in sorted(data.keys()):
            value = data[key]
            if len(key) > 40:
                raise ValueError("chk keys are limited to 40 characters: %s" % key)
            head = "%-40s  kind=" % key
            if value is None:
                handle.write(head + "none  None\n")
            elif isinstance(value, str):
                handle.write(head + "str   %s\n" % value)
            elif isinstance(value, (bool, np.bool_)):
                handle.write(head + "bln   %s\n" % bool(value))
            elif isinstance(value, (int, np.integer)):
                handle.write(head + "int   %d\n" % int(value))
            elif isinstance(value, (float, np.floating)):
                handle.write(head + "flt   % 22.15e\n" % float(value))
            else:
                arr = np.asarray(value)
                if arr.dtype.kind in "iu":
                    kind = "intar"
                elif arr.dtype.kind == "f":
                    kind = "fltar"
                elif arr.dtype.kind == "b":
                    kind = "blnar"
                elif arr.dtype.kind in "US":
                    kind = "strar"
                else:
                    raise TypeError("Cannot store %s of dtype %s" % (key, arr.dtype))
                handle.write(head + "%s %s\n" % (kind, ",".join(str(n) for n in arr.shape)))
                flat = arr.ravel()
                for start in range(0, flat.size, 4):
                    handle.write(" ".join(_fmt(v) for v in flat[start:start + 4]) + "\n")
