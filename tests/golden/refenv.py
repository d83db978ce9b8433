"""Make the UNMODIFIED reference importable in the build container (never on the GPU box).

Puts the molmod/h5py stand-ins and /root/reference on ``sys.path`` and lets ``micmec.pes.ext``
(the reference's Cython+C Domain) resolve to the copy built by ``make -C oracle ref`` in
``oracle/_ref/``.  Used only by ``make_golden.py`` and by ``-m "not gpu"`` tests that are skipped
when /root/reference is absent.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("MICMEC_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE, "micmec"))


def setup():
    """Return the imported reference ``micmec`` package."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE)
    for path in (REFERENCE, os.path.join(HERE, "stubs"), ROOT):
        if path not in sys.path:
            sys.path.insert(0, path)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import micmec.pes

    refdir = os.path.join(ROOT, "oracle", "_ref")
    if refdir not in micmec.pes.__path__:
        micmec.pes.__path__.append(refdir)
    import micmec.pes.ext  # noqa: F401  (fails loudly if `make -C oracle ref` was not run)
    import micmec

    return micmec


def use_model(model):
    """Rebind the per-cell functions mmff.py looks up at call time (mmff.py:380-385)."""
    import importlib
    from micmec.pes import mmff

    module = {
        "original": "micmec.pes.nanocell_original",
        "default": "micmec.pes.nanocell",
    }[model]
    mod = importlib.import_module(module)
    mmff.elastic_energy_nanocell = mod.elastic_energy_nanocell
    mmff.grad_elastic_energy_nanocell = mod.grad_elastic_energy_nanocell
