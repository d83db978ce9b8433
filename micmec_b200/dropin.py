"""Run the reference's scripts on this package without editing them.

    python -m micmec_b200.dropin simulations/md.py input.chk out.h5 -temp 300 -press 1

``install()`` registers this package's mirrors under the module names the reference uses, so that
``from micmec.pes.mmff import ForcePartMechanical`` or ``from micmec.sampling.nvt import LangevinThermostat`` in
user code resolve to the B200 implementations:

    micmec.system                          -> micmec_b200.system
    micmec.pes.mmff / micmec.pes.mech      -> micmec_b200.pes.mmff
    micmec.sampling.{verlet, nvt, npt, opt, dof, trajectory, iterative, utils} -> micmec_b200.sampling.*
    micmec.log                             -> micmec_b200.log
    molmod.units / molmod.constants        -> micmec_b200.units           (only when molmod itself is not installed)

When the real ``micmec`` package is importable it stays in place for everything else (builder, analysis,
``micmec.utils``) and only the modules above are replaced; otherwise a bare ``micmec`` namespace is created.
"""
import importlib
import runpy
import sys
import types

__all__ = ["install", "MIRRORS"]

MIRRORS = {
    "micmec.system": "micmec_b200.system",
    "micmec.log": "micmec_b200.log",
    "micmec.pes.mmff": "micmec_b200.pes.mmff",
    "micmec.pes.mech": "micmec_b200.pes.mmff",
    "micmec.sampling.iterative": "micmec_b200.sampling.iterative",
    "micmec.sampling.utils": "micmec_b200.sampling.utils",
    "micmec.sampling.verlet": "micmec_b200.sampling.verlet",
    "micmec.sampling.nvt": "micmec_b200.sampling.nvt",
    "micmec.sampling.npt": "micmec_b200.sampling.npt",
    "micmec.sampling.dof": "micmec_b200.sampling.dof",
    "micmec.sampling.opt": "micmec_b200.sampling.opt",
    "micmec.sampling.trajectory": "micmec_b200.sampling.trajectory",
}


def _package(name):
    """The real package when it imports, else an empty namespace registered under that name."""
    try:
        return importlib.import_module(name)
    except Exception:
        module = types.ModuleType(name)
        module.__path__ = []
        sys.modules[name] = module
        return module


_installed = None


def install():
    """Idempotent.  Returns the list of module names that now point at this package."""
    global _installed
    if _installed is not None:
        return list(_installed)
    installed = []
    for pkg in ("micmec", "micmec.pes", "micmec.sampling"):
        module = _package(pkg)
        if "." in pkg:
            setattr(sys.modules[pkg.rsplit(".", 1)[0]], pkg.rsplit(".", 1)[1], module)
    for alias, target in MIRRORS.items():
        module = importlib.import_module(target)
        sys.modules[alias] = module
        parent, leaf = alias.rsplit(".", 1)
        setattr(sys.modules[parent], leaf, module)
        installed.append(alias)
    try:
        import molmod.units  # noqa: F401
    except Exception:
        from . import units

        molmod = _package("molmod")
        for leaf in ("units", "constants"):
            sys.modules["molmod." + leaf] = units
            setattr(molmod, leaf, units)
            installed.append("molmod." + leaf)
        for name in units.__all__:
            if not hasattr(molmod, name):
                setattr(molmod, name, getattr(units, name))
    _installed = installed
    return list(installed)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        sys.stderr.write(__doc__)
        return 2
    install()
    from .log import log

    log.set_level(log.medium)  # scripts print their progress tables, like the reference's default log level
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
