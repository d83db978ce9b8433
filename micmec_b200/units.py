"""Physical constants and unit factors in atomic units.

The reference takes these from the third-party package ``molmod`` (>= 1.4.1, not vendored under
/root/reference): ``molmod.boltzmann`` is used at micmec/pes/mmff.py:33,390,393,
micmec/sampling/verlet.py:175, micmec/sampling/nvt.py:397-454 and micmec/sampling/npt.py:596,739;
the unit factors are used by the ``simulations/*.py`` scripts.  The values below are the CODATA-2002
based numbers molmod ships; ``angstrom`` and ``pascal`` reproduce the figures stored in the
reference fixtures (10 A = 18.89726133921252 bohr, 50 GPa = 1.699465791378921e-3 a.u.).
"""

# Boltzmann constant in hartree / kelvin.
boltzmann = 3.1668154051341965e-06

kelvin = 1.0
# time
second = 1.0 / 2.418884326500e-17
femtosecond = 1e-15 * second
picosecond = 1e-12 * second
# length
meter = 1.0 / 0.5291772083e-10
angstrom = 1e-10 * meter
nanometer = 1e-9 * meter
# energy
joule = 1.0 / 4.35974381e-18
avogadro = 6.0221415e23
kjmol = 1.0e3 * joule / avogadro
electronvolt = 1.0 / 27.2113845
# pressure
pascal = joule / meter**3
bar = 1e5 * pascal
# mass
amu = 1e-3 / avogadro / 9.1093826e-31

__all__ = [name for name in dir() if not name.startswith("_")]
