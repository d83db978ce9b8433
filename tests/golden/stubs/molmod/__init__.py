"""Minimal stand-in for ``molmod`` (constants + units re-exported at top level, as molmod does)."""
from micmec_b200.units import *  # noqa: F401,F403
