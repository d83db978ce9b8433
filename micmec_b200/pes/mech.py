"""Alias module: BASELINE.json names ``micmec.pes.mech.ForcePartMechanical``; the class lives in ``mmff``."""
from .mmff import ForcePart, ForcePartMechanical, MicMecForceField  # noqa: F401

__all__ = ["MicMecForceField", "ForcePart", "ForcePartMechanical"]
