"""Micromechanical potential energy surface on the GPU (drop-in for ``micmec.pes``)."""
