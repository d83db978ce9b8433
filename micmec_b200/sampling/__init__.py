"""Phase-space sampling on the device (drop-in for the hot part of ``micmec.sampling``)."""
