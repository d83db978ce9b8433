"""Test infrastructure: a ``ForcePart`` evaluated by the CPU oracle, so that the HOST logic of this package (hook
algebra, RNG order, integrator bookkeeping, optimisers) can be checked against the reference's golden trajectories
without a GPU.  Never part of the product path: the shipped ``ForcePartMechanical`` has no CPU code."""
import numpy as np

from micmec_b200.pes.mmff import ForcePart
from oracle import oracle as orc


class OracleForcePart(ForcePart):
    def __init__(self, system, model="original"):
        ForcePart.__init__(self, "micmec", system)
        self.system = system
        self.oracle = orc.Oracle(system, model=model)
        self.evaluations = 0

    def _internal_compute(self, gpos, vtens):
        e, g, v = self.oracle.compute(self.system.pos, np.array(self.system.domain.rvecs), gpos=gpos is not None,
                                      vtens=vtens is not None)
        self.evaluations += 1
        if gpos is not None:
            gpos[:] = g
        if vtens is not None:
            vtens[:] = v
        return e
