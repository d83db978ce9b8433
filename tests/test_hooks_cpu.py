"""Host logic of the Verlet hooks that run in host-driven mode (Langevin / Berendsen / CSVR / Andersen / GLE
thermostats, Langevin / Berendsen barostats, MTK with a non-NHC thermostat or its own chain) against trajectories
recorded from the UNMODIFIED reference with the same RNG seed.  Forces come from the CPU oracle here (test
infrastructure); tests/test_hooks_gpu.py repeats the cases on the CUDA force part."""
import pytest

import hookcases
from oraclepart import OracleForcePart


@pytest.mark.parametrize("tag", hookcases.CASES)
def test_hook_trajectory_matches_reference(tag):
    hookcases.run_case(tag, lambda system: OracleForcePart(system))


def test_tbcombination_rejects_unsupported_pairs_and_orders_arguments():
    import numpy as np
    from micmec_b200.sampling import nvt, npt
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU
    from micmec_b200.pes.mmff import MicMecForceField
    from micmec_b200.units import pascal

    system = System.periodic_grid((2, 2, 2), TYPE_FCU, explicit=True)
    mmf = MicMecForceField(system, [OracleForcePart(system)])
    baro = npt.LangevinBarostat(mmf, 300.0, 1e6 * pascal)
    thermo = nvt.LangevinThermostat(300.0)
    tbc = npt.TBCombination(baro, thermo)  # swapped on purpose (npt.py:69-77)
    assert tbc.thermostat is thermo and tbc.barostat is baro and not tbc.native
    with pytest.raises(TypeError):
        npt.TBCombination(nvt.CSVRThermostat(300.0), baro)
    native = npt.TBCombination(nvt.NHCThermostat(300.0), npt.MTKBarostat(mmf, 300.0, 1e6 * pascal))
    assert native.native
    own = npt.MTKBarostat(mmf, 300.0, 1e6 * pascal, baro_thermo=nvt.NHCThermostat(300.0))
    assert not own.native and not npt.TBCombination(nvt.NHCThermostat(300.0), own).native
    assert np.allclose(np.array(system.domain.rvecs), np.array(system.domain.rvecs).T)  # symmetrised by the barostats
