"""The host-driven Verlet hooks on top of the CUDA force part: same reference trajectories as test_hooks_cpu.py,
every force evaluation through mm_compute."""
import pytest

import hookcases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", hookcases.CASES)
def test_hook_trajectory_matches_reference_on_gpu(tag):
    from micmec_b200.pes.mmff import ForcePartMechanical

    verlet = hookcases.run_case(tag, lambda system: ForcePartMechanical(system))
    assert verlet.mmf.parts[0].launches > 0
