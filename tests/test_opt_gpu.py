"""QNOptimizer + DOF mappings on top of the CUDA force part: the reference's recorded optimiser runs
(tests/golden/opt_*.npz), every energy / gradient / virial through mm_compute."""
import pytest

import optcases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", optcases.CASES)
def test_qn_optimizer_matches_reference_on_gpu(tag):
    from micmec_b200.pes.mmff import ForcePartMechanical

    opt = optcases.run_case(tag, lambda system: ForcePartMechanical(system))
    assert opt.mmf.parts[0].launches > 0
