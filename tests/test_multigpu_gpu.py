"""Multi-GPU parity of the z-slab decomposition as a driver-visible test: spawns tests/multigpu_check.py under
torch.distributed.run on 2, 4 and 8 GPUs of this box (each skipped when the box has fewer).  The script integrates a
perturbed grid as P z-slabs (NVE, NVT, NPT, 30 steps each) and rank 0 repeats the run on one GPU; gathered positions,
velocities, gradients and all scalars must agree to 1e-10 / 1e-8 (1-GPU vs P-GPU equality).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_count():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("nproc,shape,model", [(2, "40,24,32", "original"), (4, "40,24,32", "original"), (8, "40,24,64", "original"),
                                               (2, "40,24,32", "default")])
def test_slab_decomposition_matches_single_gpu(nproc, shape, model):
    if device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    env = dict(os.environ, MULTIGPU_SHAPE=shape, MULTIGPU_MODEL=model)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(29540 + nproc), os.path.join(ROOT, "tests", "multigpu_check.py")]
    proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    sys.stdout.write(proc.stdout[-4000:])
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert "multigpu ok (%d slabs" % nproc in proc.stdout
