"""A script written against the reference's module names (tests/scripts/md_like.py, the call pattern of
simulations/md.py) runs unedited through ``python -m micmec_b200.dropin`` on the GPU: NVE stays device resident,
Langevin NVT / NPT run host-driven with every force evaluation on the GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("flags, device_mode", [([], True), (["-temp", "300"], False), (["-temp", "300", "-press", "1"], False)])
def test_md_script_runs_unedited(tmp_path, flags, device_mode):
    from micmec_b200.system import System
    from micmec_b200.celltypes import TYPE_FCU

    system = System.periodic_grid((3, 3, 3), TYPE_FCU, explicit=True)
    system.pos[:] = system.pos + 0.2 * np.random.default_rng(0).standard_normal(system.pos.shape)
    chk = str(tmp_path / "in.chk")
    system.to_file(chk)
    out_fn = str(tmp_path / "out.h5")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests", "stubs_h5")]))
    cmd = [sys.executable, "-m", "micmec_b200.dropin", os.path.join(ROOT, "tests", "scripts", "md_like.py"), chk, out_fn,
           "-steps", "20"] + flags
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "md_like done: device_mode=%s counter=20" % device_mode in out.stdout, out.stdout
    traj = np.load(out_fn + ".npz")
    assert traj["trajectory/pos"].shape == (5, 27, 3) and list(traj["trajectory/counter"]) == [0, 5, 10, 15, 20]
    assert np.all(np.isfinite(traj["trajectory/epot"])) and np.all(traj["trajectory/temp"] < 2000.0)
    assert traj["system/pos"].shape == (27, 3)
