"""``micmec_b200.dropin`` makes the reference's module names resolve to this package (checked in a subprocess so the
test session's own ``sys.modules`` stays clean)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = """
import sys
sys.path.insert(0, %r)
from micmec_b200 import dropin
names = dropin.install()
assert dropin.install() == names  # idempotent
from micmec.system import System
from micmec.pes.mmff import MicMecForceField, ForcePartMechanical
from micmec.sampling.verlet import VerletIntegrator, VerletScreenLog
from micmec.sampling.nvt import NHCThermostat, LangevinThermostat
from micmec.sampling.npt import MTKBarostat, TBCombination, LangevinBarostat
from micmec.sampling.opt import QNOptimizer, OptScreenLog
from micmec.sampling.dof import CartesianDOF, StrainCellDOF
from micmec.sampling.trajectory import HDF5Writer
from micmec.log import log
import micmec.sampling.nvt
for obj in (System, ForcePartMechanical, VerletIntegrator, LangevinThermostat, LangevinBarostat, QNOptimizer, StrainCellDOF, HDF5Writer):
    assert obj.__module__.startswith("micmec_b200."), obj
assert micmec.sampling.nvt is sys.modules["micmec_b200.sampling.nvt"]
print("ok", len(names))
"""


def test_reference_module_names_resolve_to_this_package():
    env = dict(os.environ, PYTHONPATH="")  # neither the reference nor the molmod stand-in on the path
    out = subprocess.run([sys.executable, "-c", CODE % ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("ok")
