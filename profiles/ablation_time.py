#!/usr/bin/env python
"""Time the MD step of ablated library builds (profiles/ablation.sh) - run on the GPU box:
    python profiles/ablation_time.py nve 0 "0 1 2 4 8 16 32"
Each build runs in its own process (MICMEC_B200_LIB); numerics of ablated builds are meaningless, so the NaN check of
bench.py is bypassed by timing mm_md_run directly."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, ctypes, numpy as np, torch
sys.path.insert(0, %r)
import bench
from micmec_b200 import _lib
from micmec_b200.pes.mmff import MicMecForceField, ForcePartMechanical
from micmec_b200.sampling.verlet import VerletIntegrator
from micmec_b200.sampling.nvt import NHCThermostat
from micmec_b200.sampling.npt import MTKBarostat, TBCombination
ens, variant, grid = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lib = _lib.load()
p = bench.md_params(ens)
system, vel0 = bench.make_state(grid)
part = ForcePartMechanical(system, model="original", device=0)
_lib.check(lib.mm_set_option(part.handle, b"variant", variant))
mmf = MicMecForceField(system, [part])
stream = torch.cuda.Stream()
_lib.check(lib.mm_set_stream(part.handle, ctypes.c_void_p(stream.cuda_stream)))
hooks = []
thermo = NHCThermostat(p["temp"], timecon=p["timecon_thermo"], chain_vel0=p["chain_vel0"], chain_pos0=np.zeros(3), restart=True) if p["thermo"] else None
baro = MTKBarostat(mmf, p["temp"], p["press"], timecon=p["timecon_baro"], vel_press0=p["vel_press0"], restart=True) if p["baro"] else None
if thermo is not None and baro is not None: hooks.append(TBCombination(thermo, baro))
elif thermo is not None: hooks.append(thermo)
verlet = VerletIntegrator(mmf, timestep=p["timestep"], hooks=hooks, vel0=vel0)
md = verlet._md
lib.mm_md_run(md, 3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); lib.mm_md_run(md, 20); e1.record(stream); torch.cuda.synchronize()
print("MS_PER_STEP %%.4f" %% (e0.elapsed_time(e1) / 20))
''' % ROOT

if __name__ == "__main__":
    ens, variant = sys.argv[1], sys.argv[2]
    grid = os.environ.get("GRID", "256")
    for a in sys.argv[3].split():
        env = dict(os.environ, MICMEC_B200_LIB=os.path.join(ROOT, "profiles", "ablate", "lib_%s.so" % a))
        try:  # an ablated build may hang (e.g. no barriers + staged loads): bound every child
            out = subprocess.run([sys.executable, "-c", CHILD, ens, variant, grid], env=env, capture_output=True, text=True, timeout=90)
        except subprocess.TimeoutExpired:
            print(json.dumps({"ablate": int(a), "ensemble": ens, "variant": int(variant), "ms_per_step": None, "err": "timeout"}), flush=True)
            continue
        ms = [l for l in out.stdout.splitlines() if l.startswith("MS_PER_STEP")]
        print(json.dumps({"ablate": int(a), "ensemble": ens, "variant": int(variant), "ms_per_step": float(ms[0].split()[1]) if ms else None,
                          "err": None if ms else out.stderr[-300:]}), flush=True)
