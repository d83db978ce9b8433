#!/bin/bash
# round 2, call I (1 GPU): defaults after the tail became its own launch; 64^3; full GPU test suite
set -x
O=gpurun_out/r02i_1; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
run() { # tag args...
  tag=$1; shift
  $B "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_$tag.json"))
    r=d["roofline"]
    print("$tag", "ms/step %.4f" % d["ms_per_step"], "step_kernel %.4f" % r["kernel_ms"], "force %.4f" % r.get("force_only_kernel",{}).get("kernel_ms",0), d["config"]["kernel_tiling"], "launches", d["gpu_launches"], "epot %.10e" % d["check"]["epot"])
except Exception as e:
    print("$tag FAILED", e); print(open("$O/bench_$tag.err").read()[-1500:])
PY
}
run npt_default --ensemble npt
run npt_wrap1 --ensemble npt --wrap 1
run nve_default --ensemble nve
run nve_wrap1 --ensemble nve --wrap 1
run nvt64_default --ensemble nvt --grid 64
run nvt64_wrap1 --ensemble nvt --grid 64 --wrap 1
run nvt64_wrap1_tik --ensemble nvt --grid 64 --wrap 1 --tail-in-kernel 1
run npt64_default --ensemble npt --grid 64
run npt64_wrap1 --ensemble npt --grid 64 --wrap 1
