#!/bin/bash
# round 2, call J: structured path of the eight-corner (`default`) model; full GPU test suite; scalar-kernel latency after unrolling
set -x
O=gpurun_out/r02j; mkdir -p $O
python -m pytest tests -m gpu -x -q --durations=6 > $O/pytest_gpu.log 2>&1; tail -14 $O/pytest_gpu.log
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e"
run() { # tag args...
  tag=$1; shift
  $B "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_$tag.json"))
    r=d["roofline"]
    print("$tag", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "step_kernel %.4f" % r["kernel_ms"], "force %.4f" % r.get("force_only_kernel",{}).get("kernel_ms",0), d["config"]["kernel_tiling"], "launches", d["gpu_launches"], "epot %.10e" % d["check"]["epot"])
except Exception as e:
    print("$tag FAILED", e); print(open("$O/bench_$tag.err").read()[-1500:])
PY
}
run npt_original --ensemble npt
run npt_default_model --ensemble npt --model default
run nve_default_model --ensemble nve --model default
run npt_default_model_indexed --ensemble npt --model default --generic --grid 128
run npt_default_model_128 --ensemble npt --model default --grid 128
