#!/bin/bash
set -x
O=gpurun_out/r02p; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
for ens in npt nve; do
  $B --ensemble $ens > $O/bench_$ens.json 2> $O/bench_$ens.err
  python - <<PY
import json
d=json.load(open("$O/bench_$ens.json")); r=d["roofline"]
print("$ens", "ms/step %.4f" % d["ms_per_step"], "step_kernel %.4f" % r["kernel_ms"], "force %.4f" % r.get("force_only_kernel",{}).get("kernel_ms",0), "epot %.10e" % d["check"]["epot"])
PY
done
