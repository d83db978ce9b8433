#!/bin/bash
# round 2, call C: first run of k_march2 - GPU tests on the default tiling, then the tile configurations side by side
set -x
O=gpurun_out/r02c; mkdir -p $O
python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; tail -15 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
for cfg in "--rpt 2" "--rpt 1" "--rpt 3" "--rpt 2 --tile-rows 4" "--march2 0"; do
  tag=$(echo $cfg | tr -d ' -')
  for ens in npt nve; do
    $B --ensemble $ens $cfg > $O/bench_${ens}_$tag.json 2> $O/bench_${ens}_$tag.err
    python - <<PY
import json
try:
    d=json.load(open("$O/bench_${ens}_$tag.json"))
    r=d["roofline"]
    print("$ens $cfg", "ms/step %.4f" % d["ms_per_step"], "step_kernel %.4f" % r["kernel_ms"], "force %.4f" % r.get("force_only_kernel",{}).get("kernel_ms",0), d["config"]["kernel_tiling"], "drift %.2e" % d["check"]["econs_drift"], "epot %.10e" % d["check"]["epot"])
except Exception as e:
    print("$ens $cfg FAILED", e); print(open("$O/bench_${ens}_$tag.err").read()[-800:])
PY
  done
done
