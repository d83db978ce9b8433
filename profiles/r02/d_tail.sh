#!/bin/bash
# round 2, call D: k_march2 with periodic images on load + fused tail; PIN variants; tail on/off
set -x
O=gpurun_out/r02d; mkdir -p $O
python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu.log 2>&1; tail -12 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
run() { # tag ens args...
  tag=$1; ens=$2; shift 2
  $B --ensemble $ens "$@" > $O/bench_${ens}_$tag.json 2> $O/bench_${ens}_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_${ens}_$tag.json"))
    r=d["roofline"]
    print("$ens $tag", "ms/step %.4f" % d["ms_per_step"], "step_kernel %.4f" % r["kernel_ms"], "force %.4f" % r.get("force_only_kernel",{}).get("kernel_ms",0), d["config"]["kernel_tiling"], "launches", d["gpu_launches"], "drift %.2e" % d["check"]["econs_drift"], "epot %.10e" % d["check"]["epot"])
except Exception as e:
    print("$ens $tag FAILED", e); print(open("$O/bench_${ens}_$tag.err").read()[-1500:])
PY
}
for ens in npt nve nvt; do
  run default $ens
  run notail $ens --tail 0
  run pinstep3 $ens --pin-step 3
  run pinforce0 $ens --pin-force 0
  run rpt1 $ens --rpt 1
  run rpt1pin3 $ens --rpt 1 --pin-step 3
done
run g64 nvt --grid 64
run g64 npt --grid 64
