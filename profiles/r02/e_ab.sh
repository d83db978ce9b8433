#!/bin/bash
# round 2, call E: A/B of k_march2 with the periodic images on load (wrap 1) vs from ghost nodes (wrap 0); ncu of both
set -x
O=gpurun_out/r02e; mkdir -p $O
python -m pytest tests/test_force_gpu.py tests/test_md_gpu.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
run() { # tag ens args...
  tag=$1; ens=$2; shift 2
  $B --ensemble $ens "$@" > $O/bench_${ens}_$tag.json 2> $O/bench_${ens}_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_${ens}_$tag.json"))
    r=d["roofline"]
    print("$ens $tag", "ms/step %.4f" % d["ms_per_step"], "step_kernel %.4f" % r["kernel_ms"], "force %.4f" % r.get("force_only_kernel",{}).get("kernel_ms",0), d["config"]["kernel_tiling"], "launches", d["gpu_launches"], "epot %.10e" % d["check"]["epot"])
except Exception as e:
    print("$ens $tag FAILED", e); print(open("$O/bench_${ens}_$tag.err").read()[-1500:])
PY
}
for ens in nve npt; do
  run wrap0 $ens --wrap 0
  run wrap1notail $ens --wrap 1 --tail 0
  run wrap1tail $ens --wrap 1 --tail 1
done
for wv in 0 1; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_march2 -s 6 -c 1 -o $O/nve_wrap$wv \
    python bench.py --ensemble nve --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --wrap $wv --tail 0 > $O/ncu_nve_wrap$wv.log 2>&1; tail -2 $O/ncu_nve_wrap$wv.log | cut -c1-200
done
