#!/bin/bash
# round 2, call G: TMA issue spread over the warps, cheaper ticket; A/B as in call F
set -x
O=gpurun_out/r02g; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
run() { # tag ens args...
  tag=$1; ens=$2; shift 2
  $B --ensemble $ens "$@" > $O/bench_${ens}_$tag.json 2> $O/bench_${ens}_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_${ens}_$tag.json"))
    r=d["roofline"]
    print("$ens $tag", "ms/step %.4f" % d["ms_per_step"], "step_kernel %.4f" % r["kernel_ms"], "force %.4f" % r.get("force_only_kernel",{}).get("kernel_ms",0), "tail", d["config"]["kernel_tiling"]["fused_tail"], "launches", d["gpu_launches"], "epot %.10e" % d["check"]["epot"])
except Exception as e:
    print("$ens $tag FAILED", e); print(open("$O/bench_${ens}_$tag.err").read()[-1500:])
PY
}
for ens in npt nve; do
  run default $ens
  run unroll2 $ens --unroll 2
  run notail $ens --tail 0
  run wrap0 $ens --wrap 0
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_march2 -s 10 -c 3 -o /tmp/npt_default \
    python bench.py --ensemble npt --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_npt.log 2>&1; tail -2 $O/ncu_npt.log | cut -c1-200
ncu -i /tmp/npt_default.ncu-rep --page raw --csv > $O/npt_default_raw.csv 2>/dev/null
ncu -i /tmp/npt_default.ncu-rep --page source --csv > $O/npt_default_source.csv 2>/dev/null
ls -la $O | head -5
