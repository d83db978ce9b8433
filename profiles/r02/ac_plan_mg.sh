#!/bin/bash
set -x
N=${1:-2}
O=gpurun_out/r02ac_$N; mkdir -p $O
run() { # tag args...
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 20 --no-e2e --no-cpu-baseline "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_$tag.json") if l.startswith("{")][-1])
    r=d["roofline"]; t=d["config"]["kernel_tiling"]
    print("N=$N $tag", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "step_kernel %.4f" % r["kernel_ms"], "force %.4f" % r.get("force_only_kernel",{}).get("kernel_ms",0), "blocks", t["blocks"], "eff", t["schedule_efficiency"], "epot %.12e" % d["check"]["epot"])
except Exception as e:
    print("N=$N $tag FAILED", e); print(open("$O/bench_$tag.err").read()[-1500:])
PY
}
run npt_plan1 --ensemble npt
run npt_plan0 --ensemble npt --plan 0
run nve_plan1 --ensemble nve
run nve_plan0 --ensemble nve --plan 0
