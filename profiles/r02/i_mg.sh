#!/bin/bash
# round 2, call I (N GPUs): slab parity test + strong-scaling bench lines, fused tail / images on load vs separate launches
set -x
N=${1:-2}
O=gpurun_out/r02i_$N; mkdir -p $O
shape=40,24,32; [ $N -eq 8 ] && shape=40,24,64
MULTIGPU_SHAPE=$shape timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > $O/multigpu_check.log 2>&1; tail -6 $O/multigpu_check.log
run() { # tag args...
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 10 --no-e2e "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_$tag.json") if l.startswith("{")][-1])
    r=d["roofline"]; c=d["check"]
    print("N=$N $tag", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "step_kernel %.4f" % r["kernel_ms"], "force %.4f" % r.get("force_only_kernel",{}).get("kernel_ms",0), d["config"]["kernel_tiling"], "launches", d["gpu_launches"], "epot %.12e ekin %.12e econs %.12e rv0 %.12e" % (c["epot"], c["ekin"], c["econs"], c["rvecs"][0]))
except Exception as e:
    print("N=$N $tag FAILED", e); print(open("$O/bench_$tag.err").read()[-1500:])
PY
}
run npt_default --ensemble npt


run nve_default --ensemble nve


