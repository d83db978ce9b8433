#!/bin/bash
set -x
N=4
O=gpurun_out/r02t_$N; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; head -12 $O/topo.txt
run() { # tag args...
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 20 --no-e2e --no-cpu-baseline "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_$tag.json") if l.startswith("{")][-1])
    r=d["roofline"]
    print("N=$N $tag", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "step_kernel %.4f" % r["kernel_ms"], "force %.4f" % r.get("force_only_kernel",{}).get("kernel_ms",0), d["config"]["kernel_tiling"]["tail"], "launches", d["gpu_launches"])
except Exception as e:
    print("N=$N $tag FAILED", e); print(open("$O/bench_$tag.err").read()[-1500:])
PY
}
run npt_default --ensemble npt
run npt_default_again --ensemble npt
run npt_wrap0 --ensemble npt --wrap 0
run npt_tik --ensemble npt --tail-in-kernel 1
run npt_notail --ensemble npt --tail 0
