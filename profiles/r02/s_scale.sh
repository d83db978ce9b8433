#!/bin/bash
# round 2, call S (N GPUs): the driver's own scaling commands (bench.py --gpus N --steps 200 --warmup 20), both arms, + config 5
set -x
N=${1:-2}
O=gpurun_out/r02s_$N; mkdir -p $O
if [ $N -eq 1 ]; then
  python bench.py --gpus 1 --steps 200 --warmup 20 > $O/bench.json 2> $O/bench.err
  python bench.py --config5 10240 > $O/config5.json 2> $O/config5.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 200 --warmup 20 > $O/bench.json 2> $O/bench.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 200 --warmup 20 --ensemble nve --no-cpu-baseline > $O/bench_nve.json 2> $O/bench_nve.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --config5 10240 > $O/config5.json 2> $O/config5.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        c = d.get("check", {})
        print(f.split("/")[-1], "N=%s" % d.get("n_gpus"), "value %.4e %s" % (d["value"], d["unit"]), "ms/step %.4f" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), {k: c[k] for k in c if k in ("epot", "ekin", "econs", "converged", "failed")}, (c.get("rvecs") or [None])[0])
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-1200:])
PY
