#!/bin/bash
# N-GPU parity + strong-scaling bench lines of the fused halo exchange: usage  profiles/mg_scale.sh N
set -u
N=${1:-4}
T="timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29521 tests/multigpu_check.py 2>&1 | tail -2
for e in nve npt; do
  $T --master-port 29523 bench.py --gpus $N --ensemble $e --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/p${N}_${e}_fused.json
done
MICMEC_B200_PEER=0 $T --master-port 29524 bench.py --gpus $N --ensemble nve --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/p${N}_nve_nccl.json
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/p${N}_*.json')):
    try:
        d=json.load(open(f)); print(f, '%.3e'%d['value'], d['ms_per_step'], d.get('gpu_launches'))
    except Exception as e: print(f, 'FAILED', e)
"
