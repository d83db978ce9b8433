#!/bin/bash
# 2-GPU parity + bench of the slab exchange modes (run on a box with >= 2 GPUs): fused halo / peer inboxes / NCCL
set -u
T="timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 tests/multigpu_check.py 2>&1 | tail -3
MICMEC_B200_FUSED=0 $T --master-port 29512 tests/multigpu_check.py 2>&1 | tail -1
for e in nve npt; do
  $T --master-port 29513 bench.py --gpus 2 --ensemble $e --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/p2_${e}_fused.json
  MICMEC_B200_FUSED=0 $T --master-port 29514 bench.py --gpus 2 --ensemble $e --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/p2_${e}_peer.json
done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/p2_*.json')):
    try:
        d=json.load(open(f)); print(f, '%.3e'%d['value'], d['ms_per_step'], d.get('gpu_launches'))
    except Exception as e: print(f, 'FAILED', e)
"
