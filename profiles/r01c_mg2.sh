#!/bin/bash
# 2-GPU check of HEAD (variant 14 default): parity of 2 slabs vs 1 GPU, the bench line exactly as the driver launches it
set -u
mkdir -p gpurun_out
T="timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 tests/multigpu_check.py 2>&1 | tail -3
$T --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 20 2>gpurun_out/p2_npt.err | tail -1 > gpurun_out/p2_npt_fused.json
$T --master-port 29514 bench.py --gpus 2 --ensemble nve --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/p2_nve_fused.json
$T --master-port 29515 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/p2_ref.json
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/p2_*.json')):
    try:
        d=json.load(open(f)); print(f, '%.3e'%d['value'], d['ms_per_step'], d.get('gpu_launches'), d['config']['parallelism'][:60], (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f, 'FAILED', e)
"
tail -5 gpurun_out/p2_npt.err
