#!/bin/bash
# round 2, call AF (8 GPUs): slab parity at 8 slabs and the driver's NPT scaling command with the two-class block schedule
set -x
O=gpurun_out/r02af_8; mkdir -p $O
timeout 150 python -m pytest tests/test_multigpu_gpu.py -q -k "8-40" > $O/pytest_mg8.log 2>&1; tail -4 $O/pytest_mg8.log
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench.json 2> $O/bench.err
python -c "import json; d=json.load(open('$O/bench.json')); r=d['roofline']; t=d['config']['kernel_tiling']; print('N=8 npt ms/step %.4f value %.3e step_kernel %.4f force %.4f blocks %s eff %s epot %.12e' % (d['ms_per_step'], d['value'], r['kernel_ms'], (r.get('force_only_kernel') or {}).get('kernel_ms', 0), t.get('blocks'), t.get('schedule_efficiency'), d['check']['epot']))"
