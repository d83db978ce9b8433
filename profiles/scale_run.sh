#!/bin/bash
# strong-scaling sweep of bench.py on one box: usage  scale_run.sh "1 2 4 8" npt
set -u
NS=${1:-"1 2 4 8"}; ENS=${2:-npt}; PORT=29600
for n in $NS; do
  PORT=$((PORT+1))
  if [ "$n" = "1" ]; then
    python bench.py --ensemble $ENS --no-e2e --no-cpu-baseline 2>&1 | tail -1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $n --ensemble $ENS --no-e2e --no-cpu-baseline 2>&1 | tail -1
  fi
done
